"""ctypes mirror of the POD types and enums declared in include/skeres.h.

Only layouts live here (no library loading), so both the product binding (`_lib.py`) and the
test-side oracle binding (`tests/oracle_lib.py`) can share them.
"""
import ctypes as C

ABI_VERSION = 3

# sk_status
OK, ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_UNSUPPORTED, ERR_NCCL, ERR_IO, ERR_INTERNAL = range(7)

# sk_functor_id
FUNCTOR_SNAVELY_REPROJECTION_ERROR = 1
FUNCTOR_EXPONENTIAL_RESIDUAL = 2
FUNCTOR_HELLO_WORLD = 3
FUNCTOR_POWELL_F1, FUNCTOR_POWELL_F2, FUNCTOR_POWELL_F3, FUNCTOR_POWELL_F4 = 4, 5, 6, 7
FUNCTOR_POWELL_ANALYTIC_F2 = 8
FUNCTOR_TEST_BILINEAR_SCALAR = 100
FUNCTOR_TEST_BILINEAR_VECTOR3 = 101
FUNCTOR_TEST_SUM10 = 102

MAX_PARAMETER_BLOCKS = 10
MAX_CONSTS = 4

# sk_loss_type
LOSS_TRIVIAL, LOSS_HUBER, LOSS_CAUCHY, LOSS_TOLERANT = 0, 1, 2, 3

# ceres/types.h enumerators (ceres.i:137)
DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR = range(7)
IDENTITY, JACOBI, SCHUR_JACOBI, CLUSTER_JACOBI, CLUSTER_TRIDIAGONAL = range(5)
LINE_SEARCH, TRUST_REGION = 0, 1
LEVENBERG_MARQUARDT, DOGLEG = 0, 1
CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE = range(5)

LINEAR_SOLVER_NAMES = {
    DENSE_NORMAL_CHOLESKY: "DENSE_NORMAL_CHOLESKY", DENSE_QR: "DENSE_QR",
    SPARSE_NORMAL_CHOLESKY: "SPARSE_NORMAL_CHOLESKY", DENSE_SCHUR: "DENSE_SCHUR",
    SPARSE_SCHUR: "SPARSE_SCHUR", ITERATIVE_SCHUR: "ITERATIVE_SCHUR", CGNR: "CGNR",
}
TERMINATION_NAMES = {
    CONVERGENCE: "CONVERGENCE", NO_CONVERGENCE: "NO_CONVERGENCE", FAILURE: "FAILURE",
    USER_SUCCESS: "USER_SUCCESS", USER_FAILURE: "USER_FAILURE",
}

KF_NAMES = ["evaluate_jacobian", "evaluate_cost", "schur_setup", "schur_matvec", "pcg_vector",
            "back_substitute", "dense", "lm", "comm", "pcg_solve"]
KF_COUNT = len(KF_NAMES)

COMM_UNIQUE_ID_BYTES = 128


class DoublePointer(C.Structure):
    """sk_double_pointer: (array handle, offset) — the reference's SWIGTYPE_p_double."""
    _fields_ = [("array", C.c_void_p), ("offset", C.c_int64)]


class SolverOptions(C.Structure):
    """sk_solver_options (Solver.Options, ceres.i:151)."""
    _fields_ = [
        ("minimizer_type", C.c_int32),
        ("trust_region_strategy_type", C.c_int32),
        ("linear_solver_type", C.c_int32),
        ("preconditioner_type", C.c_int32),
        ("max_num_iterations", C.c_int32),
        ("max_num_consecutive_invalid_steps", C.c_int32),
        ("min_linear_solver_iterations", C.c_int32),
        ("max_linear_solver_iterations", C.c_int32),
        ("jacobi_scaling", C.c_int32),
        ("minimizer_progress_to_stdout", C.c_int32),
        ("num_threads", C.c_int32),
        ("profile_kernels", C.c_int32),
        ("initial_trust_region_radius", C.c_double),
        ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double),
        ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double),
        ("max_lm_diagonal", C.c_double),
        ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("eta", C.c_double),
        ("max_solver_time_in_seconds", C.c_double),
        ("comm", C.c_void_p),
        ("residual_blocks_are_local", C.c_int32),
        ("reserved_", C.c_int32),
    ]


def default_options() -> SolverOptions:
    """The Ceres 1.x defaults written by sk_solver_options_init, in pure Python
    so the oracle binding does not need the product library."""
    o = SolverOptions()
    o.minimizer_type = TRUST_REGION
    o.trust_region_strategy_type = LEVENBERG_MARQUARDT
    o.linear_solver_type = SPARSE_NORMAL_CHOLESKY
    o.preconditioner_type = JACOBI
    o.max_num_iterations = 50
    o.max_num_consecutive_invalid_steps = 5
    o.min_linear_solver_iterations = 0
    o.max_linear_solver_iterations = 500
    o.jacobi_scaling = 1
    o.minimizer_progress_to_stdout = 0
    o.num_threads = 1
    o.profile_kernels = 0
    o.initial_trust_region_radius = 1e4
    o.max_trust_region_radius = 1e16
    o.min_trust_region_radius = 1e-32
    o.min_relative_decrease = 1e-3
    o.min_lm_diagonal = 1e-6
    o.max_lm_diagonal = 1e32
    o.function_tolerance = 1e-6
    o.gradient_tolerance = 1e-10
    o.parameter_tolerance = 1e-8
    o.eta = 1e-1
    o.max_solver_time_in_seconds = 1e9
    o.comm = None
    return o


class IterationSummary(C.Structure):
    """sk_iteration_summary (ceres IterationSummary)."""
    _fields_ = [
        ("iteration", C.c_int32),
        ("step_is_valid", C.c_int32),
        ("step_is_nonmonotonic", C.c_int32),
        ("step_is_successful", C.c_int32),
        ("linear_solver_iterations", C.c_int32),
        ("reserved_", C.c_int32),
        ("cost", C.c_double),
        ("cost_change", C.c_double),
        ("gradient_max_norm", C.c_double),
        ("gradient_norm", C.c_double),
        ("step_norm", C.c_double),
        ("relative_decrease", C.c_double),
        ("trust_region_radius", C.c_double),
        ("eta", C.c_double),
        ("iteration_time_in_seconds", C.c_double),
        ("cumulative_time_in_seconds", C.c_double),
    ]


class SolverSummaryData(C.Structure):
    """sk_solver_summary_data (Solver.Summary)."""
    _fields_ = [
        ("termination_type", C.c_int32),
        ("num_successful_steps", C.c_int32),
        ("num_unsuccessful_steps", C.c_int32),
        ("num_iterations", C.c_int32),
        ("linear_solver_type_used", C.c_int32),
        ("preconditioner_type_used", C.c_int32),
        ("num_gpus", C.c_int32),
        ("reserved_", C.c_int32),
        ("initial_cost", C.c_double),
        ("final_cost", C.c_double),
        ("fixed_cost", C.c_double),
        ("num_parameter_blocks", C.c_int64),
        ("num_parameters", C.c_int64),
        ("num_residual_blocks", C.c_int64),
        ("num_residuals", C.c_int64),
        ("num_residual_evaluations", C.c_int64),
        ("num_jacobian_evaluations", C.c_int64),
        ("num_linear_solves", C.c_int64),
        ("total_linear_solver_iterations", C.c_int64),
        ("num_kernel_launches", C.c_int64),
        ("total_time_in_seconds", C.c_double),
        ("preprocessor_time_in_seconds", C.c_double),
        ("minimizer_time_in_seconds", C.c_double),
        ("minimizer_device_time_in_seconds", C.c_double),
        ("kernel_ms", C.c_double * KF_COUNT),
        ("kernel_launches", C.c_int64 * KF_COUNT),
    ]
