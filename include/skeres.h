/*
 * skeres.h — C ABI of libskeres.so, the B200-native Levenberg–Marquardt solver core that
 * replaces the libceres + SWIG/JNI stack behind the Scala API of fgcallari/skeres.
 *
 * Every entry point is plain `extern "C"` with pointers and sizes only, callable from JNI or
 * Panama (see INTEGRATION.md for the reference-side stubs). Each declaration cites the
 * reference interface it replaces as  path:line  relative to the skeres source tree.
 *
 * Conventions
 *   - every function returning `int` returns an sk_status (0 = SK_OK); the message for the last
 *     failure on the calling thread is available from sk_last_error();
 *   - no exception ever crosses this boundary (the reference never lets one cross JNI either);
 *   - handles are opaque and owned by the library until the matching *_destroy call: nothing is
 *     freed as a garbage-collection side effect (contrast ceres.i:160-167 `%newobject`);
 *   - all parameter memory is resident in device (HBM) memory.  A "double pointer" of the
 *     reference (`SWIGTYPE_p_double`, package.scala:11-13) becomes (array handle, offset);
 *   - the library has NO CPU fallback: every numeric entry point needs an sm_100a GPU and
 *     fails with SK_ERR_CUDA when none is present.
 *   - exports are not re-entrant per handle; one host thread drives one solve.
 */
#ifndef SKERES_H_
#define SKERES_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKERES_ABI_VERSION 3

typedef enum sk_status {
  SK_OK = 0,
  SK_ERR_INVALID_ARGUMENT = 1,  /* bad handle, index out of range, size mismatch            */
  SK_ERR_CUDA = 2,              /* no device, allocation failure, launch failure            */
  SK_ERR_UNSUPPORTED = 3,       /* unregistered functor / option without a device path      */
  SK_ERR_NCCL = 4,              /* communicator failure                                     */
  SK_ERR_IO = 5,                /* BAL file could not be read                               */
  SK_ERR_INTERNAL = 6
} sk_status;

/* Message describing the last error raised on this thread ("" when none). */
const char* sk_last_error(void);
int sk_abi_version(void);
/* Number of CUDA devices visible (0 when no driver / device); never fails. */
int sk_device_count(void);
/* Binds the calling thread's solves to a device ordinal (default 0). */
int sk_set_device(int ordinal);

/* ceres.i:131-135  initGoogleLogging(name) — kept so the examples' first line still links;
 * it records the name and does nothing else (there is no glog here). */
void sk_init_google_logging(const char* name);

/* ------------------------------------------------------------------------------------------
 * DoubleArray — ceres.i:95-96 (%array_class(double, DoubleArray)), RichDoubleArray.scala:14-75
 * Device-resident array of doubles.  Element accessors exist for drop-in use (each is a
 * device round trip, exactly as each one is a JNI crossing today: RichDoubleArray.scala:20);
 * the bulk upload/download calls are the real path.
 * ---------------------------------------------------------------------------------------- */
typedef struct sk_double_array sk_double_array;

int sk_double_array_create(int64_t n, sk_double_array** out);
int sk_double_array_destroy(sk_double_array* a);
int64_t sk_double_array_size(const sk_double_array* a);
int sk_double_array_upload(sk_double_array* a, int64_t offset, const double* host, int64_t n);
int sk_double_array_download(const sk_double_array* a, int64_t offset, double* host, int64_t n);
int sk_double_array_get(const sk_double_array* a, int64_t i, double* out); /* getitem */
int sk_double_array_set(sk_double_array* a, int64_t i, double value);      /* setitem */
/* Device-to-device copy between (or inside) DoubleArrays, e.g. to restore a saved start point. */
int sk_double_array_copy(sk_double_array* dst, int64_t dst_offset, const sk_double_array* src, int64_t src_offset, int64_t n);
/* Raw device address (for callers that share the CUDA context, e.g. a benchmark harness). */
void* sk_double_array_device_ptr(sk_double_array* a);

/* DoubleArraySlice::get(double*, int) — ceres.i:99-107; RichDoubleArray.slice (:52).
 * An interior pointer: the parameter block starting at `offset` inside `array`. */
typedef struct sk_double_pointer {
  sk_double_array* array; /* NULL plays the role of a NULL double* (DoubleMatrix.row, ceres.i:117) */
  int64_t offset;
} sk_double_pointer;

/* ------------------------------------------------------------------------------------------
 * LossFunction — ceres.i:160-184 PredefinedLossFunctions
 * trivial / huber / cauchy / tolerant have device implementations (Corrector applied inside the
 * evaluator kernel).  The other factories report SK_ERR_UNSUPPORTED.
 * ---------------------------------------------------------------------------------------- */
typedef struct sk_loss_function sk_loss_function;
typedef enum sk_loss_type {
  SK_LOSS_TRIVIAL = 0, /* ceres.i:170 */
  SK_LOSS_HUBER = 1,   /* ceres.i:171 */
  SK_LOSS_CAUCHY = 2,  /* ceres.i:173 */
  SK_LOSS_TOLERANT = 3 /* ceres.i:175: rho(s) = b log(1 + e^((s - a) / b)) - b log(1 + e^(-a / b)); rho'' > 0, so it is the
                          one registered loss that drives the Corrector through its alpha != 0 branch */
} sk_loss_type;

int sk_loss_trivial(sk_loss_function** out);
int sk_loss_huber(double a, sk_loss_function** out);
int sk_loss_cauchy(double a, sk_loss_function** out);
int sk_loss_soft_l_one(double a, sk_loss_function** out);                 /* ceres.i:172 unsupported */
int sk_loss_tukey(double a, sk_loss_function** out);                      /* ceres.i:174 unsupported */
int sk_loss_tolerant(double a, double b, sk_loss_function** out);         /* ceres.i:175: a >= 0, b > 0 */
int sk_loss_destroy(sk_loss_function* loss);
/* rho[0..2] = rho(s), rho'(s), rho''(s) evaluated on the device (LossFunction::Evaluate). */
int sk_loss_evaluate(const sk_loss_function* loss, double s, double rho[3]);

/* ------------------------------------------------------------------------------------------
 * CostFunction — CostFunctor.scala:31-51, SizedCostFunction.scala:6-14,
 * AutodiffCostFunction.scala:68-134.
 * `toAutoDiffCostFunction` on a *registered device functor* yields (functor id, constants);
 * the functor body runs on the GPU on dual numbers held in registers.  An arbitrary JVM
 * closure cannot: unknown ids are rejected with SK_ERR_UNSUPPORTED (no CPU fallback).
 * ---------------------------------------------------------------------------------------- */
typedef enum sk_functor_id {
  /* SimpleBundleAdjuster.scala:79-119  SnavelyReprojectionError(2; 9, 3); consts = (obsX, obsY) */
  SK_FUNCTOR_SNAVELY_REPROJECTION_ERROR = 1,
  /* CurveFitting.scala:92-98  ExponentialResidual(1; 1, 1); consts = (x, y) */
  SK_FUNCTOR_EXPONENTIAL_RESIDUAL = 2,
  /* HelloWorld.scala:11-14  HelloCostFunctor(1; 1): 10 - x; consts = () */
  SK_FUNCTOR_HELLO_WORLD = 3,
  /* Powell.scala:13-51  F1..F4 (1; 1, 1), consts = (): x1 + 10 x2 | sqrt(5) x3 - x4 (as the code computes it, :28 -- its
   * comment and PowellAnalytic.scala:37 say sqrt(5) (x3 - x4)) | (x2 - 2 x3)^2 | sqrt(10) (x1 - x4)^2 */
  SK_FUNCTOR_POWELL_F1 = 4,
  SK_FUNCTOR_POWELL_F2 = 5,
  SK_FUNCTOR_POWELL_F3 = 6,
  SK_FUNCTOR_POWELL_F4 = 7,
  /* PowellAnalytic.scala:25-43  F2a (1; 1, 1): sqrt(5) (x3 - x4), the residual of the analytic-derivative example (a
   * SizedCostFunction there; here its formula on device duals, whose derivatives are the analytic ones).  With F1, F3, F4
   * it is the problem of the Ceres tutorial, whose published log pins the solver (tests/golden). */
  SK_FUNCTOR_POWELL_ANALYTIC_F2 = 8,
  /* The three functors of AutodiffCostFuntionSpec.scala, registered so the reference's own
   * golden vectors can be replayed against the device Jet machinery. */
  SK_FUNCTOR_TEST_BILINEAR_SCALAR = 100,  /* :14-26   (1; 2, 2)   consts = (a)  */
  SK_FUNCTOR_TEST_BILINEAR_VECTOR3 = 101, /* :55-69   (3; 2, 2)   consts = (a)  */
  SK_FUNCTOR_TEST_SUM10 = 102             /* :110-119 (1; 1 x 10) consts = ()   */
} sk_functor_id;

#define SK_MAX_PARAMETER_BLOCKS 10
#define SK_MAX_CONSTS 4

/* kNumResiduals, N* (CostFunctor.scala:31) of a registered functor. */
int sk_functor_info(int functor_id, int* num_residuals, int* num_parameter_blocks,
                    int block_sizes[SK_MAX_PARAMETER_BLOCKS], int* num_consts);

/* A user-defined functor, handed over as CUDA source and compiled at run time (NVRTC, sm_100a) over the device Jet<N>: what
 * `class MyFunctor extends CostFunctor(kNumResiduals, N0, N1, ...) { def apply[T: Field : Trig : NRoot : Order : ClassTag](x: Array[T]*): Array[T] }`
 * followed by `.toAutoDiffCostFunction` is on the JVM (CostFunctor.scala:31-51; evaluated there by AutodiffCostFunction.scala:74-134
 * with one JNI up-call per residual block).  `cuda_source` must define
 *     template <class T> __device__ bool NAME(const double* consts, T const* const* x, T* residuals);
 * with x[k] = parameter block k (block_sizes[k] values) and the return value = the functor's success flag (false = the empty-array
 * convention of CostFunctor.scala:15-26).  T is double or sk::Jet<N0 + N1 + ...>: +, -, *, /, comparisons with doubles, sqrt, exp,
 * log, sin, cos are overloaded (jet.cuh, the operator set spire gives a Jet); `consts` are the per-residual-block constants given to
 * sk_cost_function_create (the constructor arguments of the Scala functor).
 * Returns a functor id (>= 1000) usable wherever a built-in id is: sk_functor_info, sk_cost_function_create / _evaluate,
 * sk_problem_add_residual_block(s) with the dense back end (DENSE_QR and the normal-Cholesky types).  A functor of the
 * bundle-adjustment shape -- 2 residuals, parameter blocks of 9 and 3, 2 constants (the observation), i.e. another camera model in
 * place of SnavelyReprojectionError (SimpleBundleAdjuster.scala:79-119) -- is ALSO compiled into the tile evaluation kernel of the
 * Schur solvers and is accepted by DENSE_SCHUR / SPARSE_SCHUR / ITERATIVE_SCHUR (one functor per problem); sk::angle_axis_rotate_point
 * (Rotation.angleAxisRotatePoint, Rotation.scala:449-522) is available to the source.  Compiling needs no device;
 * the module is loaded on first use.  SK_ERR_INVALID_ARGUMENT with the compiler log in sk_last_error() when the source does not
 * compile; SK_ERR_UNSUPPORTED when libnvrtc is not available.  Limits: <= 16 residuals, <= 10 blocks, <= 32 parameters in total. */
int sk_functor_register_source(const char* name, const char* cuda_source, int num_residuals, int num_parameter_blocks,
                               const int* block_sizes, int num_consts, int* out_functor_id);

typedef struct sk_cost_function sk_cost_function;
int sk_cost_function_create(int functor_id, const double* consts, int num_consts,
                            sk_cost_function** out);
int sk_cost_function_destroy(sk_cost_function* f);
int sk_cost_function_num_residuals(const sk_cost_function* f);

/* bool CostFunction::Evaluate(double const* const* parameters, double* residuals,
 *                             double** jacobians)   — the director up-call of ceres.i:48,
 * implemented by AutodiffCostFunction.scala:74-134.  Same contract, evaluated on the device:
 *   jacobians == NULL            -> residuals only                    (:80)
 *   jacobians[i].array == NULL   -> block i skipped                   (:118)
 *   jacobian block i is row-major kNumResiduals x N_i                 (:121-127)
 *   *ok = 0 when the functor reports failure (empty array convention, CostFunctor.scala:15-26)
 */
int sk_cost_function_evaluate(const sk_cost_function* f, const sk_double_pointer* parameters,
                              sk_double_pointer residuals, const sk_double_pointer* jacobians,
                              int* ok);
/* Host-pointer convenience with the exact Ceres signature; copies in, launches, copies out. */
int sk_cost_function_evaluate_host(const sk_cost_function* f, double const* const* parameters,
                                   double* residuals, double** jacobians, int* ok);

/* ------------------------------------------------------------------------------------------
 * Problem — Problem.scala:16-33 (CeresProblem, ceres.i:73)
 * Ownership is always DO_NOT_TAKE_OWNERSHIP (Problem.scala:10-13): cost functions, losses and
 * arrays must outlive the problem and are destroyed by their owner.
 * ---------------------------------------------------------------------------------------- */
typedef struct sk_problem sk_problem;
typedef int64_t sk_residual_block_id; /* package.scala:13 ResidualBlockId */

int sk_problem_create(sk_problem** out);
int sk_problem_destroy(sk_problem* p);

/* Problem.addResidualBlock(cost, loss, x: DoublePointer*) — Problem.scala:20-27.
 * loss == NULL means "no loss" exactly as in Ceres (treated as trivial). */
int sk_problem_add_residual_block(sk_problem* p, const sk_cost_function* cost,
                                  const sk_loss_function* loss, const sk_double_pointer* blocks,
                                  int num_blocks, sk_residual_block_id* id);

/* Bulk form of the O(n_obs) loop of SimpleBundleAdjuster.scala:139-145: n residual blocks of one
 * functor at once.
 *   consts        n x num_consts, row-major (host memory)
 *   block_offsets n x num_parameter_blocks, row-major: offset of each parameter block inside
 *                 `array` (host memory)
 * Returns the id of the first block; ids are consecutive. */
int sk_problem_add_residual_blocks(sk_problem* p, int functor_id, int64_t n, const double* consts,
                                   const sk_loss_function* loss, sk_double_array* array,
                                   const int64_t* block_offsets, sk_residual_block_id* first_id);

/* Problem::AddParameterBlock(values, size) in bulk (ceres/problem.h; the reference exposes ceres::Problem wholesale,
 * ceres.i:73): declares n parameter blocks of `block_size` doubles at `offsets` inside `array` without attaching a
 * residual block.  As in Ceres, a declared block that no residual block uses does not take part in a solve -- except in
 * the rank-local multi-GPU mode (sk_solver_options.residual_blocks_are_local), where every rank declares ALL cameras
 * so that the replicated camera table is the same on every rank whichever observations it holds. */
int sk_problem_add_parameter_blocks(sk_problem* p, sk_double_array* array, int64_t n,
                                    const int64_t* offsets, int32_t block_size);

int64_t sk_problem_num_residual_blocks(const sk_problem* p);
int64_t sk_problem_num_residuals(const sk_problem* p);
int64_t sk_problem_num_parameter_blocks(const sk_problem* p);
int64_t sk_problem_num_parameters(const sk_problem* p);

/* ------------------------------------------------------------------------------------------
 * Solver.Options / Solver.Summary — SWIG-generated from ceres/solver.h (ceres.i:151); the
 * setters the reference exercises are setLinearSolverType, setMaxNumIterations,
 * setMinimizerProgressToStdout, setMinimizerType (SimpleBundleAdjuster.scala:147-149,
 * CurveFitting.scala:119-122, Powell.scala:77-81).  Enumerator values follow ceres/types.h
 * (ceres.i:137) so the Scala enum ordinals carry over unchanged.
 * ---------------------------------------------------------------------------------------- */
typedef enum sk_linear_solver_type {
  SK_DENSE_NORMAL_CHOLESKY = 0, /* unsupported */
  SK_DENSE_QR = 1,
  SK_SPARSE_NORMAL_CHOLESKY = 2, /* unsupported */
  SK_DENSE_SCHUR = 3,
  SK_SPARSE_SCHUR = 4,
  SK_ITERATIVE_SCHUR = 5,
  SK_CGNR = 6 /* unsupported */
} sk_linear_solver_type;

typedef enum sk_preconditioner_type {
  SK_IDENTITY = 0,
  SK_JACOBI = 1,
  SK_SCHUR_JACOBI = 2,
  SK_CLUSTER_JACOBI = 3,       /* unsupported */
  SK_CLUSTER_TRIDIAGONAL = 4   /* unsupported */
} sk_preconditioner_type;

typedef enum sk_minimizer_type { SK_LINE_SEARCH = 0 /* unsupported */, SK_TRUST_REGION = 1 } sk_minimizer_type;
typedef enum sk_trust_region_strategy_type { SK_LEVENBERG_MARQUARDT = 0, SK_DOGLEG = 1 /* unsupported */ } sk_trust_region_strategy_type;

typedef enum sk_termination_type {
  SK_CONVERGENCE = 0,
  SK_NO_CONVERGENCE = 1,
  SK_FAILURE = 2,
  SK_USER_SUCCESS = 3,
  SK_USER_FAILURE = 4
} sk_termination_type;

typedef struct sk_comm sk_comm;

typedef struct sk_solver_options {
  int32_t minimizer_type;               /* TRUST_REGION */
  int32_t trust_region_strategy_type;   /* LEVENBERG_MARQUARDT */
  int32_t linear_solver_type;           /* SPARSE_NORMAL_CHOLESKY (Ceres default): runs on the dense QR back end, like DENSE_NORMAL_CHOLESKY */
  int32_t preconditioner_type;          /* JACOBI */
  int32_t max_num_iterations;           /* 50 */
  int32_t max_num_consecutive_invalid_steps; /* 5 */
  int32_t min_linear_solver_iterations; /* 0 */
  int32_t max_linear_solver_iterations; /* 500 */
  int32_t jacobi_scaling;               /* 1 */
  int32_t minimizer_progress_to_stdout; /* 0 */
  int32_t num_threads;                  /* 1; accepted and ignored on the device path */
  int32_t profile_kernels;              /* 0: no events; 1: CUDA events around every kernel family (perturbs a multi-GPU
                                           solve by ~10 %); 2: events around the implicit-Schur product only */
  double initial_trust_region_radius;   /* 1e4 */
  double max_trust_region_radius;       /* 1e16 */
  double min_trust_region_radius;       /* 1e-32 */
  double min_relative_decrease;         /* 1e-3 */
  double min_lm_diagonal;               /* 1e-6 */
  double max_lm_diagonal;               /* 1e32 */
  double function_tolerance;            /* 1e-6 */
  double gradient_tolerance;            /* 1e-10 */
  double parameter_tolerance;           /* 1e-8 */
  double eta;                           /* 1e-1 */
  double max_solver_time_in_seconds;    /* 1e9 */
  sk_comm* comm;                        /* NULL = single GPU; else the point-partitioned multi-GPU path */
  int32_t residual_blocks_are_local;    /* multi-GPU bundle adjustment only.  0 (default): every rank passes the WHOLE
                                           problem and the solver keeps this rank's point range (sk_partition_points);
                                           the solution of every point is published to every rank.  1: every rank passes
                                           only the residual blocks of ITS OWN points (no point on two ranks) and declares
                                           all cameras with sk_problem_add_parameter_blocks; ingestion then costs
                                           O(local observations) and a rank's array receives all cameras and its own
                                           points. */
  int32_t reserved_;
} sk_solver_options;

/* Writes the Ceres 1.x defaults listed above. */
void sk_solver_options_init(sk_solver_options* o);

/* One row of Solver::Summary::iterations (ceres/iteration_callback.h IterationSummary). */
typedef struct sk_iteration_summary {
  int32_t iteration;
  int32_t step_is_valid;
  int32_t step_is_nonmonotonic;
  int32_t step_is_successful;
  int32_t linear_solver_iterations;
  int32_t reserved_;
  double cost;
  double cost_change;
  double gradient_max_norm;
  double gradient_norm;
  double step_norm;
  double relative_decrease;
  double trust_region_radius;
  double eta;
  double iteration_time_in_seconds;
  double cumulative_time_in_seconds;
} sk_iteration_summary;

/* Kernel families timed when profile_kernels = 1 (device time from CUDA events). */
typedef enum sk_kernel_family {
  SK_KF_EVALUATE_JACOBIAN = 0, /* residual + Jacobian evaluation (+ gradient / column norms)   */
  SK_KF_EVALUATE_COST = 1,     /* residual-only evaluation of the candidate point              */
  SK_KF_SCHUR_SETUP = 2,       /* E^T E inverses, reduced rhs, SchurJacobi blocks / explicit S */
  SK_KF_SCHUR_MATVEC = 3,      /* implicit-Schur product: the PCG inner kernel                 */
  SK_KF_PCG_VECTOR = 4,        /* fused PCG vector updates, dots and preconditioner apply       */
  SK_KF_BACK_SUBSTITUTE = 5,   /* point back-substitution + model cost change                  */
  SK_KF_DENSE = 6,             /* dense QR / Cholesky                                          */
  SK_KF_LM = 7,                /* LM diagonal, step acceptance, radius update, small vector ops */
  SK_KF_COMM = 8,              /* NCCL allreduce                                               */
  SK_KF_PCG_SOLVE = 9,         /* fused PCG solve: one persistent kernel per linear solve (products + vector phases).
                                  When it runs, kernel_launches[SK_KF_SCHUR_MATVEC] counts the products executed INSIDE it and
                                  (profile_kernels) kernel_ms[SK_KF_SCHUR_MATVEC] / [SK_KF_PCG_VECTOR] hold the time its first CTA
                                  spent in the product / vector phases (device clock), kernel_ms[SK_KF_PCG_SOLVE] the CUDA-event
                                  time of the launches */
  SK_KF_COUNT = 10
} sk_kernel_family;

typedef struct sk_solver_summary_data {
  int32_t termination_type;
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  int32_t num_iterations;        /* rows in `iterations` */
  int32_t linear_solver_type_used;
  int32_t preconditioner_type_used;
  int32_t num_gpus;
  int32_t reserved_;
  double initial_cost;
  double final_cost;
  double fixed_cost;
  int64_t num_parameter_blocks;
  int64_t num_parameters;
  int64_t num_residual_blocks;
  int64_t num_residuals;
  int64_t num_residual_evaluations;  /* calls of the residual-only evaluator (whole problem) */
  int64_t num_jacobian_evaluations;
  int64_t num_linear_solves;
  int64_t total_linear_solver_iterations;
  int64_t num_kernel_launches;       /* kernels of this library launched by the solve */
  double total_time_in_seconds;      /* host wall clock */
  double preprocessor_time_in_seconds;
  double minimizer_time_in_seconds;
  double minimizer_device_time_in_seconds; /* CUDA-event time of the minimizer on the solver's stream */
  double kernel_ms[SK_KF_COUNT];     /* profile_kernels only, else 0 */
  int64_t kernel_launches[SK_KF_COUNT];
} sk_solver_summary_data;

typedef struct sk_solver_summary sk_solver_summary;
int sk_solver_summary_create(sk_solver_summary** out);
int sk_solver_summary_destroy(sk_solver_summary* s);
int sk_solver_summary_get(const sk_solver_summary* s, sk_solver_summary_data* out);
/* Copies up to `capacity` rows; returns the number of rows available through *count. */
int sk_solver_summary_iterations(const sk_solver_summary* s, sk_iteration_summary* out,
                                 int32_t capacity, int32_t* count);
const char* sk_solver_summary_message(const sk_solver_summary* s);
/* Summary.briefReport() / fullReport() — CurveFitting.scala:131, SimpleBundleAdjuster.scala:154.
 * The returned string lives until the summary is destroyed or reused. */
const char* sk_solver_summary_brief_report(sk_solver_summary* s);
const char* sk_solver_summary_full_report(sk_solver_summary* s);
int sk_solver_summary_is_solution_usable(const sk_solver_summary* s);

/* ceres.solve(options, problem, summary) — SimpleBundleAdjuster.scala:152, CurveFitting.scala:127.
 * Runs trust-region Levenberg–Marquardt entirely on the device; on return the parameter arrays
 * hold the solution (when usable), still in device memory. */
int sk_solve(const sk_solver_options* options, sk_problem* problem, sk_solver_summary* summary);

/* ceres.solve split in its two halves so that a caller can keep the preprocessed problem (Ceres'
 * Preprocessor: program, Schur ordering, and here the tiled device layout) resident in HBM and run
 * the minimizer more than once:  sk_solve == create + minimize + destroy.
 *   max_num_iterations_override < 0 keeps options->max_num_iterations; otherwise it must not exceed it.
 * Every sk_solver_minimize call starts from the CURRENT contents of the parameter arrays. */
typedef struct sk_solver sk_solver;
int sk_solver_create(const sk_solver_options* options, sk_problem* problem, sk_solver** out);
int sk_solver_minimize(sk_solver* solver, int32_t max_num_iterations_override, sk_solver_summary* summary);
int sk_solver_destroy(sk_solver* solver);
/* Measurement aid (no reference counterpart): mean device time, in milliseconds, of `reps` back-to-back launches of the
 * implicit Schur-complement product (ImplicitSchurComplement::RightMultiply) on the linearisation the last
 * sk_solver_minimize call left behind, timed with CUDA events on the solver's stream.  ITERATIVE_SCHUR solvers only, after
 * at least one LM iteration.  bench.py and tools/ use it to time the dominant kernel in isolation. */
int sk_solver_time_schur_product(sk_solver* solver, int32_t reps, double* out_ms_per_launch);

/* ------------------------------------------------------------------------------------------
 * Batched independent small problems (BASELINE.json configs[3]): n_problems CurveFitting-shaped
 * problems (CurveFitting.scala:92-122: ExponentialResidual, two scalar blocks m and c, DENSE_QR,
 * trust-region LM), one thread per problem, one launch per LM iteration.
 *   x, y    device arrays laid out [n_obs][n_problems]  (observation-major SoA)
 *   mc      device array [2][n_problems]: m then c, initial values in, solution out
 *   out_*   optional host arrays of n_problems entries (may be NULL)
 * ---------------------------------------------------------------------------------------- */
int sk_curve_fit_batch_solve(const sk_solver_options* options, int64_t n_problems, int32_t n_obs,
                             const sk_double_array* x, const sk_double_array* y,
                             sk_double_array* mc, double* out_initial_cost, double* out_final_cost,
                             int32_t* out_num_iterations, int32_t* out_termination_type,
                             sk_solver_summary* summary);

/* ------------------------------------------------------------------------------------------
 * BalProblem.fromFile — SimpleBundleAdjuster.scala:37-77.  Parses the BAL text layout
 * (`<ncam> <npt> <nobs>`, nobs lines `cam pt x y`, then 9*ncam + 3*npt values) and leaves the
 * parameter vector (cameras first, :28-33) in a device DoubleArray.
 * ---------------------------------------------------------------------------------------- */
typedef struct sk_bal_problem sk_bal_problem;
int sk_bal_problem_from_file(const char* path, sk_bal_problem** out);
int sk_bal_problem_destroy(sk_bal_problem* b);
int32_t sk_bal_problem_num_cameras(const sk_bal_problem* b);
int32_t sk_bal_problem_num_points(const sk_bal_problem* b);
int32_t sk_bal_problem_num_observations(const sk_bal_problem* b);
sk_double_array* sk_bal_problem_parameters(sk_bal_problem* b);       /* :26  parameters */
const int32_t* sk_bal_problem_camera_index(const sk_bal_problem* b); /* :24  host memory */
const int32_t* sk_bal_problem_point_index(const sk_bal_problem* b);  /* :23  host memory */
const double* sk_bal_problem_observations(const sk_bal_problem* b);  /* :25  host memory, 2*nobs */
/* The loop of SimpleBundleAdjuster.scala:139-145 as one call. */
int sk_bal_problem_build(sk_bal_problem* b, const sk_loss_function* loss, sk_problem* problem);
/* mutableCameraForObservation(i) / mutablePointForObservation(i) (SimpleBundleAdjuster.scala:31-33) for observations
 * [0, n_obs) at once, as offsets into the BalProblem parameter array (cameras first, :28-29): out[2i] = 9 * camera_index[i],
 * out[2i + 1] = 9 * num_cameras + 3 * point_index[i].  Host-only helper: the interleaved array is what
 * sk_problem_add_residual_blocks takes.  Fails on an index outside [0, num_cameras) / [0, num_points). */
int sk_bal_block_offsets(int64_t n_obs, const int32_t* camera_index, const int32_t* point_index,
                         int32_t num_cameras, int32_t num_points, int64_t* out_offsets);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU (no counterpart in the reference, which is single-threaded): one process per GPU,
 * bundle-adjustment problems partitioned by point, cameras replicated, NCCL allreduce of the
 * camera-block quantities.  The unique id is created on rank 0 and shipped to the other ranks
 * by the host program (torch.distributed, MPI, a file ...).
 * ---------------------------------------------------------------------------------------- */
#define SK_COMM_UNIQUE_ID_BYTES 128
int sk_comm_get_unique_id(char id[SK_COMM_UNIQUE_ID_BYTES]);
int sk_comm_create(const char id[SK_COMM_UNIQUE_ID_BYTES], int rank, int world_size, sk_comm** out);
int sk_comm_destroy(sk_comm* c);
int sk_comm_rank(const sk_comm* c);
int sk_comm_world_size(const sk_comm* c);

/* Host-only helper (no GPU needed): the contiguous point ranges of the partition.
 * point_ptr has n_points + 1 entries (CSR offsets of the point-sorted observation list);
 * out_begin has world_size + 1 entries.  Balanced by observation count. */
int sk_partition_points(int64_t n_points, const int64_t* point_ptr, int world_size,
                        int64_t* out_begin);

/* Device memory freed by destroyed solvers / arrays is cached for the next allocation of the same size (cudaMalloc and
 * cudaFree of GB-sized buffers cost 0.1-0.4 s per solver otherwise).  sk_release_cached_memory returns it to the driver --
 * the counterpart of what a JVM host does with a native arena (no reference equivalent: libceres uses malloc). */
int sk_release_cached_memory(void);
int64_t sk_cached_memory_bytes(void);

#ifdef __cplusplus
}
#endif
#endif /* SKERES_H_ */
