"""ctypes loader of libskeres.so (the C ABI of include/skeres.h).

The library is built in-tree by `make -C skeres_b200/csrc` (see `__graft_entry__.build()`).  There is
no fallback of any kind: if the shared object is missing the import fails loudly, and every numeric
entry point fails with SK_ERR_CUDA when no B200 is visible.
"""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# SKERES_LIB: development only -- an alternative build of the same library (kernel A/B variants under gpurun_variants/)
LIB_PATH = os.environ.get("SKERES_LIB") or os.path.join(_HERE, "libskeres.so")


class SkeresError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"libskeres error {status}: {message}")
        self.status = status


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C skeres_b200/csrc` "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    P = C.POINTER
    sig = {
        "sk_last_error": (C.c_char_p, []),
        "sk_abi_version": (i32, []),
        "sk_device_count": (i32, []),
        "sk_set_device": (i32, [i32]),
        "sk_init_google_logging": (None, [C.c_char_p]),
        "sk_double_array_create": (i32, [i64, P(vp)]),
        "sk_double_array_destroy": (i32, [vp]),
        "sk_double_array_size": (i64, [vp]),
        "sk_double_array_upload": (i32, [vp, i64, vp, i64]),
        "sk_double_array_download": (i32, [vp, i64, vp, i64]),
        "sk_double_array_get": (i32, [vp, i64, P(dbl)]),
        "sk_double_array_set": (i32, [vp, i64, dbl]),
        "sk_double_array_device_ptr": (vp, [vp]),
        "sk_double_array_copy": (i32, [vp, i64, vp, i64, i64]),
        "sk_loss_trivial": (i32, [P(vp)]),
        "sk_loss_huber": (i32, [dbl, P(vp)]),
        "sk_loss_cauchy": (i32, [dbl, P(vp)]),
        "sk_loss_soft_l_one": (i32, [dbl, P(vp)]),
        "sk_loss_tukey": (i32, [dbl, P(vp)]),
        "sk_loss_tolerant": (i32, [dbl, dbl, P(vp)]),
        "sk_loss_destroy": (i32, [vp]),
        "sk_loss_evaluate": (i32, [vp, dbl, P(dbl)]),
        "sk_functor_info": (i32, [i32, P(i32), P(i32), P(i32), P(i32)]),
        "sk_functor_register_source": (i32, [C.c_char_p, C.c_char_p, i32, i32, P(i32), i32, P(i32)]),
        "sk_cost_function_create": (i32, [i32, vp, i32, P(vp)]),
        "sk_cost_function_destroy": (i32, [vp]),
        "sk_cost_function_num_residuals": (i32, [vp]),
        "sk_cost_function_evaluate": (i32, [vp, vp, _abi.DoublePointer, vp, P(i32)]),
        "sk_cost_function_evaluate_host": (i32, [vp, vp, vp, vp, P(i32)]),
        "sk_problem_create": (i32, [P(vp)]),
        "sk_problem_destroy": (i32, [vp]),
        "sk_problem_add_residual_block": (i32, [vp, vp, vp, vp, i32, P(i64)]),
        "sk_problem_add_residual_blocks": (i32, [vp, i32, i64, vp, vp, vp, vp, P(i64)]),
        "sk_problem_add_parameter_blocks": (i32, [vp, vp, i64, vp, i32]),
        "sk_bal_block_offsets": (i32, [i64, vp, vp, i32, i32, vp]),
        "sk_problem_num_residual_blocks": (i64, [vp]),
        "sk_problem_num_residuals": (i64, [vp]),
        "sk_problem_num_parameter_blocks": (i64, [vp]),
        "sk_problem_num_parameters": (i64, [vp]),
        "sk_solver_options_init": (None, [P(_abi.SolverOptions)]),
        "sk_solver_summary_create": (i32, [P(vp)]),
        "sk_solver_summary_destroy": (i32, [vp]),
        "sk_solver_summary_get": (i32, [vp, P(_abi.SolverSummaryData)]),
        "sk_solver_summary_iterations": (i32, [vp, vp, i32, P(i32)]),
        "sk_solver_summary_message": (C.c_char_p, [vp]),
        "sk_solver_summary_brief_report": (C.c_char_p, [vp]),
        "sk_solver_summary_full_report": (C.c_char_p, [vp]),
        "sk_solver_summary_is_solution_usable": (i32, [vp]),
        "sk_solve": (i32, [P(_abi.SolverOptions), vp, vp]),
        "sk_solver_create": (i32, [P(_abi.SolverOptions), vp, P(vp)]),
        "sk_solver_minimize": (i32, [vp, i32, vp]),
        "sk_solver_time_schur_product": (i32, [vp, i32, P(dbl)]),
        "sk_solver_destroy": (i32, [vp]),
        "sk_curve_fit_batch_solve": (i32, [P(_abi.SolverOptions), i64, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
        "sk_bal_problem_from_file": (i32, [C.c_char_p, P(vp)]),
        "sk_bal_problem_destroy": (i32, [vp]),
        "sk_bal_problem_num_cameras": (i32, [vp]),
        "sk_bal_problem_num_points": (i32, [vp]),
        "sk_bal_problem_num_observations": (i32, [vp]),
        "sk_bal_problem_parameters": (vp, [vp]),
        "sk_bal_problem_camera_index": (vp, [vp]),
        "sk_bal_problem_point_index": (vp, [vp]),
        "sk_bal_problem_observations": (vp, [vp]),
        "sk_bal_problem_build": (i32, [vp, vp, vp]),
        "sk_comm_get_unique_id": (i32, [vp]),
        "sk_comm_create": (i32, [vp, i32, i32, P(vp)]),
        "sk_comm_destroy": (i32, [vp]),
        "sk_comm_rank": (i32, [vp]),
        "sk_comm_world_size": (i32, [vp]),
        "sk_partition_points": (i32, [i64, vp, i32, vp]),
        "sk_release_cached_memory": (i32, []),
        "sk_cached_memory_bytes": (i64, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)          # AttributeError here == the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return L, sorted(sig)


lib, DECLARED_SYMBOLS = _load()


def check(status):
    if status != _abi.OK:
        raise SkeresError(status, lib.sk_last_error().decode())
