"""ctypes binding of oracle/liboracle.so — the CPU parity oracle.  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing under skeres_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from skeres_b200 import _abi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "liboracle.so")
_lib = None


def build():
    subprocess.run(["make", "-C", os.path.join(_ROOT, "oracle")], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.oracle_problem_create.restype = C.c_void_p
        L.oracle_problem_create.argtypes = [C.c_void_p, C.c_int64]
        L.oracle_problem_destroy.argtypes = [C.c_void_p]
        L.oracle_problem_add_residual_blocks.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int,
                                                         C.c_double, C.c_void_p]
        L.oracle_problem_evaluate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_char_p, C.c_int]
        L.oracle_evaluate.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_loss_evaluate.argtypes = [C.c_int, C.c_double, C.c_double, C.c_void_p]
        L.oracle_loss_evaluate2.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.oracle_problem_add_residual_blocks2.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int,
                                                          C.c_double, C.c_double, C.c_void_p]
        L.oracle_angle_axis_rotate_point.argtypes = [C.c_void_p] * 3
        L.oracle_angle_axis_to_rotation_matrix.argtypes = [C.c_void_p] * 2
        L.oracle_functor_info.argtypes = [C.c_int] + [C.c_void_p] * 4
        L.oracle_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def functor_info(fid):
    nres, nblk, nc = C.c_int(), C.c_int(), C.c_int()
    sizes = (C.c_int * _abi.MAX_PARAMETER_BLOCKS)()
    ok = lib().oracle_functor_info(fid, C.byref(nres), C.byref(nblk), sizes, C.byref(nc))
    assert ok, f"unknown functor {fid}"
    return nres.value, [sizes[i] for i in range(nblk.value)], nc.value


def evaluate(fid, consts, params, want_jacobians=True, skip_blocks=()):
    """CostFunction::Evaluate. Returns (ok, residuals, [jacobian blocks or None])."""
    nres, sizes, _ = functor_info(fid)
    consts = np.ascontiguousarray(consts, dtype=np.float64)
    blocks = [np.ascontiguousarray(p, dtype=np.float64) for p in params]
    pp = (C.c_void_p * len(blocks))(*[b.ctypes.data for b in blocks])
    res = np.zeros(nres)
    jacs = [None if (i in skip_blocks) else np.zeros((nres, sizes[i])) for i in range(len(sizes))]
    if want_jacobians:
        jp = (C.c_void_p * len(blocks))(*[(j.ctypes.data if j is not None else None) for j in jacs])
    else:
        jp, jacs = None, None
    ok = lib().oracle_evaluate(fid, _p(consts), pp, _p(res), jp)
    return bool(ok), res, jacs


def rotate_point(aa, pt):
    aa = np.ascontiguousarray(aa, dtype=np.float64); pt = np.ascontiguousarray(pt, dtype=np.float64)
    out = np.zeros(3)
    lib().oracle_angle_axis_rotate_point(_p(aa), _p(pt), _p(out))
    return out


def rotation_matrix(aa):
    """Column-major 3x3 as in Rotation.scala:206-209; returned as a (3,3) array R[i, j]."""
    aa = np.ascontiguousarray(aa, dtype=np.float64)
    R = np.zeros(9)
    lib().oracle_angle_axis_to_rotation_matrix(_p(aa), _p(R))
    return R.reshape(3, 3).T.copy()


def loss(kind, a, s, b=0.0):
    rho = np.zeros(3)
    lib().oracle_loss_evaluate2(kind, float(a), float(b), float(s), _p(rho))
    return rho


class Summary:
    def __init__(self, data, iterations, message):
        self.data, self.iterations, self.message = data, iterations, message

    def __getattr__(self, k):
        return getattr(self.data, k)


class OracleProblem:
    """One parameter vector (host numpy array, updated in place by solve) + residual blocks."""

    def __init__(self, params):
        self.params = np.ascontiguousarray(params, dtype=np.float64).copy()
        self._h = lib().oracle_problem_create(_p(self.params), self.params.size)
        self.num_residual_blocks = 0
        self.num_residuals = 0
        self._jac_size = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_problem_destroy(self._h)
            self._h = None

    def add_residual_blocks(self, fid, consts, block_offsets, loss_type=_abi.LOSS_TRIVIAL, loss_a=0.0, loss_b=0.0):
        nres, sizes, nc = functor_info(fid)
        off = np.ascontiguousarray(block_offsets, dtype=np.int64).reshape(-1, len(sizes))
        n = off.shape[0]
        consts = np.ascontiguousarray(consts, dtype=np.float64).reshape(n, nc) if nc else np.zeros((n, 0))
        ok = lib().oracle_problem_add_residual_blocks2(self._h, fid, n, _p(consts), loss_type, float(loss_a), float(loss_b), _p(off))
        assert ok, "oracle_problem_add_residual_blocks failed"
        self.num_residual_blocks += n
        self.num_residuals += n * nres
        self._jac_size += n * nres * sum(sizes)

    def evaluate(self):
        cost = C.c_double()
        r = np.zeros(self.num_residuals); g = np.zeros(self.params.size); J = np.zeros(self._jac_size)
        ok = lib().oracle_problem_evaluate(self._h, C.byref(cost), _p(r), _p(g), _p(J))
        assert ok, "oracle evaluate failed"
        return cost.value, r, g, J

    def solve(self, options, threads=None):
        if threads is not None:
            lib().oracle_set_num_threads(int(threads))
        data = _abi.SolverSummaryData()
        cap = max(int(options.max_num_iterations) + 2, 4)
        its = (_abi.IterationSummary * cap)()
        cnt = C.c_int()
        msg = C.create_string_buffer(512)
        lib().oracle_solve(C.byref(options), self._h, C.byref(data), its, cap, C.byref(cnt), msg, 512)
        return Summary(data, [its[i] for i in range(min(cnt.value, cap))], msg.value.decode())
