"""Host-side mirror of the reference's Scala API over the C ABI of libskeres.so.

The reference's host language is Scala (JVM), whose toolchain is absent from this image, so the
host side above the C ABI is written in Python with the same names, argument meaning and error
behaviour as the reference interface for this path:

    reference (file:line)                                   here
    ------------------------------------------------------  -----------------------------------
    DoubleArray / RichDoubleArray  (ceres.i:95-96,           DoubleArray (+ slice -> DoublePointer)
        RichDoubleArray.scala:14-75)
    CostFunctor / AutoDiffCostFunctor(kNumResiduals, N*)     AutoDiffCostFunctor
        .toAutoDiffCostFunction  (CostFunctor.scala:31-51)       .toAutoDiffCostFunction()
    AutoDiffCostFunction.evaluate                            CostFunction.evaluate(...)
        (AutodiffCostFunction.scala:74-134)
    SnavelyReprojectionError (SimpleBundleAdjuster.scala:79) SnavelyReprojectionError
    ExponentialResidual      (CurveFitting.scala:92)         ExponentialResidual
    PredefinedLossFunctions  (ceres.i:160-184)               PredefinedLossFunctions
    Problem.addResidualBlock (Problem.scala:20-27)           Problem.addResidualBlock(cost, loss, *x)
    Solver.Options / Solver.Summary (ceres.i:151)            Solver.Options / Solver.Summary
    ceres.solve(options, problem, summary)                   ceres.solve(options, problem, summary)
    BalProblem.fromFile (SimpleBundleAdjuster.scala:37-77)   BalProblem.fromFile / fromArrays

All numeric work happens inside libskeres.so on the GPU; this module only marshals handles.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._lib import SkeresError, check, lib

LinearSolverType = type("LinearSolverType", (), {k: getattr(_abi, k) for k in (
    "DENSE_NORMAL_CHOLESKY", "DENSE_QR", "SPARSE_NORMAL_CHOLESKY", "DENSE_SCHUR", "SPARSE_SCHUR", "ITERATIVE_SCHUR", "CGNR")})
PreconditionerType = type("PreconditionerType", (), {k: getattr(_abi, k) for k in (
    "IDENTITY", "JACOBI", "SCHUR_JACOBI", "CLUSTER_JACOBI", "CLUSTER_TRIDIAGONAL")})
MinimizerType = type("MinimizerType", (), {"LINE_SEARCH": _abi.LINE_SEARCH, "TRUST_REGION": _abi.TRUST_REGION})
TerminationType = type("TerminationType", (), {k: getattr(_abi, k) for k in (
    "CONVERGENCE", "NO_CONVERGENCE", "FAILURE", "USER_SUCCESS", "USER_FAILURE")})


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _destroy(name, handle, _lib=lib):
    """Handle destructor usable from __del__.  At interpreter shutdown module globals (including this function's own
    name and `lib`) are set to None, so everything needed is bound as a default argument at definition time; the
    __del__ methods below bind `_destroy` the same way."""
    if _lib is not None and handle:
        getattr(_lib, name)(handle)


class DoublePointer:
    """SWIGTYPE_p_double (package.scala:11): an interior pointer = (array, offset)."""

    def __init__(self, array, offset=0):
        self.array, self.offset = array, int(offset)

    def slice(self, start):                      # RichDoubleArray.slice, RichDoubleArray.scala:52
        return DoublePointer(self.array, self.offset + int(start))

    def get(self, i=0):
        return self.array.get(self.offset + i)

    def set(self, i, value):
        self.array.set(self.offset + i, value)

    def toArray(self, n):                        # RichDoubleArray.toArray, :60-70
        return self.array.toArray(n, self.offset)

    def _c(self):
        return _abi.DoublePointer(self.array._h if self.array is not None else None, self.offset)


class DoubleArray:
    """carrays.i DoubleArray (ceres.i:95-96), resident in device memory."""

    def __init__(self, n):
        h = C.c_void_p()
        check(lib.sk_double_array_create(int(n), C.byref(h)))
        self._h, self.n = h, int(n)

    def __del__(self, _destroy=_destroy):
        if getattr(self, "_h", None):
            _destroy("sk_double_array_destroy", self._h)
            self._h = None

    @classmethod
    def fromArray(cls, values):
        v = np.ascontiguousarray(values, dtype=np.float64).ravel()
        a = cls(v.size)
        a.copyFrom(v)
        return a

    def __len__(self):
        return self.n

    def get(self, i):                            # getitem: one device round trip (cf. one JNI crossing, :20)
        out = C.c_double()
        check(lib.sk_double_array_get(self._h, int(i), C.byref(out)))
        return out.value

    def set(self, i, value):
        check(lib.sk_double_array_set(self._h, int(i), float(value)))

    def copyFrom(self, values, offset=0):        # RichDoubleArray.copyFrom (:30-40), bulk
        v = np.ascontiguousarray(values, dtype=np.float64).ravel()
        check(lib.sk_double_array_upload(self._h, int(offset), _vp(v), v.size))

    def toArray(self, n=None, offset=0):
        n = self.n - offset if n is None else int(n)
        out = np.empty(n, dtype=np.float64)
        check(lib.sk_double_array_download(self._h, int(offset), _vp(out), n))
        return out

    def toPointer(self):
        return DoublePointer(self, 0)

    def slice(self, start):
        return DoublePointer(self, start)

    def device_ptr(self):
        return lib.sk_double_array_device_ptr(self._h)

    def copyFromArray(self, src, n=None, dst_offset=0, src_offset=0):
        """Device-to-device copy from another DoubleArray."""
        n = min(self.n - dst_offset, src.n - src_offset) if n is None else int(n)
        check(lib.sk_double_array_copy(self._h, int(dst_offset), src._h, int(src_offset), n))


class LossFunction:
    def __init__(self, handle, kind, a=0.0, b=0.0):
        self._h, self.kind, self.a, self.b = handle, kind, a, b

    def __del__(self, _destroy=_destroy):
        if getattr(self, "_h", None):
            _destroy("sk_loss_destroy", self._h)
            self._h = None

    def evaluate(self, s):
        """LossFunction::Evaluate: (rho, rho', rho'')."""
        rho = (C.c_double * 3)()
        check(lib.sk_loss_evaluate(self._h, float(s), rho))
        return np.array(rho[:])


class PredefinedLossFunctions:
    """ceres.i:160-184. trivial / huber / cauchy / tolerant run on the device; the others are rejected."""

    @staticmethod
    def _make(fn, kind, *args):
        h = C.c_void_p()
        check(fn(*[float(a) for a in args], C.byref(h)))
        return LossFunction(h, kind, args[0] if args else 0.0, args[1] if len(args) > 1 else 0.0)

    @staticmethod
    def trivialLoss():
        return PredefinedLossFunctions._make(lib.sk_loss_trivial, _abi.LOSS_TRIVIAL)

    @staticmethod
    def huberLoss(a):
        return PredefinedLossFunctions._make(lib.sk_loss_huber, _abi.LOSS_HUBER, a)

    @staticmethod
    def cauchyLoss(a):
        return PredefinedLossFunctions._make(lib.sk_loss_cauchy, _abi.LOSS_CAUCHY, a)

    @staticmethod
    def softLOneLoss(a):
        return PredefinedLossFunctions._make(lib.sk_loss_soft_l_one, -1, a)

    @staticmethod
    def tukeyLoss(a):
        return PredefinedLossFunctions._make(lib.sk_loss_tukey, -1, a)

    @staticmethod
    def tolerantLoss(a, b):
        return PredefinedLossFunctions._make(lib.sk_loss_tolerant, _abi.LOSS_TOLERANT, a, b)


def functor_info(functor_id):
    nres, nblk, nc = C.c_int(), C.c_int(), C.c_int()
    sizes = (C.c_int * _abi.MAX_PARAMETER_BLOCKS)()
    check(lib.sk_functor_info(int(functor_id), C.byref(nres), C.byref(nblk), sizes, C.byref(nc)))
    return nres.value, [sizes[i] for i in range(nblk.value)], nc.value


class CostFunction:
    """AutoDiffCostFunction of a registered device functor (AutodiffCostFunction.scala:68)."""

    def __init__(self, functor_id, consts):
        self.functor_id = int(functor_id)
        self.kNumResiduals, self.N, nc = functor_info(functor_id)
        self.consts = np.ascontiguousarray(consts, dtype=np.float64).ravel()
        h = C.c_void_p()
        check(lib.sk_cost_function_create(self.functor_id, _vp(self.consts), self.consts.size, C.byref(h)))
        self._h = h

    def __del__(self, _destroy=_destroy):
        if getattr(self, "_h", None):
            _destroy("sk_cost_function_destroy", self._h)
            self._h = None

    def evaluate(self, parameters, residuals, jacobians):
        """bool evaluate(parameters: DoublePointerPointer, residuals: DoublePointer, jacobians: DoublePointerPointer)
        (AutodiffCostFunction.scala:74-78).  `parameters` / `jacobians` are sequences of DoublePointer
        (a RichDoubleMatrix row each); jacobians None == NULL (residuals only); an entry None == NULL row."""
        pp = (_abi.DoublePointer * len(self.N))(*[p._c() for p in parameters])
        jp = None
        if jacobians is not None:
            jp = (_abi.DoublePointer * len(self.N))(*[(j._c() if j is not None else _abi.DoublePointer(None, 0)) for j in jacobians])
        ok = C.c_int()
        check(lib.sk_cost_function_evaluate(self._h, pp, residuals._c(), jp, C.byref(ok)))
        return bool(ok.value)

    def evaluate_host(self, parameters, want_jacobians=True, skip_blocks=()):
        """Same contract on host arrays (the raw Ceres ABI): returns (ok, residuals, jacobian blocks)."""
        blocks = [np.ascontiguousarray(p, dtype=np.float64) for p in parameters]
        pp = (C.c_void_p * len(blocks))(*[b.ctypes.data for b in blocks])
        res = np.zeros(self.kNumResiduals)
        jacs = None
        jp = None
        if want_jacobians:
            jacs = [None if i in skip_blocks else np.zeros((self.kNumResiduals, n)) for i, n in enumerate(self.N)]
            jp = (C.c_void_p * len(blocks))(*[(j.ctypes.data if j is not None else None) for j in jacs])
        ok = C.c_int()
        check(lib.sk_cost_function_evaluate_host(self._h, pp, _vp(res), jp, C.byref(ok)))
        return bool(ok.value), res, jacs


class AutoDiffCostFunctor:
    """CostFunctor.scala:40-51.  Subclasses name a REGISTERED device functor; an arbitrary Python/JVM
    closure cannot run on the GPU and is rejected by the library (no CPU fallback)."""
    functor_id = None

    def __init__(self, kNumResiduals, *N):
        assert kNumResiduals > 0, f"Nonpositive number of residuals specified: {kNumResiduals}"
        assert all(n > 0 for n in N), f"Nonpositive parameter block sizes specified: {N}"
        self.kNumResiduals, self.N = kNumResiduals, list(N)

    def consts(self):
        return []

    def toAutoDiffCostFunction(self):
        if self.functor_id is None:
            raise SkeresError(_abi.ERR_UNSUPPORTED, f"{type(self).__name__} is not a registered device functor")
        cf = CostFunction(self.functor_id, self.consts())
        assert cf.kNumResiduals == self.kNumResiduals and cf.N == self.N, "functor shape does not match the device registration"
        return cf


class SourceCostFunctor(AutoDiffCostFunctor):
    """A functor written by the user, as CUDA source over T = double / Jet<N> (sk_functor_register_source): the device
    counterpart of subclassing CostFunctor on the JVM (CostFunctor.scala:31-51).

        F = SourceCostFunctor.define("MyResidual", SRC, 1, [1, 1], num_consts=2)     # compile once (NVRTC)
        cost = F(x_i, y_i).toAutoDiffCostFunction()                                   # per residual block: its constants
    """
    @staticmethod
    def define(name, cuda_source, kNumResiduals, N, num_consts=0):
        sizes = (C.c_int * len(N))(*N)
        fid = C.c_int(0)
        check(lib.sk_functor_register_source(name.encode(), cuda_source.encode(), int(kNumResiduals), len(N), sizes, int(num_consts), C.byref(fid)))

        class _Defined(SourceCostFunctor):
            functor_id = fid.value

            def __init__(self, *consts):
                assert len(consts) == num_consts, f"{name} takes {num_consts} constants"
                super().__init__(kNumResiduals, *N)
                self._consts = [float(c) for c in consts]

            def consts(self):
                return self._consts
        _Defined.__name__ = name
        return _Defined


class SnavelyReprojectionError(AutoDiffCostFunctor):
    """SimpleBundleAdjuster.scala:79-119."""
    functor_id = _abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR

    def __init__(self, observedX, observedY):
        super().__init__(2, 9, 3)
        self.observedX, self.observedY = float(observedX), float(observedY)

    def consts(self):
        return [self.observedX, self.observedY]


class ExponentialResidual(AutoDiffCostFunctor):
    """CurveFitting.scala:92-98."""
    functor_id = _abi.FUNCTOR_EXPONENTIAL_RESIDUAL

    def __init__(self, x, y):
        super().__init__(1, 1, 1)
        self.x, self.y = float(x), float(y)

    def consts(self):
        return [self.x, self.y]


class HelloCostFunctor(AutoDiffCostFunctor):
    """HelloWorld.scala:11-14: 10 - x."""
    functor_id = _abi.FUNCTOR_HELLO_WORLD

    def __init__(self):
        super().__init__(1, 1)

    def consts(self):
        return []


class _PowellFunctor(AutoDiffCostFunctor):
    def __init__(self):
        super().__init__(1, 1, 1)

    def consts(self):
        return []


class Powell:
    """Powell.scala:13-51: the four autodiff functors (F2 as the code computes it, sqrt(5) x3 - x4)."""
    class F1(_PowellFunctor): functor_id = _abi.FUNCTOR_POWELL_F1
    class F2(_PowellFunctor): functor_id = _abi.FUNCTOR_POWELL_F2
    class F3(_PowellFunctor): functor_id = _abi.FUNCTOR_POWELL_F3
    class F4(_PowellFunctor): functor_id = _abi.FUNCTOR_POWELL_F4
    class F2a(_PowellFunctor): functor_id = _abi.FUNCTOR_POWELL_ANALYTIC_F2   # PowellAnalytic.scala:25-43: sqrt(5) (x3 - x4)


class Problem:
    """Problem.scala:16-33."""

    def __init__(self):
        h = C.c_void_p()
        check(lib.sk_problem_create(C.byref(h)))
        self._h = h
        self._keep = []          # Problem.scala:29-32: keep cost/loss/arrays alive

    def __del__(self, _destroy=_destroy):
        if getattr(self, "_h", None):
            _destroy("sk_problem_destroy", self._h)
            self._h = None

    def addResidualBlock(self, cost, loss, *x):
        self._keep.extend([cost, loss, *[p.array for p in x]])
        blocks = (_abi.DoublePointer * len(x))(*[p._c() for p in x])
        rid = C.c_int64()
        check(lib.sk_problem_add_residual_block(self._h, cost._h, loss._h if loss is not None else None, blocks, len(x), C.byref(rid)))
        return rid.value

    def addResidualBlocks(self, functor_id, consts, loss, array, block_offsets):
        """Bulk form of the loop at SimpleBundleAdjuster.scala:139-145."""
        nres, sizes, nc = functor_info(functor_id)
        off = np.ascontiguousarray(block_offsets, dtype=np.int64).reshape(-1, len(sizes))
        n = off.shape[0]
        consts = np.ascontiguousarray(consts, dtype=np.float64).reshape(n, nc) if nc else None
        self._keep.extend([loss, array])
        rid = C.c_int64()
        check(lib.sk_problem_add_residual_blocks(self._h, int(functor_id), n, _vp(consts), loss._h if loss is not None else None,
                                                 array._h, _vp(off), C.byref(rid)))
        return rid.value

    def addParameterBlocks(self, array, offsets, size):
        """Problem::AddParameterBlock(values, size) in bulk: declares blocks without attaching residual blocks."""
        off = np.ascontiguousarray(offsets, dtype=np.int64).ravel()
        self._keep.append(array)
        check(lib.sk_problem_add_parameter_blocks(self._h, array._h, off.size, _vp(off), int(size)))

    def numResidualBlocks(self):
        return lib.sk_problem_num_residual_blocks(self._h)

    def numResiduals(self):
        return lib.sk_problem_num_residuals(self._h)

    def numParameterBlocks(self):
        return lib.sk_problem_num_parameter_blocks(self._h)

    def numParameters(self):
        return lib.sk_problem_num_parameters(self._h)


class Solver:
    class Options:
        """Solver.Options with the SWIG-style setters the reference calls."""

        def __init__(self):
            self._o = _abi.SolverOptions()
            lib.sk_solver_options_init(C.byref(self._o))
            self._comm = None

        def __getattr__(self, k):
            if k.startswith("_"):
                raise AttributeError(k)
            return getattr(self._o, k)

        def __setattr__(self, k, v):
            if k.startswith("_"):
                object.__setattr__(self, k, v)
            elif k == "comm":
                self._comm = v
                self._o.comm = v._h if v is not None else None
            else:
                if not hasattr(self._o, k):
                    raise AttributeError(k)
                setattr(self._o, k, v)

        def setLinearSolverType(self, t): self._o.linear_solver_type = int(t)
        def setPreconditionerType(self, t): self._o.preconditioner_type = int(t)
        def setMaxNumIterations(self, n): self._o.max_num_iterations = int(n)
        def setMinimizerProgressToStdout(self, b): self._o.minimizer_progress_to_stdout = int(bool(b))
        def setMinimizerType(self, t): self._o.minimizer_type = int(t)
        def setNumThreads(self, n): self._o.num_threads = int(n)

    class Summary:
        def __init__(self):
            h = C.c_void_p()
            check(lib.sk_solver_summary_create(C.byref(h)))
            self._h = h

        def __del__(self, _destroy=_destroy):
            if getattr(self, "_h", None):
                _destroy("sk_solver_summary_destroy", self._h)
                self._h = None

        @property
        def data(self):
            d = _abi.SolverSummaryData()
            check(lib.sk_solver_summary_get(self._h, C.byref(d)))
            return d

        def __getattr__(self, k):
            if k.startswith("_"):
                raise AttributeError(k)
            return getattr(self.data, k)

        @property
        def iterations(self):
            cnt = C.c_int32()
            check(lib.sk_solver_summary_iterations(self._h, None, 0, C.byref(cnt)))
            rows = (_abi.IterationSummary * max(cnt.value, 1))()
            check(lib.sk_solver_summary_iterations(self._h, rows, cnt.value, C.byref(cnt)))
            return [rows[i] for i in range(cnt.value)]

        @property
        def message(self):
            return lib.sk_solver_summary_message(self._h).decode()

        def briefReport(self):
            return lib.sk_solver_summary_brief_report(self._h).decode()

        def fullReport(self):
            return lib.sk_solver_summary_full_report(self._h).decode()

        def isSolutionUsable(self):
            return bool(lib.sk_solver_summary_is_solution_usable(self._h))

        def kernel_times(self):
            d = self.data
            return {_abi.KF_NAMES[i]: (d.kernel_ms[i], d.kernel_launches[i]) for i in range(_abi.KF_COUNT)}


class ceres:
    """The `ceres` SWIG module object (ceres.solve, ceres.initGoogleLogging)."""

    @staticmethod
    def initGoogleLogging(name):
        lib.sk_init_google_logging(name.encode())

    @staticmethod
    def solve(options, problem, summary):
        check(lib.sk_solve(C.byref(options._o), problem._h, summary._h))


class PreparedSolver:
    """sk_solver: the preprocessed problem kept resident in HBM; minimize() can be called repeatedly
    and always starts from the current contents of the parameter arrays."""

    def __init__(self, options, problem):
        h = C.c_void_p()
        check(lib.sk_solver_create(C.byref(options._o), problem._h, C.byref(h)))
        self._h, self._keep = h, (options, problem)

    def minimize(self, summary=None, max_num_iterations=-1):
        summary = summary if summary is not None else Solver.Summary()
        check(lib.sk_solver_minimize(self._h, int(max_num_iterations), summary._h))
        return summary

    def timeSchurProduct(self, reps=50):
        """Measurement aid: mean device ms of one implicit Schur-complement product on the last linearisation."""
        ms = C.c_double()
        check(lib.sk_solver_time_schur_product(self._h, int(reps), C.byref(ms)))
        return ms.value

    def close(self, _destroy=_destroy):
        if getattr(self, "_h", None):
            _destroy("sk_solver_destroy", self._h)
            self._h = None

    __del__ = close


class Communicator:
    """One rank of the point-partitioned multi-GPU solve (no reference counterpart)."""

    def __init__(self, unique_id, rank, world_size):
        h = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), _abi.COMM_UNIQUE_ID_BYTES)
        check(lib.sk_comm_create(buf, int(rank), int(world_size), C.byref(h)))
        self._h, self.rank, self.world_size = h, rank, world_size

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(_abi.COMM_UNIQUE_ID_BYTES)
        check(lib.sk_comm_get_unique_id(buf))
        return buf.raw

    def __del__(self, _destroy=_destroy):
        if getattr(self, "_h", None):
            _destroy("sk_comm_destroy", self._h)
            self._h = None


def partition_points(point_ptr, world_size):
    """sk_partition_points: contiguous point ranges balanced by observation count (host only)."""
    ptr = np.ascontiguousarray(point_ptr, dtype=np.int64)
    out = np.zeros(world_size + 1, dtype=np.int64)
    check(lib.sk_partition_points(ptr.size - 1, _vp(ptr), int(world_size), _vp(out)))
    return out


class BalProblem:
    """BalProblem (SimpleBundleAdjuster.scala:18-77) with device-resident parameters."""

    def __init__(self, numCameras, numPoints, cameraIndex, pointIndex, observations, parameters):
        self.numCameras, self.numPoints = int(numCameras), int(numPoints)
        self.cameraIndex = np.ascontiguousarray(cameraIndex, dtype=np.int32)
        self.pointIndex = np.ascontiguousarray(pointIndex, dtype=np.int32)
        self.observations = np.ascontiguousarray(observations, dtype=np.float64).ravel()
        self.numObservations = int(self.cameraIndex.size)
        self.numParameters = 9 * self.numCameras + 3 * self.numPoints
        self.parameters = parameters            # DoubleArray

    @classmethod
    def fromArrays(cls, data):
        """From a skeres_b200.synth.BalData (or anything with the same fields)."""
        return cls(data.num_cameras, data.num_points, data.camera_index, data.point_index, data.observations,
                   DoubleArray.fromArray(data.parameters))

    @classmethod
    def fromFile(cls, path):
        h = C.c_void_p()
        check(lib.sk_bal_problem_from_file(str(path).encode(), C.byref(h)))
        try:
            nc, npt, no = (lib.sk_bal_problem_num_cameras(h), lib.sk_bal_problem_num_points(h), lib.sk_bal_problem_num_observations(h))
            cam = np.ctypeslib.as_array(C.cast(lib.sk_bal_problem_camera_index(h), C.POINTER(C.c_int32)), (no,)).copy()
            pt = np.ctypeslib.as_array(C.cast(lib.sk_bal_problem_point_index(h), C.POINTER(C.c_int32)), (no,)).copy()
            obs = np.ctypeslib.as_array(C.cast(lib.sk_bal_problem_observations(h), C.POINTER(C.c_double)), (2 * no,)).copy()
            params = DoubleArray(9 * nc + 3 * npt)
            src = lib.sk_bal_problem_parameters(h)
            tmp = np.empty(9 * nc + 3 * npt)
            check(lib.sk_double_array_download(src, 0, _vp(tmp), tmp.size))
            params.copyFrom(tmp)
        finally:
            lib.sk_bal_problem_destroy(h)
        return cls(nc, npt, cam, pt, obs, params)

    def mutableCameras(self):
        return self.parameters.toPointer()

    def mutablePoints(self):
        return self.parameters.slice(9 * self.numCameras)

    def mutableCameraForObservation(self, i):
        return self.mutableCameras().slice(int(self.cameraIndex[i]) * 9)

    def mutablePointForObservation(self, i):
        return self.mutablePoints().slice(int(self.pointIndex[i]) * 3)

    def blockOffsets(self, o0=0, o1=None):
        """Offsets of mutableCameraForObservation(i) / mutablePointForObservation(i) for observations [o0, o1), interleaved."""
        o1 = self.numObservations if o1 is None else o1
        off = np.empty((o1 - o0, 2), dtype=np.int64)
        check(lib.sk_bal_block_offsets(o1 - o0, _vp(self.cameraIndex[o0:o1]), _vp(self.pointIndex[o0:o1]), self.numCameras, self.numPoints,
                                       _vp(off)))
        return off

    def buildProblem(self, loss=None, functor_id=_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR):
        """The residual-block loop of SimpleBundleAdjuster.scala:134-145, in bulk.  functor_id: SnavelyReprojectionError, or a
        functor of the same shape (2; 9, 3; constants = the observation) defined from source (SourceCostFunctor.define)."""
        problem = Problem()
        loss = loss if loss is not None else PredefinedLossFunctions.trivialLoss()
        problem.addResidualBlocks(functor_id, self.observations.reshape(-1, 2), loss, self.parameters, self.blockOffsets())
        return problem

    def localRange(self, rank, world_size):
        """Observation range [o0, o1) of this rank's points -- the partition of sk_partition_points (first point whose
        observation prefix reaches rank / world of the total), found without building the point CSR: observations are
        sorted by point, so a boundary is the first point start at or behind the target observation."""
        n = self.numObservations
        pt = self.pointIndex

        def boundary(r):
            if r <= 0:
                return 0
            if r >= world_size:
                return n
            j = (n * r) // world_size
            while 0 < j < n and pt[j] == pt[j - 1]:
                j += 1
            return j

        return boundary(rank), boundary(rank + 1)

    def buildLocalProblem(self, rank, world_size, loss=None):
        """Multi-GPU ingestion in O(local observations): only the residual blocks of this rank's points, every camera
        declared (Problem::AddParameterBlock).  Solve with Options.residual_blocks_are_local = 1 and a communicator."""
        o0, o1 = self.localRange(rank, world_size)
        # only this rank's share is touched (the whole list is N times longer): sorted by point inside, true point boundaries outside
        loc = self.pointIndex[o0:o1]
        assert np.all(loc[1:] >= loc[:-1]), "rank-local ingestion needs observations sorted by point"
        assert (o0 == 0 or self.pointIndex[o0 - 1] < self.pointIndex[o0]) and \
               (o1 == self.numObservations or o1 == o0 or self.pointIndex[o1 - 1] < self.pointIndex[o1]), "observations are not sorted by point"

        problem = Problem()
        loss = loss if loss is not None else PredefinedLossFunctions.trivialLoss()
        problem.addParameterBlocks(self.parameters, 9 * np.arange(self.numCameras, dtype=np.int64), 9)
        off = self.blockOffsets(o0, o1)
        problem.addResidualBlocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, self.observations[2 * o0:2 * o1].reshape(-1, 2), loss,
                                  self.parameters, off)
        return problem


def build_share_problem(bal, loss=None):
    """Multi-GPU ingestion from a rank's OWN share: `bal` holds only this rank's residual blocks (global camera / point
    indices) and the full parameter array; every camera is declared (Problem::AddParameterBlock).  Solve with
    Options.residual_blocks_are_local = 1 and a communicator."""
    problem = Problem()
    loss = loss if loss is not None else PredefinedLossFunctions.trivialLoss()
    problem.addParameterBlocks(bal.parameters, 9 * np.arange(bal.numCameras, dtype=np.int64), 9)
    problem.addResidualBlocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, bal.observations.reshape(-1, 2), loss, bal.parameters, bal.blockOffsets())
    return problem


def curve_fit_batch_solve(options, x, y, mc, want_details=True):
    """sk_curve_fit_batch_solve on device arrays x, y [n_obs][n], mc [2][n] (DoubleArray)."""
    n = mc.n // 2
    n_obs = x.n // n
    summary = Solver.Summary()
    ic = np.empty(n) if want_details else None
    fc = np.empty(n) if want_details else None
    it = np.empty(n, dtype=np.int32) if want_details else None
    tt = np.empty(n, dtype=np.int32) if want_details else None
    check(lib.sk_curve_fit_batch_solve(C.byref(options._o), n, n_obs, x._h, y._h, mc._h, _vp(ic), _vp(fc), _vp(it), _vp(tt), summary._h))
    return summary, ic, fc, it, tt
