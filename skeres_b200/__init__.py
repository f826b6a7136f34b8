"""skeres_b200 — B200-native Levenberg–Marquardt solver core behind the skeres (Scala/Ceres) API.

`skeres_b200.api` mirrors the reference interface over the C ABI of libskeres.so (include/skeres.h);
`skeres_b200.synth` generates the synthetic BAL-shaped inputs.  Importing the package itself is
cheap; the shared library is loaded on first use of the API names below and there is no fallback
when it is missing.
"""
_API = ("DoubleArray", "DoublePointer", "LossFunction", "PredefinedLossFunctions", "CostFunction", "AutoDiffCostFunctor",
        "SnavelyReprojectionError", "ExponentialResidual", "Problem", "Solver", "ceres", "Communicator", "BalProblem",
        "LinearSolverType", "PreconditionerType", "MinimizerType", "TerminationType", "SkeresError", "functor_info",
        "partition_points", "curve_fit_batch_solve", "PreparedSolver")


def __getattr__(name):
    if name in _API:
        from . import api
        return getattr(api, name)
    raise AttributeError(name)
