"""BASELINE configs[1]: SimpleBundleAdjuster on the Ladybug-49 shape with the explicit reduced camera system (DENSE_SCHUR, the
reference's own choice at SimpleBundleAdjuster.scala:148, or SPARSE_SCHUR): LM iteration time and the share of the dense
factorisation (FP64 tensor-core Cholesky update, dense_kernels.cu).  Prints one JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skeres_b200 import _abi, api, synth
shape = sys.argv[1] if len(sys.argv) > 1 else "ladybug-49"
lst = _abi.SPARSE_SCHUR if (len(sys.argv) > 2 and sys.argv[2] == "sparse") else _abi.DENSE_SCHUR
d = synth.make_bal(shape, seed=1)
out = {}
for profile in (0, 1):
    bal = api.BalProblem.fromArrays(d)
    o = api.Solver.Options(); o.setLinearSolverType(lst); o.profile_kernels = profile
    solver = api.PreparedSolver(o, bal.buildProblem())
    x0 = api.DoubleArray.fromArray(d.parameters)
    solver.minimize()                                   # warm-up
    bal.parameters.copyFromArray(x0)
    s = solver.minimize()
    its = max(s.num_iterations - 1, 0)
    if profile == 0:
        out.update({"workload": f"{shape}: {d.num_cameras} cameras, {d.num_points} points, {d.num_observations} observations, "
                                f"{_abi.LINEAR_SOLVER_NAMES[lst] if hasattr(_abi, 'LINEAR_SOLVER_NAMES') else lst}",
                    "reduced_system_size": 9 * d.num_cameras, "lm_iterations": its, "final_cost": s.final_cost,
                    "ms_per_lm_iteration": 1e3 * s.data.minimizer_device_time_in_seconds / max(its, 1), "kernel_launches": int(s.num_kernel_launches)})
    else:
        kt = s.kernel_times()
        tot = sum(v[0] for v in kt.values())
        out["kernel_family_ms_per_lm_iteration"] = {k: round(v[0] / max(its, 1), 4) for k, v in kt.items() if v[1]}
        out["dense_share_of_kernel_time"] = kt["dense"][0] / tot if tot > 0 else None
    solver.close()
print(json.dumps(out))
