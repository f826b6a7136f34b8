"""Development helper: LM rows of the Huber BA case, CUDA path vs oracle, to find where they separate."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from skeres_b200 import _abi, api, synth
d = synth.make_bal("small", seed=8)
d.observations[::37] += 25.0
p = O.OracleProblem(d.parameters)
p.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets(), _abi.LOSS_HUBER, 1.5)
oo = _abi.default_options(); oo.linear_solver_type = _abi.DENSE_SCHUR
so = p.solve(oo)
bal = api.BalProblem.fromArrays(d)
prob = bal.buildProblem(api.PredefinedLossFunctions.huberLoss(1.5))
o = api.Solver.Options(); o.setLinearSolverType(_abi.DENSE_SCHUR)
s = api.Solver.Summary(); api.ceres.solve(o, prob, s)
print("gpu   :", s.message, "final %.12e" % s.final_cost, len(s.iterations))
print("oracle:", so.message, "final %.12e" % so.final_cost, len(so.iterations))
first = None
for k, (a, b) in enumerate(zip(s.iterations, so.iterations)):
    rel = abs(a.cost - b.cost) / max(abs(b.cost), 1e-300)
    flag = "" if (a.step_is_successful == b.step_is_successful and rel < 1e-6) else "  <-- differs"
    if flag and first is None: first = k
    print(f"{k:3d} ok={a.step_is_successful}/{b.step_is_successful} cost {a.cost:.10e} / {b.cost:.10e} rel {rel:.1e} rho {a.relative_decrease:+.4e} / {b.relative_decrease:+.4e} radius {a.trust_region_radius:.3e} / {b.trust_region_radius:.3e}{flag}")
print("first differing row:", first)
