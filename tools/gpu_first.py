"""First end-to-end check on a B200: device path vs the CPU oracle (development helper)."""
import json, sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from skeres_b200 import _abi, synth, api

def rows(s): return [(r.iteration, r.cost, r.cost_change, r.gradient_max_norm, r.step_norm, r.relative_decrease, r.trust_region_radius, r.linear_solver_iterations) for r in s.iterations]

# 1. golden vectors at the evaluate boundary
cf = api.CostFunction(_abi.FUNCTOR_TEST_BILINEAR_SCALAR, [1.0])
print("bilinear", cf.evaluate_host([[1, 2], [3, 4]]))
cf = api.CostFunction(_abi.FUNCTOR_TEST_BILINEAR_VECTOR3, [1.0])
print("vec3", cf.evaluate_host([[1, 2], [3, 4]]))
# 2. curve fitting
d = json.load(open(os.path.join(ROOT, "tests/golden/curve_fitting_data.json")))
m = api.DoubleArray(1); c = api.DoubleArray(1)
loss = api.PredefinedLossFunctions.trivialLoss()
prob = api.Problem()
for x, y in zip(d["x"], d["y"]):
    prob.addResidualBlock(api.ExponentialResidual(x, y).toAutoDiffCostFunction(), loss, m.toPointer(), c.toPointer())
opt = api.Solver.Options(); opt.setMaxNumIterations(25); opt.setLinearSolverType(_abi.DENSE_QR); opt.setMinimizerProgressToStdout(True)
summ = api.Solver.Summary()
api.ceres.solve(opt, prob, summ)
print(summ.briefReport()); print("Final", m.get(0), c.get(0), summ.message)
# 3. BA
for shape in ["small", "ladybug-49"]:
    data = synth.make_bal(shape, seed=1)
    for lst, prec in [(_abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI), (_abi.DENSE_SCHUR, _abi.JACOBI)]:
        op = O.OracleProblem(data.parameters)
        op.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, data.observations.reshape(-1, 2), data.block_offsets())
        oo = _abi.default_options(); oo.linear_solver_type = lst; oo.preconditioner_type = prec
        so = op.solve(oo)
        bal = api.BalProblem.fromArrays(data)
        prob = bal.buildProblem()
        opt = api.Solver.Options(); opt.setLinearSolverType(lst); opt.setPreconditionerType(prec); opt.profile_kernels = 1
        summ = api.Solver.Summary()
        t = time.time(); api.ceres.solve(opt, prob, summ); dt = time.time() - t
        x = bal.parameters.toArray()
        print(f"== {shape} lst={lst}: gpu {summ.message} | oracle {so.message}")
        print("   gpu   ", summ.initial_cost, summ.final_cost, len(summ.iterations), [r.linear_solver_iterations for r in summ.iterations], "%.3fs" % dt)
        print("   oracle", so.initial_cost, so.final_cost, len(so.iterations), [r.linear_solver_iterations for r in so.iterations])
        print("   max rel param diff", np.max(np.abs(x - op.params) / np.maximum(np.abs(op.params), 1e-3)))
        for a, b in zip(rows(summ), rows(so)):
            print("    ", " ".join("%.6e" % v if isinstance(v, float) else str(v) for v in a)); print("   o ", " ".join("%.6e" % v if isinstance(v, float) else str(v) for v in b))
        print("   kernels", {k: (round(v[0], 3), v[1]) for k, v in summ.kernel_times().items()})
