"""Development helper: matvec kernel A/B. Run once per kernel (SKERES_MATVEC unset / =pf); prints costs with
full precision (the two kernels must agree bitwise), PCG counts and the matvec time per executed launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skeres_b200 import _abi, api, synth
d = synth.make_bal("venice-1778", seed=1)
bal = api.BalProblem.fromArrays(d); prob = bal.buildProblem()
o = api.Solver.Options(); o.setLinearSolverType(_abi.ITERATIVE_SCHUR); o.setPreconditionerType(_abi.SCHUR_JACOBI)
o.setMaxNumIterations(8); o.profile_kernels = int(os.environ.get("AB_PROFILE", "2"))
solver = api.PreparedSolver(o, prob)
x0 = api.DoubleArray.fromArray(d.parameters)
solver.minimize()                      # warm-up
bal.parameters.copyFromArray(x0)
s = solver.minimize()
kt = s.kernel_times()
ms, n = kt["schur_matvec"]
print("kernel", os.environ.get("SKERES_MATVEC", "default"))
print("costs", " ".join(repr(r.cost) for r in s.iterations))
print("pcg", [r.linear_solver_iterations for r in s.iterations])
print("matvec: %.3f ms over %d executed launches = %.4f ms each ; device time %.2f ms" % (ms, n, ms / max(n, 1), 1e3 * s.minimizer_device_time_in_seconds))
print("families", {k: (round(v[0], 2), v[1]) for k, v in kt.items() if v[1]})
