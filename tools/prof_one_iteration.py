"""Profiling helper: a Venice-1778-shaped problem, a few LM iterations (ITERATIVE_SCHUR + SCHUR_JACOBI)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skeres_b200 import _abi, api, synth
its = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
c, p, o = synth.SHAPES["venice-1778"]
d = synth.make_bal(n_cam=int(c * scale), n_pt=int(p * scale), n_obs=int(o * scale), seed=1)
bal = api.BalProblem.fromArrays(d)
problem = bal.buildProblem()
opt = api.Solver.Options()
opt.setLinearSolverType(_abi.ITERATIVE_SCHUR); opt.setPreconditionerType(_abi.SCHUR_JACOBI); opt.setMaxNumIterations(its)
s = api.Solver.Summary()
api.ceres.solve(opt, problem, s)
print(s.briefReport(), [r.linear_solver_iterations for r in s.iterations], "launches", s.num_kernel_launches)
