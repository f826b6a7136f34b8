"""Development helper: how far apart are the residual vectors of two optima reached by different linear solvers?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from skeres_b200 import _abi, api, synth
d = synth.make_bal("small", seed=4)
def res_at(x):
    p = O.OracleProblem(x); p.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets()); return p.evaluate()[1]
p = O.OracleProblem(d.parameters); p.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets())
oo = _abi.default_options(); oo.linear_solver_type = _abi.DENSE_SCHUR; oo.function_tolerance = 1e-12; oo.max_num_iterations = 60
so = p.solve(oo)
bal = api.BalProblem.fromArrays(d); prob = bal.buildProblem()
o = api.Solver.Options(); o.setLinearSolverType(_abi.ITERATIVE_SCHUR); o.setPreconditionerType(_abi.SCHUR_JACOBI)
o.function_tolerance = 1e-12; o.max_num_iterations = 60; o.eta = 1e-3
s = api.Solver.Summary(); api.ceres.solve(o, prob, s)
x = bal.parameters.toArray()
rg, ro = res_at(x), res_at(p.params)
print("cost gpu %.12e oracle %.12e rel %.2e" % (s.final_cost, so.final_cost, abs(s.final_cost - so.final_cost) / so.final_cost))
print("max |r_gpu - r_oracle| = %.3e px ; rms residual %.3f px ; max rel param diff %.3e" % (np.max(np.abs(rg - ro)), np.sqrt(np.mean(ro**2)), np.max(np.abs(x - p.params) / np.maximum(np.abs(p.params), 1e-2))))
