"""Development aid: the bench's end-to-end sequence (host arrays -> problem -> solve -> download) with the host trace on."""
import os, sys, time
os.environ["SKERES_TRACE_HOST"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skeres_b200 import _abi, api, synth
data = synth.make_bal("venice-1778", seed=1)
for rep in range(3):
    t0 = time.time()
    bal = api.BalProblem.fromArrays(data); t1 = time.time()
    problem = bal.buildProblem(); t2 = time.time()
    opt = api.Solver.Options(); opt.setLinearSolverType(_abi.ITERATIVE_SCHUR); opt.setPreconditionerType(_abi.SCHUR_JACOBI); opt.setMaxNumIterations(10)
    summ = api.Solver.Summary()
    api.ceres.solve(opt, problem, summ); t3 = time.time()
    out = bal.parameters.toArray(); t4 = time.time()
    print(f"rep {rep}: fromArrays {t1-t0:.3f}  buildProblem {t2-t1:.3f}  solve {t3-t2:.3f} (preprocessor {summ.preprocessor_time_in_seconds:.3f}, minimizer {summ.minimizer_time_in_seconds:.3f})  toArray {t4-t3:.3f}  total {t4-t0:.3f}", flush=True)
    del problem, bal
