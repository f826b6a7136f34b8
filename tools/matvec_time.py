"""Development helper: device time of the implicit Schur product alone (sk_solver_time_schur_product) on the Venice-1778 shape.
Kernel variant by environment: SKERES_MATVEC=tma|classic|rows, SKERES_MATVEC_SUMS=chunked|serial, SKERES_LIB=<variant .so>."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skeres_b200 import _abi, api, synth
shape = sys.argv[1] if len(sys.argv) > 1 else "venice-1778"
d = synth.make_bal(shape, seed=1)
bal = api.BalProblem.fromArrays(d); prob = bal.buildProblem()
o = api.Solver.Options(); o.setLinearSolverType(_abi.ITERATIVE_SCHUR); o.setPreconditionerType(_abi.SCHUR_JACOBI)
o.setMaxNumIterations(3)
solver = api.PreparedSolver(o, prob)
s = solver.minimize()
modes = [int(m) for m in os.environ.get("TIME_MODES", "0").split(",")]
for mode in modes[:-1]:
    os.environ["SKERES_TIME_MODE"] = str(mode)
    t = [solver.timeSchurProduct(200) for _ in range(3)]
    print("time mode %d: %.4f ms per product (runs %s)" % (mode, min(t), ["%.4f" % m for m in t]))
os.environ["SKERES_TIME_MODE"] = str(modes[-1])
ms = [solver.timeSchurProduct(200) for _ in range(3)]
nbytes = d.num_observations * 196 + d.num_points * 72 + d.num_cameras * 144
print("variant matvec=%s sums=%s lib=%s : %.4f ms per product (runs %s) = %.0f GB/s algorithmic ; solve cost %.9e pcg %s" % (
    os.environ.get("SKERES_MATVEC", "tma"), os.environ.get("SKERES_MATVEC_SUMS", "chunked"), os.path.basename(os.environ.get("SKERES_LIB", "libskeres.so")),
    min(ms), ["%.4f" % m for m in ms], nbytes / min(ms) / 1e6, s.final_cost, [r.linear_solver_iterations for r in s.iterations]))
