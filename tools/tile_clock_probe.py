"""Development probe (needs a library built with -DSK_TILE_CLK, see profiles/r02_v14_tile_clocks.md): per-tile SM cycles and per-CTA
start / end times of the stand-alone implicit-Schur product on the Venice-1778 shape, two passes, written to an .npz."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skeres_b200 import _abi, api, synth
from skeres_b200 import _lib
out = sys.argv[1]
d = synth.make_bal("venice-1778", seed=1)
bal = api.BalProblem.fromArrays(d); prob = bal.buildProblem()
o = api.Solver.Options(); o.setLinearSolverType(_abi.ITERATIVE_SCHUR); o.setPreconditionerType(_abi.SCHUR_JACOBI)
o.setMaxNumIterations(3)
os.environ["SKERES_PCG"] = "sequence"        # the stand-alone product kernel is the instrumented one
solver = api.PreparedSolver(o, prob)
s = solver.minimize()
L = C.CDLL(_lib.LIB_PATH)
L.sk_debug_tile_clocks.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
n = 1 << 17
res = {}
for rep in range(3):
    ms = solver.timeSchurProduct(1 if rep < 2 else 50)
    clk = np.zeros(n, dtype=np.uint32); cta = np.zeros((2, 2048), dtype=np.uint64)
    rc = L.sk_debug_tile_clocks(clk.ctypes.data, n, cta.ctypes.data, 2048)
    assert rc == 0, rc
    res[f"clk{rep}"] = clk; res[f"cta{rep}"] = cta; res[f"ms{rep}"] = ms
np.savez_compressed(out, **res)
print("saved", out, [res[f"ms{r}"] for r in range(3)])
