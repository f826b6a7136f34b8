"""BASELINE.json configs[3]: 1M independent CurveFitting-sized LM problems (DENSE_QR per problem, one launch per LM
iteration).  Prints problems/s, LM iterations/s over the batch and the achieved fraction of the HBM roofline for
k_batch_iterate: algorithmic bytes per launch = N_active * (67 * 16 + 2 * 16)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skeres_b200 import _abi, api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
x, y, truth = synth.make_curve_fit_batch(n, seed=1)
xa, ya = api.DoubleArray.fromArray(x), api.DoubleArray.fromArray(y)
o = api.Solver.Options(); o.setLinearSolverType(_abi.DENSE_QR); o.setMaxNumIterations(25); o.profile_kernels = 1
best = None
for rep in range(4):
    mc = api.DoubleArray(2 * n)
    t = time.time()
    summary, ic, fc, it, tt = api.curve_fit_batch_solve(o, xa, ya, mc)
    wall = time.time() - t
    dev = 1e-3 * float(sum(summary.data.kernel_ms[:]))          # CUDA events around every launch (k_batch_init + k_batch_iterate)
    if rep and (best is None or dev < best[0]):
        best = (dev, wall, summary.num_kernel_launches, it.copy())
dev, wall, launches, it = best
total_it = int(it.sum())
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6536.7) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6536.7
bytes_total = total_it * (67 * 16 + 2 * 16)
print(json.dumps({"workload": f"{n} CurveFitting-sized problems, DENSE_QR, max 25 LM iterations", "device_s": dev, "wall_s": wall,
                  "launches": int(launches), "problems_per_s": n / dev, "lm_iterations_total": total_it,
                  "lm_iterations_per_s": total_it / dev, "iterations_per_problem": [int(it.min()), float(it.mean()), int(it.max())],
                  "algorithmic_GBps": bytes_total / dev / 1e9, "hbm_peak_GBps": peak, "frac_of_hbm": bytes_total / dev / 1e9 / peak,
                  "all_converged": bool(np.all(tt == _abi.CONVERGENCE))}))
