"""Development aid: LM rows of the oracle and of the device side by side on the long-track cases (and the converged
IDENTITY case).  `python tools/long_track_rows.py full` prints the full SCHUR_JACOBI runs the *_full_run test compares."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import numpy as np
import oracle_lib as oracle
import skeres_b200 as sk
from skeres_b200 import _abi, synth
import test_gpu_parity as T

CONV = dict(eta=1e-10, max_linear_solver_iterations=3000, max_num_iterations=3)
CASES = {
    "short": [
        ("big/SJ 3 its", T.LONG_TRACK_CASE, _abi.SCHUR_JACOBI, dict(max_num_iterations=3)),
        ("small/SJ 4 its", T.LONG_TRACK_SMALL, _abi.SCHUR_JACOBI, dict(max_num_iterations=4)),
        ("big/J", T.LONG_TRACK_CASE, _abi.JACOBI, {}),
        ("big/J 1 it eta 1e-8", T.LONG_TRACK_CASE, _abi.JACOBI, dict(max_num_iterations=1, eta=1e-8, max_linear_solver_iterations=3000)),
    ],
    "full": [
        ("big/SJ full", T.LONG_TRACK_CASE, _abi.SCHUR_JACOBI, {}),
        ("small/SJ full", T.LONG_TRACK_SMALL, _abi.SCHUR_JACOBI, {}),
        ("small-3/IDENTITY converged", dict(shape="small", seed=3), _abi.IDENTITY, CONV),
        ("small-3/JACOBI converged", dict(shape="small", seed=3), _abi.JACOBI, CONV),
    ],
}
for name, case, prec, kw in CASES[sys.argv[1] if len(sys.argv) > 1 else "short"]:
    d = synth.make_bal(**case)
    p, so = T.oracle_ba(oracle, d, _abi.ITERATIVE_SCHUR, prec, **kw)
    bal, s = T.gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, prec, **kw)
    print(f"== {name}: obs {d.num_observations}  term {s.termination_type}/{so.termination_type}  rows {len(s.iterations)}/{len(so.iterations)}")
    for a, b in zip(s.iterations, so.iterations):
        print(f"  it {a.iteration:2d} cost {a.cost:.10e} {b.cost:.10e} rel {abs(a.cost-b.cost)/b.cost:.1e}  radius {a.trust_region_radius:.6e} {b.trust_region_radius:.6e}"
              f"  pcg {a.linear_solver_iterations:4d} {b.linear_solver_iterations:4d}  ok {a.step_is_successful}/{b.step_is_successful}  gmax rel {abs(a.gradient_max_norm-b.gradient_max_norm)/max(b.gradient_max_norm,1e-30):.1e}")
    print(f"  final cost rel {abs(s.final_cost-so.final_cost)/so.final_cost:.2e}  param diff {T.rel_param_diff(bal.parameters.toArray(), p.params):.2e}")
