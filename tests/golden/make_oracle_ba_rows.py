"""Generates oracle_<case>_rows.json: the ORACLE's (not the reference's -- libceres cannot be built here) LM trajectory of
a BASELINE.json-sized synthetic bundle adjustment, solved with ITERATIVE_SCHUR + SCHUR_JACOBI.  The oracle needs minutes
at these sizes, so the `-m gpu` tests compare the device solve with this committed output row by row instead of
re-running the oracle on the GPU box (tests/test_gpu_parity.py: test_venice_rows_follow_the_oracle_fixture, ...).

Run from the repo root:   python tests/golden/make_oracle_ba_rows.py venice-1778 13
                          python tests/golden/make_oracle_ba_rows.py final-13682 3
The problem is synth.make_bal(shape, seed=1): the same input bench.py times."""
import json, os, sys, time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..")); sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np
import oracle_lib as oracle
from skeres_b200 import _abi, synth

N_POINTS_KEPT = 300      # parameter digest: every camera + this many evenly spaced points


def digest_indices(n_cam, n_pt):
    pts = np.unique(np.linspace(0, n_pt - 1, N_POINTS_KEPT).astype(np.int64))
    idx = [np.arange(9 * n_cam, dtype=np.int64)]
    for p in pts:
        idx.append(9 * n_cam + 3 * p + np.arange(3, dtype=np.int64))
    return np.concatenate(idx)


if __name__ == "__main__":
    shape, iters = sys.argv[1], int(sys.argv[2])
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    t0 = time.time()
    d = synth.make_bal(shape, seed=seed)
    p = oracle.OracleProblem(d.parameters)
    p.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets(), _abi.LOSS_TRIVIAL, 0.0)
    o = _abi.default_options()
    o.linear_solver_type = _abi.ITERATIVE_SCHUR
    o.preconditioner_type = _abi.SCHUR_JACOBI
    o.max_num_iterations = iters
    s = p.solve(o, threads=os.cpu_count())
    idx = digest_indices(d.num_cameras, d.num_points)
    out = {"case": {"shape": shape, "seed": seed, "max_num_iterations": iters, "linear_solver": "ITERATIVE_SCHUR", "preconditioner": "SCHUR_JACOBI"},
           "provenance": "oracle/oracle.cc (Ceres-algorithm restatement, NOT libceres), -ffp-contract=off, %d threads" % (os.cpu_count() or 1),
           "input_checksum": float(np.sum(d.parameters) + np.sum(d.observations)),
           "n_cam": d.num_cameras, "n_pt": d.num_points, "n_obs": d.num_observations,
           "termination_type": int(s.termination_type), "message": s.message, "initial_cost": s.initial_cost, "final_cost": s.final_cost,
           "num_successful_steps": s.num_successful_steps, "num_unsuccessful_steps": s.num_unsuccessful_steps,
           "iterations": [{"iteration": it.iteration, "cost": it.cost, "cost_change": it.cost_change, "trust_region_radius": it.trust_region_radius,
                           "relative_decrease": it.relative_decrease, "step_norm": it.step_norm,
                           "step_is_valid": int(it.step_is_valid), "step_is_successful": int(it.step_is_successful),
                           "linear_solver_iterations": it.linear_solver_iterations, "gradient_max_norm": it.gradient_max_norm,
                           "gradient_norm": it.gradient_norm}
                          for it in s.iterations],
           "param_digest_points": N_POINTS_KEPT,
           "param_digest": [float(v) for v in p.params[idx]],
           "oracle_wall_s": time.time() - t0}
    name = "oracle_%s_rows.json" % shape.replace("-", "_")
    json.dump(out, open(os.path.join(HERE, name), "w"))
    print(name, len(s.iterations), s.initial_cost, s.final_cost, [it.linear_solver_iterations for it in s.iterations], "%.0f s" % (time.time() - t0))
