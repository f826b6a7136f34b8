"""Generates oracle_long_tracks_dense_schur.json: the ORACLE's (not the reference's) DENSE_SCHUR solve of a synthetic
problem with tracks longer than one device tile.  The oracle's dense reduced solve takes minutes at this camera
count, so the GPU test compares against this committed output instead of re-running it on the GPU box.
Run from the repo root:  python tests/golden/make_oracle_long_tracks.py"""
import json, os, sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..")); sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np
import oracle_lib as oracle
from skeres_b200 import _abi, synth

CASE = dict(n_cam=270, n_pt=220, n_obs=2200, seed=11, long_tracks=(257, 270, 264))

if __name__ == "__main__":
    d = synth.make_bal(**CASE)
    p = oracle.OracleProblem(d.parameters)
    p.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets(), _abi.LOSS_TRIVIAL, 0.0)
    o = _abi.default_options()
    o.linear_solver_type = _abi.DENSE_SCHUR
    s = p.solve(o)
    out = {"case": {k: (list(v) if isinstance(v, tuple) else v) for k, v in CASE.items()},
           "input_checksum": float(np.sum(d.parameters) + np.sum(d.observations)),
           "termination_type": int(s.termination_type), "initial_cost": s.initial_cost, "final_cost": s.final_cost,
           "num_successful_steps": s.num_successful_steps, "num_unsuccessful_steps": s.num_unsuccessful_steps,
           "iterations": [{"iteration": it.iteration, "cost": it.cost, "trust_region_radius": it.trust_region_radius,
                           "step_is_valid": int(it.step_is_valid), "step_is_successful": int(it.step_is_successful),
                           "linear_solver_iterations": it.linear_solver_iterations, "gradient_max_norm": it.gradient_max_norm}
                          for it in s.iterations],
           "params": [float(v) for v in p.params]}
    json.dump(out, open(os.path.join(HERE, "oracle_long_tracks_dense_schur.json"), "w"))
    print(len(s.iterations), s.initial_cost, s.final_cost)
