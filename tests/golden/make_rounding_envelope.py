"""Generates schur_jacobi_rounding_envelope.json: how far the ORACLE moves away from ITSELF when nothing but the rounding of
its arithmetic changes.  Two kinds of perturbation, both leaving the algorithm untouched:

  * `fma`: the same oracle.cc compiled with -ffp-contract=fast (the compiler may fuse a*b+c) instead of
    -ffp-contract=off -- what any other compiler, CPU or GPU may legitimately do to the same source;
  * `variant 1..3`: algebraically identical re-orderings of the Schur eliminator (oracle.cc: g_schur_variant):
    G^T (E^-1 G) instead of (G^T E^-1) G; sum F^T F and sum G^T E^-1 G accumulated separately and subtracted once;
    points eliminated in reverse order.

Why this exists (parity gap 1): with ITERATIVE_SCHUR + SCHUR_JACOBI at Ceres' default eta = 0.1 the
parameters a solve ends at are NOT a well-conditioned function of the input -- the oracle differs from its own FMA build by
7e-4 on the Ladybug-49 shape and by 4e-2 on the long-track case, while the cost agrees to 1e-9 (the synthetic problems do
not fix the gauge, and a truncated CG step depends on the preconditioner's last bits).  With the JACOBI preconditioner, or
with converged linear solves (eta = 1e-10), the same perturbations move the parameters by 1e-12 .. 4e-7.  So the 1e-5
parameter tolerance of the north star is checked where it is attainable (converged solves, JACOBI, exact Schur solvers),
and for the truncated SCHUR_JACOBI runs the device is held to a multiple of this envelope instead.

Run from the repo root:  python tests/golden/make_rounding_envelope.py"""
import ctypes as C
import json, os, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np
import oracle_lib as O
from skeres_b200 import _abi, synth

CASES = {
    "ladybug-49": dict(shape="ladybug-49", seed=1),
    "small-3": dict(shape="small", seed=3),
    "long-tracks": dict(n_cam=600, n_pt=1500, n_obs=12000, seed=5, long_tracks=(257, 600, 513)),
}
PRECS = {"SCHUR_JACOBI": _abi.SCHUR_JACOBI, "JACOBI": _abi.JACOBI, "IDENTITY": _abi.IDENTITY}
TIGHT = dict(eta=1e-10, max_linear_solver_iterations=3000)


def rel(a, b, floor=1e-2):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def solve(so_path, d, prec, variant=0, **kw):
    O._lib, O._SO = None, so_path
    L = O.lib()
    L.oracle_set_schur_rounding_variant.argtypes = [C.c_int]
    L.oracle_set_schur_rounding_variant(variant)
    p = O.OracleProblem(d.parameters)
    p.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets())
    o = _abi.default_options()
    o.linear_solver_type, o.preconditioner_type = _abi.ITERATIVE_SCHUR, prec
    for k, v in kw.items():
        setattr(o, k, v)
    s = p.solve(o, threads=4)
    L.oracle_set_schur_rounding_variant(0)
    return p.params.copy(), s


def compare(x, s, x0, s0):
    rows = min(len(s.iterations), len(s0.iterations))
    return {"param_rel_diff": rel(x, x0), "final_cost_rel_diff": abs(s.final_cost - s0.final_cost) / abs(s0.final_cost),
            "rows": [len(s0.iterations), len(s.iterations)],
            "max_row_cost_rel_diff": max(abs(a.cost - b.cost) / abs(b.cost) for a, b in zip(s.iterations[:rows], s0.iterations[:rows])),
            "pcg_iterations": [[r.linear_solver_iterations for r in s0.iterations], [r.linear_solver_iterations for r in s.iterations]]}


if __name__ == "__main__":
    base = os.path.join(ROOT, "oracle", "liboracle.so")
    fma = os.path.join(tempfile.mkdtemp(), "liboracle_fma.so")
    subprocess.run(["g++", "-O3", "-march=x86-64-v3", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=fast", "-w", "-shared", "-o", fma,
                    os.path.join(ROOT, "oracle", "oracle.cc")], check=True)
    out = {"what": __doc__.split("\n\n")[0], "cases": {}}
    for cname, ckw in CASES.items():
        d = synth.make_bal(**ckw)
        for pname, prec in PRECS.items():
            if pname == "IDENTITY" and cname != "small-3":
                continue
            for ename, ekw in (("eta0.1", {}), ("eta1e-10", TIGHT)):
                x0, s0 = solve(base, d, prec, **ekw)
                e = {"fma": compare(*solve(fma, d, prec, **ekw), x0, s0)}
                if pname == "SCHUR_JACOBI":
                    for v in (1, 2, 3):
                        e["variant%d" % v] = compare(*solve(base, d, prec, variant=v, **ekw), x0, s0)
                e["envelope_param_rel_diff"] = max(v["param_rel_diff"] for v in e.values())
                out["cases"]["%s/%s/%s" % (cname, pname, ename)] = e
                print(cname, pname, ename, "envelope %.2e" % e["envelope_param_rel_diff"], {k: "%.1e" % v["param_rel_diff"] for k, v in e.items() if isinstance(v, dict)})
    json.dump(out, open(os.path.join(HERE, "schur_jacobi_rounding_envelope.json"), "w"), indent=1)
