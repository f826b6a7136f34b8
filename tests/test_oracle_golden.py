"""Pins the CPU oracle against every golden vector / known answer the reference holds for the hot path
(SURVEY.md §8(c)).  No GPU needed.

  * AutodiffCostFuntionSpec.scala  — exact residuals + Jacobian layout at the evaluate boundary
  * RotationSpec.scala:616-655     — angleAxisRotatePoint vs rotation matrix, 10,000 seeded trials, 1e-9
  * CurveFitting.scala:22-90       — known answers of BASELINE.md §3 and the upstream tutorial log
  * independent numpy / scipy cross-checks of the restated Ceres linear algebra
"""
import json
import os

import numpy as np
import pytest

from skeres_b200 import _abi, synth

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    return json.load(open(os.path.join(HERE, "golden", name)))


def powell_problem(oracle, f2):
    """Powell.scala:56-76 / PowellAnalytic.scala: four residual blocks over the scalar blocks x1..x4 from (3, -1, 0, 1)."""
    p = oracle.OracleProblem(np.array([3.0, -1.0, 0.0, 1.0]))
    for fid, blocks in zip([_abi.FUNCTOR_POWELL_F1, f2, _abi.FUNCTOR_POWELL_F3, _abi.FUNCTOR_POWELL_F4], [(0, 1), (2, 3), (1, 2), (0, 3)]):
        p.add_residual_blocks(fid, np.zeros((1, 0)), np.array([blocks]))
    o = _abi.default_options()
    o.linear_solver_type, o.max_num_iterations = _abi.DENSE_QR, 100          # Powell.scala:80-81
    return p, o


def check_rows_against_log(rows, gold, only=None):
    sig = lambda v, digits: float(f"{v:.{digits}e}")
    assert len(rows) == len(gold["rows"])
    for row, g in zip(rows, gold["rows"]):
        if only is not None and g[0] not in only:
            continue
        assert row.iteration == g[0] and row.step_is_successful
        assert sig(row.cost, 6) == g[1] and sig(row.cost_change, 2) == g[2] and sig(row.gradient_max_norm, 2) == g[3]
        assert sig(row.step_norm, 2) == g[4] and sig(row.relative_decrease, 2) == g[5] and sig(row.trust_region_radius, 2) == g[6]


def test_hello_world_reproduces_ceres_tutorial_log(oracle):
    """HelloWorld.scala:11-35 (x: 0.5 -> 10): every figure of the published three-row log, recalled before the oracle was run."""
    gold = load("ceres_tutorial_helloworld_log.json")
    p = oracle.OracleProblem(np.array([0.5]))
    p.add_residual_blocks(_abi.FUNCTOR_HELLO_WORLD, np.zeros((1, 0)), np.array([[0]]))
    o = _abi.default_options()
    o.linear_solver_type = _abi.DENSE_QR        # the example keeps Ceres' default solver; one scalar block: any exact solver is the same
    s = p.solve(o)
    check_rows_against_log(s.iterations, gold)
    assert _abi.TERMINATION_NAMES[s.termination_type] == gold["final"]["termination"]
    assert abs(p.params[0] - gold["final"]["x"]) < 1e-7 and float(f"{s.final_cost:.6e}") == gold["final"]["final_cost"]


def test_powell_reproduces_ceres_tutorial_log(oracle):
    """PowellAnalytic.scala's problem (f2 = sqrt(5) (x3 - x4)) is the Ceres tutorial's: rows 0-4 and 14 and the final point of
    the published log were recalled before the oracle was run; all 15 rows are kept as a regression vector."""
    gold = load("ceres_tutorial_powell_log.json")
    p, o = powell_problem(oracle, _abi.FUNCTOR_POWELL_ANALYTIC_F2)
    s = p.solve(o)
    check_rows_against_log(s.iterations, gold, only=gold["recalled_rows"])
    check_rows_against_log(s.iterations, gold)
    assert _abi.TERMINATION_NAMES[s.termination_type] == gold["final"]["termination"] and "Gradient tolerance reached" in s.message
    assert [float(f"{v:.5e}") for v in p.params] == [float(f"{v:.5e}") for v in gold["final"]["x"]]
    assert float(f"{s.initial_cost:.6e}") == gold["final"]["initial_cost"] and float(f"{s.final_cost:.6e}") == gold["final"]["final_cost"]


def test_powell_as_the_reference_computes_it(oracle):
    """Powell.scala:28 evaluates sqrt(5) x3 - x4 (its comment says sqrt(5) (x3 - x4)): a different problem with the same
    minimiser 0.  Known answers: initial cost 105.5 (by hand: (49 + 1 + 1 + 160) / 2), all four parameters -> 0."""
    p, o = powell_problem(oracle, _abi.FUNCTOR_POWELL_F2)
    cost, r, g, J = p.evaluate()
    assert np.allclose(r, [3.0 - 10.0, np.sqrt(5.0) * 0.0 - 1.0, (-1.0 - 0.0) ** 2, np.sqrt(10.0) * (3.0 - 1.0) ** 2], rtol=1e-15)
    assert abs(cost - 0.5 * (49.0 + 1.0 + 1.0 + 160.0)) < 1e-12
    s = p.solve(o)
    assert s.termination_type == _abi.CONVERGENCE and s.final_cost < 1e-14 and np.all(np.abs(p.params) < 1e-3)


# ---------------------------------------------------------------------------------------------------
def test_autodiff_spec_vectors(oracle):
    for case in load("autodiff_spec_vectors.json")["cases"]:
        ok, res, jacs = oracle.evaluate(case["functor"], case["consts"], case["parameters"], want_jacobians=False)
        assert ok and jacs is None
        assert res.tolist() == case["residuals"], case["name"]           # `must be` == exact equality
        ok, res, jacs = oracle.evaluate(case["functor"], case["consts"], case["parameters"])
        assert ok and res.tolist() == case["residuals"]
        for j, want in zip(jacs, case["jacobians"]):
            assert j.ravel().tolist() == want, case["name"]             # row-major kNumResiduals x N_i


def test_null_jacobian_row_is_skipped(oracle):
    """AutodiffCostFunction.scala:118 — a NULL row means that block is held constant."""
    case = load("autodiff_spec_vectors.json")["cases"][1]
    ok, res, jacs = oracle.evaluate(case["functor"], case["consts"], case["parameters"], skip_blocks=(0,))
    assert ok and jacs[0] is None and jacs[1].ravel().tolist() == case["jacobians"][1]


class JavaRandom:
    """java.util.Random (scala.util.Random wraps it): 48-bit LCG, nextDouble = (next(26)<<27 + next(27)) * 2^-53."""

    def __init__(self, seed):
        self.seed = (seed ^ 0x5DEECE66D) & ((1 << 48) - 1)

    def _next(self, bits):
        self.seed = (self.seed * 0x5DEECE66D + 0xB) & ((1 << 48) - 1)
        return self.seed >> (48 - bits)

    def next_double(self):
        return ((self._next(26) << 27) + self._next(27)) * (1.0 / (1 << 53))


@pytest.mark.parametrize("theta_scale", [np.pi, 1.0e-16])
def test_rotate_point_matches_rotation_matrix(oracle, theta_scale):
    """RotationSpec.scala:616-635 (random angles) and :636-655 (near-zero angles): new Random(5), kNumTrials = 10000,
    kLooseTolerance = 1e-9 absolute."""
    rnd = JavaRandom(5)
    worst = 0.0
    for _ in range(10000):
        theta = theta_scale * (2 * rnd.next_double() - 1)
        axis = np.array([2 * rnd.next_double() - 1 for _ in range(3)])
        aa = theta * (axis / np.sqrt(axis @ axis))
        p = 10.0 * np.array([2 * rnd.next_double() - 1 for _ in range(3)])
        r1 = oracle.rotate_point(aa, p)
        r2 = oracle.rotation_matrix(aa) @ p
        worst = max(worst, np.max(np.abs(r1 - r2)))
    assert worst <= 1e-9


def test_snavely_jacobian_against_finite_differences(oracle):
    rng = np.random.default_rng(0)
    d = synth.make_bal("tiny", seed=2)
    for i in rng.integers(0, d.num_observations, 20):
        cam = d.parameters[9 * d.camera_index[i]:9 * d.camera_index[i] + 9]
        pt = d.parameters[9 * d.num_cameras + 3 * d.point_index[i]:][:3]
        obs = d.observations[2 * i:2 * i + 2]
        ok, res, (F, E) = oracle.evaluate(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, obs, [cam, pt])
        x = np.concatenate([cam, pt])
        J = np.concatenate([F, E], axis=1)
        for k in range(12):
            h = 1e-6 * max(1.0, abs(x[k]))
            xp, xm = x.copy(), x.copy()
            xp[k] += h; xm[k] -= h
            rp = oracle.evaluate(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, obs, [xp[:9], xp[9:]], want_jacobians=False)[1]
            rm = oracle.evaluate(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, obs, [xm[:9], xm[9:]], want_jacobians=False)[1]
            fd = (rp - rm) / (2 * h)
            assert np.allclose(J[:, k], fd, rtol=2e-5, atol=1e-5 * np.abs(J).max())


def test_taylor_branch_of_rotate_point(oracle):
    """Rotation.scala:492-520: ||aa||^2 <= ulp(1.0) uses R*pt = pt + aa x pt and still yields derivatives."""
    cam = np.array([1e-9, -2e-9, 3e-10, 0.1, -0.2, -10.0, 800.0, 1e-3, -1e-5])
    pt = np.array([0.3, -0.4, 1.5])
    ok, res, (F, E) = oracle.evaluate(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, [1.0, 2.0], [cam, pt])
    assert ok and np.all(np.isfinite(F)) and np.all(np.isfinite(E))
    assert np.abs(F[:, :3]).max() > 1.0      # rotation derivatives are not lost in the small-angle branch


# ---------------------------------------------------------------------------------------------------
def curve_fitting_problem(oracle):
    d = load("curve_fitting_data.json")
    p = oracle.OracleProblem(np.zeros(2))
    p.add_residual_blocks(_abi.FUNCTOR_EXPONENTIAL_RESIDUAL, np.stack([d["x"], d["y"]], 1), np.tile([0, 1], (67, 1)))
    return p, np.array(d["x"]), np.array(d["y"])


def test_curve_fitting_known_answers(oracle):
    """initial cost 1.211734e+02, initial max-gradient 3.61e+02, optimum 1.056751e+00 at (0.291871, 0.131401)."""
    p, x, y = curve_fitting_problem(oracle)
    cost, r, g, J = p.evaluate()
    assert abs(cost - 1.211734e+02) < 5e-5
    assert abs(np.abs(g).max() - 3.61e+02) < 0.5
    from scipy.optimize import least_squares
    sol = least_squares(lambda t: y - np.exp(t[0] * x + t[1]), [0.0, 0.0], xtol=1e-15, ftol=1e-15, gtol=1e-15)
    assert abs(sol.cost - 1.056751) < 1e-6 and np.allclose(sol.x, [0.291871, 0.131401], atol=2e-6)


def test_curve_fitting_reproduces_ceres_tutorial_log(oracle):
    """The LM / trust-region / Jacobi-scaling / QR restatement reproduces the published Ceres log row by row."""
    gold = load("ceres_tutorial_curve_fitting_log.json")
    p, _, _ = curve_fitting_problem(oracle)
    o = _abi.default_options()
    o.linear_solver_type = _abi.DENSE_QR
    o.max_num_iterations = 25                                     # CurveFitting.scala:120
    s = p.solve(o)
    assert _abi.TERMINATION_NAMES[s.termination_type] == gold["final"]["termination"]
    assert len(s.iterations) == len(gold["rows"]) == 14
    sig = lambda v, digits: float(f"{v:.{digits}e}")
    for row, g in zip(s.iterations, gold["rows"]):
        assert row.iteration == g[0]
        if row.step_is_successful:
            assert sig(row.cost, 6) == g[1]
        assert sig(row.cost_change, 2) == g[2]
        assert sig(row.gradient_max_norm, 2) == g[3]
        assert sig(row.step_norm, 2) == g[4]
        assert sig(row.relative_decrease, 2) == g[5]
        assert sig(row.trust_region_radius, 2) == g[6]
    assert round(p.params[0], 6) == gold["final"]["m"] and round(p.params[1], 6) == gold["final"]["c"]
    assert sig(s.initial_cost, 6) == gold["final"]["initial_cost"] and sig(s.final_cost, 6) == gold["final"]["final_cost"]
    assert s.num_successful_steps == 9 and s.num_unsuccessful_steps == 5


# ---------------------------------------------------------------------------------------------------
def ba_problem(oracle, shape="tiny", seed=1, loss=(_abi.LOSS_TRIVIAL, 0.0)):
    d = synth.make_bal(shape, seed=seed)
    p = oracle.OracleProblem(d.parameters)
    p.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets(), *loss)
    return d, p


def dense_lm_step(d, p, radius=1e4):
    """Independent numpy restatement of one LM step: Jacobi scaling, LM diagonal, normal equations."""
    cost, r, g, Jv = p.evaluate()
    n = d.parameters.size
    J = np.zeros((2 * d.num_observations, n))
    off = d.block_offsets()
    Jv = Jv.reshape(-1, 24)
    for i in range(d.num_observations):
        J[2 * i:2 * i + 2, off[i, 0]:off[i, 0] + 9] = Jv[i, :18].reshape(2, 9)
        J[2 * i:2 * i + 2, off[i, 1]:off[i, 1] + 3] = Jv[i, 18:].reshape(2, 3)
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))
    Js = J * scale
    diag = np.clip((Js * Js).sum(0), 1e-6, 1e32)
    H = Js.T @ Js + np.diag(diag / radius)
    y = np.linalg.solve(H, Js.T @ r)
    return cost, g, J, -y * scale


def test_gradient_and_cost_against_numpy(oracle):
    d, p = ba_problem(oracle)
    cost, g, J, _ = dense_lm_step(d, p)
    c2, r, g2, _ = p.evaluate()
    assert np.isclose(0.5 * r @ r, c2, rtol=1e-14)
    assert np.allclose(J.T @ r, g2, rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("lst", [_abi.DENSE_SCHUR, _abi.SPARSE_SCHUR])
def test_first_schur_step_equals_dense_normal_equations(oracle, lst):
    """SchurEliminator + Cholesky + back-substitution == the full (J'J + D'D) solve (A.5)."""
    d, p = ba_problem(oracle)
    _, _, _, delta = dense_lm_step(d, p)
    x0 = p.params.copy()
    o = _abi.default_options()
    o.linear_solver_type = lst
    o.max_num_iterations = 1
    s = p.solve(o)
    assert s.iterations[1].step_is_successful
    assert np.allclose(p.params - x0, delta, rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("shape,seed", [("tiny", 1), ("small", 2), ("small", 5)])
@pytest.mark.parametrize("linear,lst,prec", [("dense", _abi.DENSE_SCHUR, _abi.JACOBI), ("pcg", _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI),
                                             ("pcg", _abi.ITERATIVE_SCHUR, _abi.JACOBI)])
def test_whole_trajectories_against_an_independent_numpy_solver(oracle, shape, seed, linear, lst, prec):
    """not one LM step but whole solves.  tests/numpy_lm.py states the trust-region loop, the LM strategy,
    the explicit Schur complement and Ceres' conjugate-gradients solver with the SchurJacobi / Jacobi preconditioner in dense
    numpy -- sharing only the functor evaluation with the oracle -- and the oracle's C++ restatement (Schur eliminator, implicit
    products, fixed-order sums) must follow it row for row: costs to 1e-11, the same radii, the same accepted / rejected steps
    and the same number of PCG iterations in every linear solve."""
    import numpy_lm
    d = synth.make_bal(shape, seed=seed)

    def evaluate(x):
        q = oracle.OracleProblem(x)
        q.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets())
        c, r, _, Jv = q.evaluate()
        return c, r, numpy_lm.dense_jacobian(d, Jv)
    rows, x = numpy_lm.solve(d, evaluate, linear=linear, preconditioner="schur_jacobi" if prec == _abi.SCHUR_JACOBI else "jacobi")
    p = oracle.OracleProblem(d.parameters)
    p.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets())
    o = _abi.default_options()
    o.linear_solver_type, o.preconditioner_type = lst, prec
    s = p.solve(o)
    assert s.termination_type == _abi.CONVERGENCE and len(rows) == len(s.iterations) >= 3
    for a, b in zip(rows, s.iterations):
        assert np.isclose(a[0], b.cost, rtol=1e-11) and np.isclose(a[1], b.trust_region_radius, rtol=1e-9)
        assert bool(a[2]) == bool(b.step_is_successful) and a[3] == b.linear_solver_iterations
    assert np.isclose(evaluate(x)[0], evaluate(p.params)[0], rtol=1e-11)       # neither side applies the step that met the tolerance


def test_iterative_schur_converges_to_the_exact_schur_optimum(oracle):
    costs = {}
    for lst, prec in [(_abi.DENSE_SCHUR, _abi.JACOBI), (_abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI), (_abi.ITERATIVE_SCHUR, _abi.JACOBI),
                      (_abi.ITERATIVE_SCHUR, _abi.IDENTITY)]:
        d, p = ba_problem(oracle, "small")
        o = _abi.default_options()
        o.linear_solver_type, o.preconditioner_type = lst, prec
        o.function_tolerance = 1e-12
        s = p.solve(o)
        assert s.termination_type == _abi.CONVERGENCE
        costs[(lst, prec)] = s.final_cost
    ref = costs[(_abi.DENSE_SCHUR, _abi.JACOBI)]
    for v in costs.values():
        assert abs(v - ref) <= 1e-8 * ref



def _snavely_residuals_scipy(x, d):
    """Independent of the oracle and of the device code: SciPy's rotation-vector class for R(angle-axis), numpy for the rest
    (SimpleBundleAdjuster.scala:79-119 as a formula: p = R X + t, projection through -z, two radial terms)."""
    from scipy.spatial.transform import Rotation
    nc = d.num_cameras
    cams = x[:9 * nc].reshape(nc, 9)[d.camera_index]
    pts = x[9 * nc:].reshape(-1, 3)[d.point_index]
    p = Rotation.from_rotvec(cams[:, :3]).apply(pts) + cams[:, 3:6]
    xp, yp = -p[:, 0] / p[:, 2], -p[:, 1] / p[:, 2]
    r2 = xp * xp + yp * yp
    f = cams[:, 6] * (1.0 + r2 * (cams[:, 7] + cams[:, 8] * r2))
    return (np.stack([f * xp, f * yp], 1) - d.observations.reshape(-1, 2)).ravel()


@pytest.mark.parametrize("seed", [1, 4])
def test_ba_residuals_and_optimum_against_scipy(oracle, seed):
    """an independent cross-check of bundle-adjustment OPTIMA (not trajectories).  The residual vector of the
    oracle equals SciPy's own rotation-vector arithmetic to 1e-11 px, and the cost the oracle's trust-region / Schur solve
    converges to equals -- to 1e-7, from below -- the cost scipy.optimize.least_squares (trust-region reflective, finite-
    difference Jacobian: none of the restated Ceres algebra) reaches from the same start."""
    from scipy.optimize import least_squares
    from scipy.sparse import lil_matrix
    d, p = ba_problem(oracle, "tiny", seed)
    cost, r, g, Jv = p.evaluate()
    r0 = _snavely_residuals_scipy(d.parameters, d)
    assert np.abs(r0 - r).max() < 1e-11 and abs(0.5 * r0 @ r0 - cost) <= 1e-13 * cost
    o = _abi.default_options()
    o.linear_solver_type, o.max_num_iterations = _abi.DENSE_SCHUR, 200
    o.function_tolerance = o.gradient_tolerance = o.parameter_tolerance = 1e-14
    s = p.solve(o)
    assert s.termination_type == _abi.CONVERGENCE
    n_obs = d.num_observations
    A = lil_matrix((2 * n_obs, d.parameters.size), dtype=int)
    i = np.arange(n_obs)
    for k in range(9):
        A[2 * i, 9 * d.camera_index + k] = 1; A[2 * i + 1, 9 * d.camera_index + k] = 1
    for k in range(3):
        A[2 * i, 9 * d.num_cameras + 3 * d.point_index + k] = 1; A[2 * i + 1, 9 * d.num_cameras + 3 * d.point_index + k] = 1
    sol = least_squares(_snavely_residuals_scipy, d.parameters, args=(d,), jac_sparsity=A, method="trf", x_scale="jac",
                        xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=200)
    assert abs(sol.cost - s.final_cost) <= 1e-7 * s.final_cost
    assert s.final_cost <= sol.cost * (1 + 1e-12)          # scipy, limited by its finite differences, stops just above


def test_tight_pcg_matches_exact_schur_step(oracle):
    """With eta -> 0 the implicit-Schur PCG solves the same reduced system the eliminator factorises (A.6/A.7)."""
    d, p1 = ba_problem(oracle)
    d, p2 = ba_problem(oracle)
    o = _abi.default_options()
    o.max_num_iterations = 1
    o.linear_solver_type = _abi.DENSE_SCHUR
    p1.solve(o)
    o.linear_solver_type, o.preconditioner_type, o.eta = _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI, 1e-14
    p2.solve(o)
    assert np.allclose(p1.params, p2.params, rtol=1e-5, atol=1e-6)


def test_unsorted_input_gives_the_same_solve(oracle):
    d = synth.make_bal("tiny", seed=4)
    perm = np.random.default_rng(0).permutation(d.num_observations)
    outs = []
    for order in (np.arange(d.num_observations), perm):
        p = oracle.OracleProblem(d.parameters)
        p.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2)[order], d.block_offsets()[order])
        o = _abi.default_options()
        o.linear_solver_type = _abi.DENSE_SCHUR
        s = p.solve(o)
        outs.append((len(s.iterations), s.final_cost, p.params.copy()))
    assert outs[0][0] == outs[1][0] and np.isclose(outs[0][1], outs[1][1], rtol=1e-10)
    assert np.allclose(outs[0][2], outs[1][2], rtol=1e-6, atol=1e-8)


def test_robust_curve_fitting_reproduces_the_ceres_tutorial_answer(oracle):
    """RobustCurveFitting.scala as written: its own 67 pairs with the two outliers (:41-42), CauchyLoss(0.5) (:107), 25
    iterations, DENSE_QR, start (0, 0).  The reference example is the Ceres tutorial's robust_curve_fitting; the published
    tutorial quotes the fit m = 0.287605, c = 0.151213 with the loss (and a visibly worse one without).  The two numbers
    were written down from memory of the published docs BEFORE the oracle was run on this data; the oracle's Corrector +
    Cauchy path reproduces them to all six digits."""
    d = load("robust_curve_fitting_data.json")
    assert d["outlier_indices"] == [19, 20] and len(d["x"]) == 67
    p = oracle.OracleProblem(np.zeros(2))
    p.add_residual_blocks(_abi.FUNCTOR_EXPONENTIAL_RESIDUAL, np.stack([d["x"], d["y"]], 1), np.tile([0, 1], (67, 1)), _abi.LOSS_CAUCHY, d["cauchy_a"])
    o = _abi.default_options()
    o.linear_solver_type, o.max_num_iterations = _abi.DENSE_QR, d["max_num_iterations"]
    s = p.solve(o)
    assert s.termination_type == _abi.CONVERGENCE
    assert [float(f"{v:.6f}") for v in p.params] == [0.287605, 0.151213]
    q = oracle.OracleProblem(np.zeros(2))                                  # without the loss the outliers drag the fit away
    q.add_residual_blocks(_abi.FUNCTOR_EXPONENTIAL_RESIDUAL, np.stack([d["x"], d["y"]], 1), np.tile([0, 1], (67, 1)))
    q.solve(o)
    assert abs(q.params[0] - 0.3) > 2 * abs(p.params[0] - 0.3) and abs(q.params[1] - 0.1) > 2 * abs(p.params[1] - 0.1)


def test_schur_jacobi_parameters_are_rounding_sensitive_at_the_default_eta(oracle):
    """DESIGN.md section 2, parity gap 1.  An algebraically identical re-ordering of the oracle's own Schur eliminator
    (G^T (E^-1 G) instead of (G^T E^-1) G) moves the parameters a truncated (eta = 0.1) ITERATIVE_SCHUR + SCHUR_JACOBI
    solve of the Ladybug-49 shape ends at by ~6e-4 relative -- same LM rows, same PCG iteration counts, costs equal to
    1e-9 -- so no implementation can meet a 1e-5 parameter tolerance against the oracle there without being bit-identical
    to it.  With converged linear solves (eta = 1e-10) the same perturbation stays below 1e-6.  The measured values are
    committed (tests/golden/schur_jacobi_rounding_envelope.json, made by make_rounding_envelope.py); the GPU parity tests
    hold the device to the 1e-5 bar on the converged solves and to a multiple of this envelope on the truncated ones."""
    import ctypes as C
    env = load("schur_jacobi_rounding_envelope.json")["cases"]
    L = oracle.lib()
    L.oracle_set_schur_rounding_variant.argtypes = [C.c_int]
    d = synth.make_bal("ladybug-49", seed=1)
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-2)))

    def solve(variant, **kw):
        L.oracle_set_schur_rounding_variant(variant)
        try:
            p = oracle.OracleProblem(d.parameters)
            p.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets())
            o = _abi.default_options()
            o.linear_solver_type, o.preconditioner_type = _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI
            for k, v in kw.items():
                setattr(o, k, v)
            summ = p.solve(o)
            return p.params.copy(), summ
        finally:
            L.oracle_set_schur_rounding_variant(0)

    x0, s0 = solve(0)
    x1, s1 = solve(1)
    assert [r.linear_solver_iterations for r in s0.iterations] == [r.linear_solver_iterations for r in s1.iterations]
    assert abs(s0.final_cost - s1.final_cost) <= 1e-8 * s0.final_cost
    gap = rel(x1, x0)
    assert gap > 10 * 1e-5, gap                                            # an order of magnitude above the north star's tolerance
    assert np.isclose(gap, env["ladybug-49/SCHUR_JACOBI/eta0.1"]["variant1"]["param_rel_diff"], rtol=1e-3)   # the committed fixture is this run
    tight = dict(eta=1e-10, max_linear_solver_iterations=3000, max_num_iterations=3)
    xt0, _ = solve(0, **tight)
    xt1, _ = solve(1, **tight)
    assert rel(xt1, xt0) < 1e-5
    # every committed envelope: truncated SCHUR_JACOBI / IDENTITY far above 1e-5 on the larger cases, JACOBI far below
    assert env["ladybug-49/SCHUR_JACOBI/eta0.1"]["envelope_param_rel_diff"] > 1e-4 > 1e-5 > env["ladybug-49/SCHUR_JACOBI/eta1e-10"]["envelope_param_rel_diff"]
    assert env["ladybug-49/JACOBI/eta0.1"]["envelope_param_rel_diff"] < 1e-9
    assert env["long-tracks/SCHUR_JACOBI/eta0.1"]["fma"]["param_rel_diff"] > 1e-2 and env["long-tracks/JACOBI/eta0.1"]["envelope_param_rel_diff"] < 1e-8


# ---------------------------------------------------------------------------------------------------
def test_loss_functions(oracle):
    """ceres/loss_function.h closed forms (A.2)."""
    for s in [0.0, 0.1, 1.0, 7.5]:
        assert oracle.loss(_abi.LOSS_TRIVIAL, 0, s).tolist() == [s, 1.0, 0.0]
        a = 0.5
        rho = oracle.loss(_abi.LOSS_CAUCHY, a, s)
        b = a * a
        assert np.allclose(rho, [b * np.log1p(s / b), 1 / (1 + s / b), -(1 / b) / (1 + s / b) ** 2], rtol=1e-14)
        rho = oracle.loss(_abi.LOSS_HUBER, a, s)
        want = [s, 1.0, 0.0] if s <= b else [2 * a * np.sqrt(s) - b, a / np.sqrt(s), -a / np.sqrt(s) / (2 * s)]
        assert np.allclose(rho, want, rtol=1e-14)


def test_robust_loss_reduces_outlier_influence(oracle):
    """RobustCurveFitting.scala:41-42,107: two outliers + CauchyLoss(0.5) stay near the clean fit."""
    d = load("curve_fitting_data.json")
    x, y = np.array(d["x"]), np.array(d["y"]).copy()
    y[10] += 8.0; y[40] -= 6.0
    fits = {}
    for name, loss in [("trivial", (_abi.LOSS_TRIVIAL, 0.0)), ("cauchy", (_abi.LOSS_CAUCHY, 0.5))]:
        p = oracle.OracleProblem(np.zeros(2))
        p.add_residual_blocks(_abi.FUNCTOR_EXPONENTIAL_RESIDUAL, np.stack([x, y], 1), np.tile([0, 1], (67, 1)), *loss)
        o = _abi.default_options()
        o.linear_solver_type = _abi.DENSE_QR
        s = p.solve(o)
        assert s.termination_type == _abi.CONVERGENCE
        fits[name] = p.params.copy()
    clean = np.array([0.291861, 0.131439])
    assert np.linalg.norm(fits["cauchy"] - clean) < 0.5 * np.linalg.norm(fits["trivial"] - clean)
