"""Parity of the CUDA path (through the C ABI of libskeres.so) with the CPU oracle.  Needs a B200.

Tolerances (BASELINE.json north_star): final cost within 1e-6 relative, parameters within 1e-5 relative,
identical iteration / termination behaviour on well-conditioned problems.  Integer quantities (iteration
counts, PCG iteration counts, termination types) are compared exactly.
"""
import json
import os

import numpy as np
import pytest

from skeres_b200 import _abi, synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
COST_RTOL = 1e-6
PARAM_RTOL = 1e-5


def load(name):
    return json.load(open(os.path.join(HERE, "golden", name)))


def rel_param_diff(a, b, floor=1e-2):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


# Truncated (eta = 0.1) PCG with SCHUR_JACOBI / IDENTITY: the end point is not a well-conditioned function of the input
# (tests/golden/make_rounding_envelope.py: the oracle moves by `envelope` against its OWN -ffp-contract=fast build and against
# algebraically identical re-orderings of its eliminator; tests/test_oracle_golden.py pins that on the CPU).  Those perturb a
# handful of operations by one rounding each; the device re-associates every sum (tiles, segments, trees), a seed one to two
# orders of magnitude larger for the same amplification, hence the factor.  The 1e-5 bar itself is asserted on the converged
# solves (test_ba_converged_linear_solves_meet_the_parameter_tolerance), on JACOBI and on the exact Schur solvers.
ENVELOPE_FACTOR = 100.0


def envelope_bound(case):
    e = load("schur_jacobi_rounding_envelope.json")["cases"][case]
    return max(PARAM_RTOL, ENVELOPE_FACTOR * e["fma"]["param_rel_diff"])


def oracle_ba(oracle, d, lst, prec=_abi.SCHUR_JACOBI, loss=(_abi.LOSS_TRIVIAL, 0.0), **opts):
    p = oracle.OracleProblem(d.parameters)
    p.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets(), *loss)
    o = _abi.default_options()
    o.linear_solver_type, o.preconditioner_type = lst, prec
    for k, v in opts.items():
        setattr(o, k, v)
    return p, p.solve(o)


def gpu_ba(sk, d, lst, prec=_abi.SCHUR_JACOBI, loss=None, order=None, functor_id=_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, **opts):
    bal = sk.BalProblem.fromArrays(d)
    if order is not None:
        bal.cameraIndex, bal.pointIndex = bal.cameraIndex[order], bal.pointIndex[order]
        bal.observations = bal.observations.reshape(-1, 2)[order].ravel()
    problem = bal.buildProblem(loss, functor_id)
    o = sk.Solver.Options()
    o.setLinearSolverType(lst)
    o.setPreconditionerType(prec)
    for k, v in opts.items():
        setattr(o, k, v)
    s = sk.Solver.Summary()
    sk.ceres.solve(o, problem, s)
    return bal, s


def assert_same_trajectory(s, so, exact_rows=True, row_rtol=1e-6):
    assert s.termination_type == so.termination_type, (s.message, so.message)
    assert len(s.iterations) == len(so.iterations)
    assert (s.num_successful_steps, s.num_unsuccessful_steps) == (so.num_successful_steps, so.num_unsuccessful_steps)
    for a, b in zip(s.iterations, so.iterations):
        assert (a.iteration, a.step_is_valid, a.step_is_successful) == (b.iteration, b.step_is_valid, b.step_is_successful)
        if exact_rows:
            assert a.linear_solver_iterations == b.linear_solver_iterations
        assert np.isclose(a.cost, b.cost, rtol=row_rtol)
        assert np.isclose(a.trust_region_radius, b.trust_region_radius, rtol=1e-4)
    assert abs(s.initial_cost - so.initial_cost) <= 1e-12 * abs(so.initial_cost)
    assert abs(s.final_cost - so.final_cost) <= COST_RTOL * abs(so.final_cost)


# --------------------------------------------------------------------------------------------------- evaluate boundary
def test_golden_vectors_host_abi(sk):
    """AutodiffCostFuntionSpec.scala replayed through sk_cost_function_evaluate_host (exact equality)."""
    for case in load("autodiff_spec_vectors.json")["cases"]:
        cf = sk.CostFunction(case["functor"], case["consts"])
        ok, res, jacs = cf.evaluate_host(case["parameters"], want_jacobians=False)
        assert ok and res.tolist() == case["residuals"] and jacs is None          # jacobians == NULL branch (:80)
        ok, res, jacs = cf.evaluate_host(case["parameters"])
        assert ok and res.tolist() == case["residuals"]
        for j, want in zip(jacs, case["jacobians"]):
            assert j.ravel().tolist() == want
        ok, res, jacs = cf.evaluate_host(case["parameters"], skip_blocks=(0,))     # NULL row (:118)
        assert ok and jacs[0] is None and jacs[1].ravel().tolist() == case["jacobians"][1]


def test_user_functors_from_source(sk):
    """the three AutodiffCostFuntionSpec functors supplied AS SOURCE STRINGS (sk_functor_register_source:
    NVRTC over the device Jet<N>) reproduce the reference's golden vectors exactly, through the same NULL-jacobians and
    NULL-row branches as the built-in registrations."""
    import user_functor_sources as U
    for (name, src, nres, sizes, nconsts), case in zip(U.SPEC, load("autodiff_spec_vectors.json")["cases"]):
        F = sk.SourceCostFunctor.define(name, src, nres, sizes, nconsts)
        cf = F(*case["consts"]).toAutoDiffCostFunction()
        assert cf.functor_id >= 1000 and cf.kNumResiduals == nres and cf.N == sizes
        ok, res, jacs = cf.evaluate_host(case["parameters"], want_jacobians=False)
        assert ok and res.tolist() == case["residuals"] and jacs is None
        ok, res, jacs = cf.evaluate_host(case["parameters"])
        assert ok and res.tolist() == case["residuals"]
        for j, want in zip(jacs, case["jacobians"]):
            assert j.ravel().tolist() == want
        ok, res, jacs = cf.evaluate_host(case["parameters"], skip_blocks=(0,))
        assert ok and jacs[0] is None and jacs[1].ravel().tolist() == case["jacobians"][1]


def test_user_functor_solves_curve_fitting_like_the_builtin(sk):
    """CurveFitting.scala with ExponentialResidual given as source: every LM row equals the built-in functor's run (the
    same arithmetic on the same duals), including the published final (0.291861, 0.131439); and a problem that MIXES
    built-in and run-time functors gives the same rows again."""
    import user_functor_sources as U
    d = load("curve_fitting_data.json")
    F = sk.SourceCostFunctor.define("UserExponentialResidual", U.EXPONENTIAL, 1, [1, 1], 2)

    def fit(make):
        m, c = sk.DoubleArray(1), sk.DoubleArray(1)
        problem = sk.Problem()
        for i, (xi, yi) in enumerate(zip(d["x"], d["y"])):
            problem.addResidualBlock(make(i, xi, yi).toAutoDiffCostFunction(), sk.PredefinedLossFunctions.trivialLoss(), m.toPointer(), c.toPointer())
        o = sk.Solver.Options()
        o.setMaxNumIterations(25); o.setLinearSolverType(_abi.DENSE_QR)
        s = sk.Solver.Summary()
        sk.ceres.solve(o, problem, s)
        return [m.get(0), c.get(0)], s
    x0, s0 = fit(lambda i, x, y: sk.ExponentialResidual(x, y))
    x1, s1 = fit(lambda i, x, y: F(x, y))
    x2, s2 = fit(lambda i, x, y: F(x, y) if i % 2 else sk.ExponentialResidual(x, y))
    for xs, s in ((x1, s1), (x2, s2)):
        assert len(s.iterations) == len(s0.iterations) == 14 and s.termination_type == s0.termination_type
        for a, b in zip(s.iterations, s0.iterations):
            assert (a.step_is_successful, a.linear_solver_iterations) == (b.step_is_successful, b.linear_solver_iterations)
            assert np.isclose(a.cost, b.cost, rtol=1e-13) and np.isclose(a.trust_region_radius, b.trust_region_radius, rtol=1e-12)
        assert np.allclose(xs, x0, rtol=1e-12)
    assert [float(f"{v:.6f}") for v in x1] == [0.291861, 0.131439]


@pytest.mark.parametrize("shape,seed,lst,prec,loss", [("small", 2, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI, None),
                                                      ("small", 5, _abi.ITERATIVE_SCHUR, _abi.JACOBI, ("huber", 1.0)),
                                                      ("tiny", 1, _abi.DENSE_SCHUR, _abi.JACOBI, None),
                                                      ("ladybug-49", 1, _abi.SPARSE_SCHUR, _abi.JACOBI, None)])
def test_user_ba_functor_runs_in_the_tile_kernels(sk, oracle, shape, seed, lst, prec, loss):
    """SURVEY 8(f) rank 4 on the hot path: SnavelyReprojectionError ported from SimpleBundleAdjuster.scala:79-119 AS A SOURCE
    STRING (sk_functor_register_source) is compiled by NVRTC into the tile evaluation kernel of the Schur solvers
    (ba_evaluate.cuh: the body the built-in functor runs in) and solves bundle adjustment like the built-in registration --
    same LM rows and PCG iteration counts, costs to 1e-9 (a full-width Jet<12> against the built-in's two 6-wide stages: the
    same derivatives, different roundings) -- and like the oracle."""
    import user_functor_sources as U
    F = sk.SourceCostFunctor.define(*U.BA_SHAPED[0])
    d = synth.make_bal(shape, seed=seed)
    mk = (lambda: getattr(sk.PredefinedLossFunctions, loss[0] + "Loss")(loss[1])) if loss else (lambda: None)
    bal0, s0 = gpu_ba(sk, d, lst, prec, loss=mk())
    bal1, s1 = gpu_ba(sk, d, lst, prec, loss=mk(), functor_id=F.functor_id)
    assert_same_trajectory(s1, s0, row_rtol=1e-9)
    assert rel_param_diff(bal1.parameters.toArray(), bal0.parameters.toArray(), 1e-2) <= 1e-6
    assert s1.num_kernel_launches == s0.num_kernel_launches
    if loss is None:
        p, so = oracle_ba(oracle, d, lst, prec)
        assert_same_trajectory(s1, so, row_rtol=1e-6)


def test_unseen_camera_model_on_the_tile_path(sk, oracle):
    """A camera model the library has never seen (division model of radial distortion, same (2; 9, 3) shape, reports failure
    for a point behind the camera), given as source: the Schur solvers run it in the tile kernels, DENSE_QR runs it on the
    dense back end (per-block addResidualBlock, AutoDiffCostFunctor.toAutoDiffCostFunction).  Checked against the oracle,
    which holds the same functor as a checker-only registration (oracle/jet.h: divisionModelReprojectionError, id 900): exact
    solvers row for row at 1e-9, ITERATIVE_SCHUR like the built-in functor's runs (same rows, same PCG counts); the two
    exact device paths agree with each other.  A functor failure at the initial point is a FAILURE of the solve."""
    import user_functor_sources as U
    F = sk.SourceCostFunctor.define(*U.BA_SHAPED[1])
    d = synth.make_bal("tiny", seed=4)

    def oracle_solve(lst, prec=_abi.JACOBI):
        p = oracle.OracleProblem(d.parameters)
        p.add_residual_blocks(900, d.observations.reshape(-1, 2), d.block_offsets())
        o = _abi.default_options()
        o.linear_solver_type, o.preconditioner_type, o.max_num_iterations = lst, prec, 8
        return p, p.solve(o)
    bal1, s1 = gpu_ba(sk, d, _abi.DENSE_SCHUR, functor_id=F.functor_id, max_num_iterations=8)
    bal2, s2 = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI, functor_id=F.functor_id, max_num_iterations=8)
    bal3 = sk.BalProblem.fromArrays(d)
    problem = sk.Problem()
    o = d.observations
    for i in range(d.num_observations):
        problem.addResidualBlock(F(o[2 * i], o[2 * i + 1]).toAutoDiffCostFunction(), None, bal3.mutableCameraForObservation(i), bal3.mutablePointForObservation(i))
    opt = sk.Solver.Options()
    opt.setLinearSolverType(_abi.DENSE_QR); opt.setMaxNumIterations(8)
    s3 = sk.Solver.Summary()
    sk.ceres.solve(opt, problem, s3)
    assert len(s1.iterations) >= 3 and s1.final_cost < 0.1 * s1.initial_cost
    p1, so1 = oracle_solve(_abi.DENSE_SCHUR)
    p2, so2 = oracle_solve(_abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI)
    p3, so3 = oracle_solve(_abi.DENSE_QR)
    assert_same_trajectory(s1, so1, row_rtol=1e-9)
    assert_same_trajectory(s2, so2, row_rtol=1e-6)
    assert_same_trajectory(s3, so3, row_rtol=1e-9)
    assert rel_param_diff(bal1.parameters.toArray(), p1.params, 1e-2) <= PARAM_RTOL
    assert rel_param_diff(bal3.parameters.toArray(), p3.params, 1e-2) <= PARAM_RTOL
    for a, b in zip(s1.iterations, s3.iterations):                       # tile path == dense path (two exact linear solvers)
        assert np.isclose(a.cost, b.cost, rtol=1e-9)
    # the built-in model on the same data is a different problem: the functor in the tile kernel really is the user's
    bal0, s0 = gpu_ba(sk, d, _abi.DENSE_SCHUR, max_num_iterations=8)
    assert abs(s0.iterations[0].cost - s1.iterations[0].cost) > 1e-6 * s1.iterations[0].cost
    # functor failure: camera 0 turned around (t -> -t puts everything it sees behind it)
    bad = synth.make_bal("tiny", seed=4)
    bad.parameters[3:6] *= -1.0
    balf, sf = gpu_ba(sk, bad, _abi.DENSE_SCHUR, functor_id=F.functor_id, max_num_iterations=3)
    assert sf.termination_type == _abi.FAILURE


def test_user_functor_is_rejected_by_the_schur_solvers(sk):
    """... unless it has the bundle-adjustment shape (the tests above)."""
    import user_functor_sources as U
    F = sk.SourceCostFunctor.define("UserExponentialResidual", U.EXPONENTIAL, 1, [1, 1], 2)
    m, c = sk.DoubleArray(1), sk.DoubleArray(1)
    problem = sk.Problem()
    problem.addResidualBlock(F(1.0, 2.0).toAutoDiffCostFunction(), None, m.toPointer(), c.toPointer())
    o = sk.Solver.Options()
    o.setLinearSolverType(_abi.ITERATIVE_SCHUR)
    with pytest.raises(sk.SkeresError) as e:
        sk.ceres.solve(o, problem, sk.Solver.Summary())
    assert e.value.status == _abi.ERR_UNSUPPORTED


def test_golden_vectors_device_pointers(sk):
    """The same through device DoubleArrays / DoublePointers, as the Scala spec does with RichDoubleMatrix."""
    case = load("autodiff_spec_vectors.json")["cases"][1]
    params = sk.DoubleArray.fromArray(np.concatenate(case["parameters"]))
    residuals = sk.DoubleArray(3)
    jac = sk.DoubleArray(12)
    cf = sk.CostFunction(case["functor"], case["consts"])
    pp = [params.toPointer(), params.slice(2)]
    assert cf.evaluate(pp, residuals.toPointer(), None) is True
    assert residuals.toArray().tolist() == case["residuals"]
    residuals.copyFrom(np.zeros(3))
    assert cf.evaluate(pp, residuals.toPointer(), [jac.toPointer(), jac.slice(6)]) is True
    assert residuals.get(0) == 10.0
    assert jac.toArray().tolist() == case["jacobians"][0] + case["jacobians"][1]


def test_double_array_semantics(sk):
    """RichDoubleArraySpec.scala / DoubleArraySliceSpec.scala: get/set/slice alias the parent buffer."""
    a = sk.DoubleArray(10)
    a.copyFrom(np.arange(10.0))
    s = a.slice(4)
    assert s.get(0) == 4.0 and s.toArray(3).tolist() == [4.0, 5.0, 6.0]
    s.set(1, -1.0)
    assert a.get(5) == -1.0
    assert s.slice(2).get(0) == 6.0
    with pytest.raises(sk.SkeresError):
        a.copyFrom(np.zeros(11))
    with pytest.raises(sk.SkeresError):
        a.get(10)


def test_snavely_evaluate_matches_oracle(sk, oracle):
    d = synth.make_bal("small", seed=11)
    rng = np.random.default_rng(0)
    idx = list(rng.integers(0, d.num_observations, 60)) + list(np.nonzero(d.camera_index == 0)[0][:10])   # camera 0: Taylor branch
    for i in idx:
        cam = d.parameters[9 * d.camera_index[i]:9 * d.camera_index[i] + 9]
        pt = d.parameters[9 * d.num_cameras + 3 * d.point_index[i]:][:3]
        obs = d.observations[2 * i:2 * i + 2]
        cf = sk.CostFunction(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, obs)
        ok, res, (F, E) = cf.evaluate_host([cam, pt])
        ok2, res_o, (Fo, Eo) = oracle.evaluate(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, obs, [cam, pt])
        assert ok and ok2
        scale = max(np.abs(Fo).max(), np.abs(Eo).max())
        assert np.allclose(res, res_o, rtol=1e-12, atol=1e-9)
        assert np.max(np.abs(F - Fo)) <= 1e-12 * scale and np.max(np.abs(E - Eo)) <= 1e-12 * scale
        ok, res_only, _ = cf.evaluate_host([cam, pt], want_jacobians=False)
        assert np.allclose(res_only, res_o, rtol=1e-12, atol=1e-9)


def test_loss_functions_match_oracle(sk, oracle):
    L = sk.PredefinedLossFunctions
    for loss, kind, a in [(L.trivialLoss(), _abi.LOSS_TRIVIAL, 0.0), (L.huberLoss(0.7), _abi.LOSS_HUBER, 0.7), (L.cauchyLoss(0.5), _abi.LOSS_CAUCHY, 0.5)]:
        for s in [0.0, 0.2, 0.49, 3.0, 40.0]:
            assert np.allclose(loss.evaluate(s), oracle.loss(kind, a, s), rtol=1e-14, atol=0)
    tol = L.tolerantLoss(0.8, 0.35)                      # rho'' > 0: the loss behind the Corrector's alpha branch
    for s in [0.0, 0.05, 0.6, 0.8, 1.3, 5.0, 13.0, 0.8 + 40 * 0.35]:
        # atol: rho(0) = b log(1 + e^(-a/b)) - c is exactly 0 on the host; the device contracts it into an FMA (~1e-18)
        assert np.allclose(tol.evaluate(s), oracle.loss(_abi.LOSS_TOLERANT, 0.8, s, 0.35), rtol=1e-12, atol=1e-15)
    with pytest.raises(sk.SkeresError):
        L.tolerantLoss(1.0, 0.0)                         # b > 0 (Ceres CHECKs the same)


def test_ba_tolerant_loss_takes_the_alpha_branch(sk, oracle):
    """Bundle adjustment with TolerantLoss(0.5, 0.3): squared residuals of ~0.5 px^2 sit where rho'' > 0, so the evaluator
    kernel's Corrector runs its alpha != 0 branch on most observations (tests/test_host_logic.py checks the branch's algebra
    on the same device code built for the host).  The solve must reach the oracle's optimum; rows are compared loosely."""
    d = synth.make_bal("small", seed=8)
    loss = sk.PredefinedLossFunctions.tolerantLoss(0.5, 0.3)
    opts = dict(max_trust_region_radius=1e12)
    p, so = oracle_ba(oracle, d, _abi.DENSE_SCHUR, loss=(_abi.LOSS_TOLERANT, 0.5, 0.3), **opts)
    bal, s = gpu_ba(sk, d, _abi.DENSE_SCHUR, loss=loss, **opts)
    assert s.termination_type == so.termination_type == _abi.CONVERGENCE
    assert abs(s.initial_cost - so.initial_cost) <= 1e-11 * so.initial_cost           # evaluator + rho(s)
    assert np.isclose(s.iterations[1].cost, so.iterations[1].cost, rtol=1e-8)         # first corrected step
    assert abs(s.final_cost - so.final_cost) <= COST_RTOL * abs(so.final_cost)
    plain = gpu_ba(sk, d, _abi.DENSE_SCHUR, **opts)[1]
    assert s.initial_cost < 0.99 * plain.initial_cost and s.final_cost < 0.5 * plain.final_cost      # the loss really is in the path (oracle: 0.970, 0.365)


# --------------------------------------------------------------------------------------------------- config 1: CurveFitting, DENSE_QR
def curve_fit(sk, loss=None, y_mod=None, max_it=25):
    d = load("curve_fitting_data.json")
    y = np.array(d["y"]) if y_mod is None else y_mod(np.array(d["y"]))
    m, c = sk.DoubleArray(1), sk.DoubleArray(1)                     # two separate DoubleArrays, CurveFitting.scala:103-106
    loss = loss if loss is not None else sk.PredefinedLossFunctions.trivialLoss()
    problem = sk.Problem()
    for xi, yi in zip(d["x"], y):
        problem.addResidualBlock(sk.ExponentialResidual(xi, yi).toAutoDiffCostFunction(), loss, m.toPointer(), c.toPointer())
    o = sk.Solver.Options()
    o.setMaxNumIterations(max_it)
    o.setLinearSolverType(_abi.DENSE_QR)
    s = sk.Solver.Summary()
    sk.ceres.solve(o, problem, s)
    return np.array([m.get(0), c.get(0)]), s, problem


def test_curve_fitting_matches_oracle_and_ceres_tutorial(sk, oracle):
    x, s, problem = curve_fit(sk)
    assert (problem.numResidualBlocks(), problem.numResiduals(), problem.numParameterBlocks(), problem.numParameters()) == (67, 67, 2, 2)
    d = load("curve_fitting_data.json")
    p = oracle.OracleProblem(np.zeros(2))
    p.add_residual_blocks(_abi.FUNCTOR_EXPONENTIAL_RESIDUAL, np.stack([d["x"], d["y"]], 1), np.tile([0, 1], (67, 1)))
    o = _abi.default_options()
    o.linear_solver_type, o.max_num_iterations = _abi.DENSE_QR, 25
    so = p.solve(o)
    assert_same_trajectory(s, so)
    assert rel_param_diff(x, p.params, 1e-6) <= PARAM_RTOL
    gold = load("ceres_tutorial_curve_fitting_log.json")
    assert len(s.iterations) == len(gold["rows"])
    for row, g in zip(s.iterations, gold["rows"]):
        if row.step_is_successful:
            assert float(f"{row.cost:.6e}") == g[1]
        assert float(f"{row.trust_region_radius:.2e}") == g[6]
    assert round(x[0], 6) == gold["final"]["m"] and round(x[1], 6) == gold["final"]["c"]
    assert "Function tolerance reached" in s.message and "CONVERGENCE" in s.briefReport()
    assert s.briefReport().startswith("Ceres Solver Report: Iterations: 14, Initial cost: 1.211734e+02, Final cost: 1.056751e+00")


def test_hello_world_example(sk, oracle):
    """HelloWorld.scala:11-35 through the mirrored API: x 0.5 -> 10, rows equal the oracle's and the Ceres tutorial's log."""
    x = sk.DoubleArray(1)
    x.set(0, 0.5)
    problem = sk.Problem()
    problem.addResidualBlock(sk.HelloCostFunctor().toAutoDiffCostFunction(), sk.PredefinedLossFunctions.trivialLoss(), x.toPointer())
    options = sk.Solver.Options()
    # the example keeps Ceres' default SPARSE_NORMAL_CHOLESKY (HelloWorld.scala:27-31 never sets a solver): it runs on the dense
    # back end, same rows as DENSE_QR
    assert options.linear_solver_type == _abi.SPARSE_NORMAL_CHOLESKY
    s0 = sk.Solver.Summary()
    sk.ceres.solve(options, problem, s0)
    assert abs(x.get(0) - 10.0) < 1e-7 and s0.linear_solver_type_used == _abi.SPARSE_NORMAL_CHOLESKY
    x.set(0, 0.5)
    options.setLinearSolverType(_abi.DENSE_QR)
    s = sk.Solver.Summary()
    sk.ceres.solve(options, problem, s)
    assert [r.cost for r in s0.iterations] == [r.cost for r in s.iterations]
    p = oracle.OracleProblem(np.array([0.5]))
    p.add_residual_blocks(_abi.FUNCTOR_HELLO_WORLD, np.zeros((1, 0)), np.array([[0]]))
    o = _abi.default_options()
    o.linear_solver_type = _abi.DENSE_QR
    assert_same_trajectory(s, p.solve(o))
    gold = load("ceres_tutorial_helloworld_log.json")
    assert len(s.iterations) == len(gold["rows"])
    for r, g in zip(s.iterations, gold["rows"]):            # the last cost is rounding-dominated (x = 10 - 3e-8): 1e-5, not 6 digits
        assert np.isclose(r.cost, g[1], rtol=1e-5) and float(f"{r.trust_region_radius:.2e}") == g[6]
        assert np.isclose(r.gradient_max_norm, g[3], rtol=1e-2) and np.isclose(r.step_norm, g[4], rtol=1e-2)
    assert abs(x.get(0) - 10.0) < 1e-7
    assert s.briefReport().startswith("Ceres Solver Report: Iterations: 3, Initial cost: 4.512500e+01, Final cost: 5.0125")


@pytest.mark.parametrize("f2", ["F2", "F2a"])
def test_powell_example(sk, oracle, f2):
    """Powell.scala:54-95 (four scalar blocks in four DoubleArrays, DENSE_QR, 100 iterations) with F2 as Powell.scala computes
    it, and with PowellAnalytic.scala's f2 -- the Ceres tutorial's problem, whose published log the rows must reproduce."""
    xs = [sk.DoubleArray(1) for _ in range(4)]
    for a, v in zip(xs, [3.0, -1.0, 0.0, 1.0]):
        a.set(0, v)
    loss = sk.PredefinedLossFunctions.trivialLoss()
    problem = sk.Problem()
    F2 = getattr(sk.Powell, f2)
    for functor, (i, j) in zip([sk.Powell.F1, F2, sk.Powell.F3, sk.Powell.F4], [(0, 1), (2, 3), (1, 2), (0, 3)]):
        problem.addResidualBlock(functor().toAutoDiffCostFunction(), loss, xs[i].toPointer(), xs[j].toPointer())
    options = sk.Solver.Options()
    options.setMaxNumIterations(100)
    options.setLinearSolverType(_abi.DENSE_QR)
    s = sk.Solver.Summary()
    sk.ceres.solve(options, problem, s)
    xout = np.array([a.get(0) for a in xs])
    p = oracle.OracleProblem(np.array([3.0, -1.0, 0.0, 1.0]))
    for fid, blocks in zip([_abi.FUNCTOR_POWELL_F1, F2.functor_id, _abi.FUNCTOR_POWELL_F3, _abi.FUNCTOR_POWELL_F4], [(0, 1), (2, 3), (1, 2), (0, 3)]):
        p.add_residual_blocks(fid, np.zeros((1, 0)), np.array([blocks]))
    o = _abi.default_options()
    o.linear_solver_type, o.max_num_iterations = _abi.DENSE_QR, 100
    so = p.solve(o)
    assert_same_trajectory(s, so)
    assert np.allclose(xout, p.params, rtol=1e-6, atol=1e-12) and np.all(np.abs(xout) < 1e-3) and s.final_cost < 1e-14
    if f2 == "F2a":
        gold = load("ceres_tutorial_powell_log.json")
        assert len(s.iterations) == len(gold["rows"])
        for r, g in zip(s.iterations, gold["rows"]):
            assert np.isclose(r.cost, g[1], rtol=2e-6) and float(f"{r.trust_region_radius:.2e}") == g[6]
            assert np.isclose(r.step_norm, g[4], rtol=1e-2) and np.isclose(r.gradient_max_norm, g[3], rtol=1e-2)
        assert [float(f"{v:.5e}") for v in xout] == [float(f"{v:.5e}") for v in gold["final"]["x"]]


def test_example_functors_at_the_evaluate_boundary(sk, oracle):
    """HelloWorld / Powell functors through AutoDiffCostFunction.evaluate: residuals and row-major Jacobian blocks equal the
    oracle's Jet evaluation, and the hand derivatives."""
    cases = [(_abi.FUNCTOR_HELLO_WORLD, [[0.5]]), (_abi.FUNCTOR_POWELL_F1, [[3.0], [-1.0]]), (_abi.FUNCTOR_POWELL_F2, [[0.25], [1.0]]),
             (_abi.FUNCTOR_POWELL_ANALYTIC_F2, [[0.25], [1.0]]), (_abi.FUNCTOR_POWELL_F3, [[-1.0], [0.5]]), (_abi.FUNCTOR_POWELL_F4, [[3.0], [1.0]])]
    hand = {_abi.FUNCTOR_HELLO_WORLD: ([9.5], [[-1.0]]), _abi.FUNCTOR_POWELL_F1: ([-7.0], [[1.0], [10.0]]),
            _abi.FUNCTOR_POWELL_F2: ([np.sqrt(5.0) * 0.25 - 1.0], [[np.sqrt(5.0)], [-1.0]]),
            _abi.FUNCTOR_POWELL_ANALYTIC_F2: ([np.sqrt(5.0) * (0.25 - 1.0)], [[np.sqrt(5.0)], [-np.sqrt(5.0)]]),
            _abi.FUNCTOR_POWELL_F3: ([4.0], [[-4.0], [8.0]]), _abi.FUNCTOR_POWELL_F4: ([np.sqrt(10.0) * 4.0], [[4.0 * np.sqrt(10.0)], [-4.0 * np.sqrt(10.0)]])}
    for fid, params in cases:
        ok, res, jac = sk.CostFunction(fid, []).evaluate_host(params)
        oko, ro, jo = oracle.evaluate(fid, [], params)
        assert ok and oko
        assert np.allclose(res, ro, rtol=4e-15, atol=0) and all(np.allclose(a, b, rtol=4e-15, atol=0) for a, b in zip(jac, jo))
        assert np.allclose(res, hand[fid][0], rtol=1e-14) and all(np.allclose(np.ravel(a), b, rtol=1e-14) for a, b in zip(jac, hand[fid][1]))


def robust_curve_fit(sk, loss, max_it):
    """RobustCurveFitting.scala:97-128 verbatim: its own data, one shared loss object, two DoubleArray(1) blocks from (0, 0)."""
    d = load("robust_curve_fitting_data.json")
    m, c = sk.DoubleArray(1), sk.DoubleArray(1)
    problem = sk.Problem()
    for xi, yi in zip(d["x"], d["y"]):
        problem.addResidualBlock(sk.ExponentialResidual(xi, yi).toAutoDiffCostFunction(), loss, m.toPointer(), c.toPointer())
    o = sk.Solver.Options()
    o.setMaxNumIterations(max_it)
    o.setLinearSolverType(_abi.DENSE_QR)
    s = sk.Solver.Summary()
    sk.ceres.solve(o, problem, s)
    return np.array([m.get(0), c.get(0)]), s


def test_robust_curve_fitting(sk, oracle):
    """RobustCurveFitting.scala with the reference's own data (outliers at :41-42), CauchyLoss(0.5) (:107) and 25 iterations
    (:117) -- the Corrector on the device.  Row by row against the oracle, and the published tutorial answer."""
    d = load("robust_curve_fitting_data.json")
    for make, kind, a in [(lambda: sk.PredefinedLossFunctions.cauchyLoss(d["cauchy_a"]), _abi.LOSS_CAUCHY, d["cauchy_a"]),
                          (lambda: sk.PredefinedLossFunctions.huberLoss(1.0), _abi.LOSS_HUBER, 1.0)]:
        x, s = robust_curve_fit(sk, make(), d["max_num_iterations"])
        p = oracle.OracleProblem(np.zeros(2))
        p.add_residual_blocks(_abi.FUNCTOR_EXPONENTIAL_RESIDUAL, np.stack([d["x"], d["y"]], 1), np.tile([0, 1], (67, 1)), kind, a)
        o = _abi.default_options()
        o.linear_solver_type, o.max_num_iterations = _abi.DENSE_QR, d["max_num_iterations"]
        so = p.solve(o)
        assert_same_trajectory(s, so)
        assert rel_param_diff(x, p.params, 1e-6) <= PARAM_RTOL
        if kind == _abi.LOSS_CAUCHY:
            assert [float(f"{v:.6f}") for v in x] == [0.287605, 0.151213]      # the Ceres tutorial's robust fit


def test_max_iterations_and_zero_iterations(sk):
    x, s, _ = curve_fit(sk, max_it=3)
    assert s.termination_type == _abi.NO_CONVERGENCE and len(s.iterations) == 4
    assert s.message == "Maximum number of iterations reached. Number of iterations: 3."
    x, s, _ = curve_fit(sk, max_it=0)
    assert len(s.iterations) == 1 and np.all(x == 0.0) and s.initial_cost == s.final_cost


# --------------------------------------------------------------------------------------------------- config 2/3: bundle adjustment
@pytest.mark.parametrize("shape,seed", [("tiny", 1), ("small", 2), ("ladybug-49", 1)])
@pytest.mark.parametrize("lst", [_abi.DENSE_SCHUR, _abi.SPARSE_SCHUR])
def test_ba_explicit_schur_matches_oracle(sk, oracle, shape, seed, lst):
    """SimpleBundleAdjuster.scala:147-149 (DENSE_SCHUR) and BASELINE configs[1] (SPARSE_SCHUR): exact reduced solve."""
    d = synth.make_bal(shape, seed=seed)
    p, so = oracle_ba(oracle, d, lst)
    bal, s = gpu_ba(sk, d, lst)
    assert_same_trajectory(s, so)
    assert rel_param_diff(bal.parameters.toArray(), p.params) <= PARAM_RTOL
    assert s.num_residual_blocks == d.num_observations and s.num_parameters == d.parameters.size


@pytest.mark.parametrize("shape,seed", [("tiny", 1), ("small", 2), ("small", 5), ("ladybug-49", 1)])
def test_ba_iterative_schur_matches_oracle(sk, oracle, shape, seed):
    """ITERATIVE_SCHUR + SCHUR_JACOBI: same LM rows and the same number of PCG iterations per LM iteration."""
    d = synth.make_bal(shape, seed=seed)
    p, so = oracle_ba(oracle, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI)
    bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI)
    assert_same_trajectory(s, so, row_rtol=1e-6)
    # small cases meet 1e-5; the Ladybug shape is held to the oracle's own rounding envelope (see envelope_bound) and to
    # 1e-5 with converged solves (next test)
    tol = PARAM_RTOL if shape != "ladybug-49" else envelope_bound("ladybug-49/SCHUR_JACOBI/eta0.1")
    assert rel_param_diff(bal.parameters.toArray(), p.params) <= tol


@pytest.mark.parametrize("shape,seed,prec", [("ladybug-49", 1, _abi.SCHUR_JACOBI), ("ladybug-49", 1, _abi.JACOBI), ("small", 3, _abi.JACOBI),
                                             ("small", 3, _abi.IDENTITY)])
def test_ba_converged_linear_solves_meet_the_parameter_tolerance(sk, oracle, shape, seed, prec):
    """The north star's parameter tolerance (1e-5 relative) where it is attainable: with the linear solves converged
    (eta = 1e-10) every LM step is unique, so what is left between the device and the oracle is rounding, not the path a
    truncated CG takes.  Same rows, PCG counts within 5 % (the Q-test flips within a few iterations once the decrease per
    iteration is at rounding level), costs to 1e-7, parameters to 1e-5 -- SCHUR_JACOBI included.  The one exception is the
    IDENTITY preconditioner: unpreconditioned CG needs 260+ iterations on a 108-unknown system, i.e. it runs long past the
    point where finite-precision CG has lost conjugacy, and its iterate at the stopping test moves with the summation order
    (measured on a B200, profiles/r02_parity_rows.md: 260 vs 270 iterations, cost 9e-9, parameters 2.3e-5, while JACOBI on
    the same problem agrees to 1e-12); it is held to the oracle's own rounding envelope instead."""
    d = synth.make_bal(shape, seed=seed)
    kw = dict(eta=1e-10, max_linear_solver_iterations=3000, max_num_iterations=3)
    p, so = oracle_ba(oracle, d, _abi.ITERATIVE_SCHUR, prec, **kw)
    bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, prec, **kw)
    assert_same_trajectory(s, so, exact_rows=False, row_rtol=1e-7)
    for a, b in zip(s.iterations, so.iterations):
        assert abs(a.linear_solver_iterations - b.linear_solver_iterations) <= max(3, 0.05 * b.linear_solver_iterations)
    tol = PARAM_RTOL if prec != _abi.IDENTITY else envelope_bound("small-3/IDENTITY/eta1e-10")
    assert rel_param_diff(bal.parameters.toArray(), p.params) <= tol


@pytest.mark.parametrize("shape,seed", [("tiny", 1), ("small", 2)])
def test_ba_jacobi_preconditioner(sk, oracle, shape, seed):
    """JACOBI is Ceres' DEFAULT preconditioner_type: a caller who only sets ITERATIVE_SCHUR gets it.  For the implicit
    Schur complement it is the inverse of the block diagonal of F'F + D^2 (block_diagonal_FtF_inverse)."""
    d = synth.make_bal(shape, seed=seed)
    p, so = oracle_ba(oracle, d, _abi.ITERATIVE_SCHUR, _abi.JACOBI)
    bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.JACOBI)
    assert_same_trajectory(s, so)
    assert rel_param_diff(bal.parameters.toArray(), p.params) <= PARAM_RTOL


def test_ba_default_preconditioner_is_accepted(sk):
    d = synth.make_bal("tiny", seed=1)
    bal = sk.BalProblem.fromArrays(d)
    o = sk.Solver.Options()                       # preconditioner_type left at the Ceres default (JACOBI)
    o.setLinearSolverType(_abi.ITERATIVE_SCHUR)
    assert o.preconditioner_type == _abi.JACOBI
    s = sk.Solver.Summary()
    sk.ceres.solve(o, bal.buildProblem(), s)
    assert s.termination_type == _abi.CONVERGENCE and s.preconditioner_type_used == _abi.JACOBI


def test_ba_identity_preconditioner(sk, oracle):
    d = synth.make_bal("small", seed=3)
    p, so = oracle_ba(oracle, d, _abi.ITERATIVE_SCHUR, _abi.IDENTITY)
    bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.IDENTITY)
    # unpreconditioned CG needs ~70+ iterations per solve here and its Q-based stopping test then flips within a
    # few iterations under a different summation order: LM rows must agree, PCG counts within 10 %.
    assert_same_trajectory(s, so, exact_rows=False)
    for a, b in zip(s.iterations, so.iterations):
        assert abs(a.linear_solver_iterations - b.linear_solver_iterations) <= max(2, 0.1 * b.linear_solver_iterations)
    # Parameters: measured 2.9e-3 on a B200; the oracle's own rounding envelope for this run is 1e-4 (envelope_bound), and
    # test_ba_converged_linear_solves_meet_the_parameter_tolerance holds the same problem to 1e-5 with converged solves.
    assert rel_param_diff(bal.parameters.toArray(), p.params) <= envelope_bound("small-3/IDENTITY/eta0.1")


# --------------------------------------------------------------------------------------------------- BASELINE-sized configs vs committed oracle rows
def _digest_indices(n_cam, n_pt, kept):
    pts = np.unique(np.linspace(0, n_pt - 1, kept).astype(np.int64))
    return np.concatenate([np.arange(9 * n_cam, dtype=np.int64)] + [9 * n_cam + 3 * p + np.arange(3, dtype=np.int64) for p in pts])


@pytest.mark.parametrize("fixture", [f for f in ("oracle_venice_1778_rows.json", "oracle_final_13682_rows.json")
                                     if os.path.exists(os.path.join(HERE, "golden", f))])
def test_baseline_sized_rows_follow_the_oracle_fixture(sk, fixture):
    """BASELINE.json configs[2] (Venice-1778 shape, the headline) and configs[4] (Final-13682 shape) on ONE GPU against the
    oracle's committed LM rows (tests/golden/make_oracle_ba_rows.py; the oracle needs 10 min / 1 h for them).  Row by row:
    same step validity / acceptance, cost to 1e-8, radius to 1e-4, identical PCG iteration counts while a solve stays below
    100 iterations and within 2 % beyond (the Q-test of a long truncated solve flips within a few iterations under a
    different summation order; first such row on the Venice shape: row 8, 153 iterations); final cost to 1e-6."""
    g = load(fixture)
    c = g["case"]
    d = synth.make_bal(c["shape"], seed=c["seed"])
    assert float(np.sum(d.parameters) + np.sum(d.observations)) == g["input_checksum"], "synthetic input differs from the fixture's"
    bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI, max_num_iterations=c["max_num_iterations"])
    rows = g["iterations"]
    assert len(s.iterations) == len(rows) and s.termination_type == g["termination_type"]
    assert abs(s.initial_cost - g["initial_cost"]) <= 1e-12 * g["initial_cost"]
    for a, b in zip(s.iterations, rows):
        assert (a.iteration, a.step_is_valid, a.step_is_successful) == (b["iteration"], b["step_is_valid"], b["step_is_successful"])
        # a rejected step's cost is the objective at a point that is then discarded; on the Final shape row 2 lands where points
        # project through near-zero depth (cost 1.2e18 against 5.8e6 around it) and moves 1.8e-8 with the summation order
        assert np.isclose(a.cost, b["cost"], rtol=1e-8 if b["step_is_successful"] else 1e-6), (a.iteration, a.cost, b["cost"])
        assert np.isclose(a.trust_region_radius, b["trust_region_radius"], rtol=1e-4)
        n = b["linear_solver_iterations"]
        # the gradient at a point reached through a long truncated solve moves with it (measured 4e-3 at row 12 of the Venice shape)
        assert np.isclose(a.gradient_max_norm, b["gradient_max_norm"], rtol=1e-3 if n < 100 else 5e-2, atol=1e-6)
        if n < 100:
            assert a.linear_solver_iterations == n, (a.iteration, a.linear_solver_iterations, n)
        else:
            assert abs(a.linear_solver_iterations - n) <= 0.02 * n, (a.iteration, a.linear_solver_iterations, n)
    assert abs(s.final_cost - g["final_cost"]) <= COST_RTOL * g["final_cost"]
    # parameters: cameras + 300 points kept in the fixture.  A truncated SCHUR_JACOBI run is held to cost, not to 1e-5 in
    # parameters (envelope_bound above); the digest still catches a wrong block or a wrong scatter.
    # (measured on the Venice shape after 13 rows with solves of up to 282 iterations: 0.054 with the set-up kernel's projector
    # form of the SchurJacobi blocks; a wrong block or scatter gives differences of order 1 and more)
    x = bal.parameters.toArray()[_digest_indices(g["n_cam"], g["n_pt"], g["param_digest_points"])]
    assert rel_param_diff(x, np.array(g["param_digest"])) <= 1e-1


LONG_TRACK_CASE = dict(n_cam=600, n_pt=1500, n_obs=12000, seed=5, long_tracks=(257, 600, 513))     # chunk tiles: 2, 3, 3
LONG_TRACK_SMALL = dict(n_cam=270, n_pt=220, n_obs=2200, seed=11, long_tracks=(257, 270, 264))     # the DENSE_SCHUR fixture's case


@pytest.mark.parametrize("case,prec,max_it", [(LONG_TRACK_CASE, _abi.JACOBI, 50), (LONG_TRACK_SMALL, _abi.JACOBI, 50),
                                              (LONG_TRACK_CASE, _abi.SCHUR_JACOBI, 2), (LONG_TRACK_SMALL, _abi.SCHUR_JACOBI, 4)])
def test_ba_long_tracks_iterative_schur(sk, oracle, case, prec, max_it):
    """Points seen by more cameras than one tile holds (257 .. 600 observations: 2 and 3 chunk tiles) go through the
    k_ba_*_giant kernels; everything the LM loop derives from them must follow the oracle row by row.
    Measured on a B200 (profiles/r01_long_track_rows*.log): with JACOBI all 13 / 10 rows agree to 1e-14 in cost through
    PCG solves of 300+ iterations, parameters to 2e-10.  SCHUR_JACOBI is compared on the rows before its first long
    PCG solve (see test_ba_long_tracks_schur_jacobi_full_run for why)."""
    d = synth.make_bal(**case)
    assert np.sort(np.bincount(d.point_index))[-3:].tolist() == sorted(case["long_tracks"])
    p, so = oracle_ba(oracle, d, _abi.ITERATIVE_SCHUR, prec, max_num_iterations=max_it)
    bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, prec, max_num_iterations=max_it)
    assert_same_trajectory(s, so, row_rtol=1e-9)
    for a, b in zip(s.iterations, so.iterations):
        assert np.isclose(a.gradient_max_norm, b.gradient_max_norm, rtol=1e-6, atol=1e-9)
    assert rel_param_diff(bal.parameters.toArray(), p.params) <= PARAM_RTOL


@pytest.mark.parametrize("case", [LONG_TRACK_CASE, LONG_TRACK_SMALL])
def test_ba_long_tracks_schur_jacobi_full_run(sk, oracle, case):
    """Full SCHUR_JACOBI runs at Ceres' default eta on the long-track problems: the least well-conditioned cases of the suite.
    The oracle itself does not reproduce its own run here under a rounding-level perturbation -- against its -ffp-contract=fast
    build the parameters move by 4e-2, and with the eliminator's two sums merely accumulated separately it takes 13 LM rows
    instead of 7 and ends 9e-5 away in cost (tests/golden/schur_jacobi_rounding_envelope.json, "long-tracks").  So rows are
    compared one by one only up to the first PCG solve of 50+ iterations (before it the two sides agree to 1e-9); after it what
    is asserted is what survives such perturbations: both sides converge, and to the same cost within 10 x the envelope."""
    d = synth.make_bal(**case)
    p, so = oracle_ba(oracle, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI)
    bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI)
    assert s.termination_type == so.termination_type == _abi.CONVERGENCE
    for a, b in zip(s.iterations, so.iterations):
        assert (a.step_is_valid, a.step_is_successful) == (b.step_is_valid, b.step_is_successful)
        if b.linear_solver_iterations >= 50:          # measured (profiles/r02_parity_rows.md): 1e-13 before this row, 1.6e-8 / 7e-5 on it
            break
        assert np.isclose(a.cost, b.cost, rtol=1e-9) and a.linear_solver_iterations == b.linear_solver_iterations
    env = load("schur_jacobi_rounding_envelope.json")["cases"]["long-tracks/SCHUR_JACOBI/eta0.1"]
    bound = 10.0 * max(v["final_cost_rel_diff"] for v in env.values() if isinstance(v, dict))
    assert abs(s.final_cost - so.final_cost) <= max(COST_RTOL, bound) * abs(so.final_cost)


def test_ba_long_tracks_first_iteration_is_exact(sk, oracle):
    """One LM iteration with a tight PCG (eta 1e-8, 252 iterations, JACOBI): gradient, Jacobi scaling, Schur set-up,
    implicit product and back-substitution of the long tracks all enter the step, and the step is unique."""
    d = synth.make_bal(**LONG_TRACK_CASE)
    kw = dict(max_num_iterations=1, eta=1e-8, max_linear_solver_iterations=3000)
    p, so = oracle_ba(oracle, d, _abi.ITERATIVE_SCHUR, _abi.JACOBI, **kw)
    bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.JACOBI, **kw)
    assert len(s.iterations) == len(so.iterations) == 2
    assert s.iterations[1].linear_solver_iterations == so.iterations[1].linear_solver_iterations
    assert np.isclose(s.iterations[1].cost, so.iterations[1].cost, rtol=1e-11)
    assert np.isclose(s.iterations[1].trust_region_radius, so.iterations[1].trust_region_radius, rtol=1e-9)
    assert rel_param_diff(bal.parameters.toArray(), p.params) <= PARAM_RTOL      # measured 6.3e-11


def test_ba_long_tracks_dense_schur_matches_oracle_fixture(sk):
    """DENSE_SCHUR on long tracks against the oracle's committed output (tests/golden/make_oracle_long_tracks.py; the
    oracle's dense reduced solve needs minutes of CPU at 270 cameras, so it is not re-run here)."""
    g = load("oracle_long_tracks_dense_schur.json")
    case = dict(g["case"]); case["long_tracks"] = tuple(case["long_tracks"])
    d = synth.make_bal(**case)
    assert float(np.sum(d.parameters) + np.sum(d.observations)) == g["input_checksum"]      # same synthetic input
    bal, s = gpu_ba(sk, d, _abi.DENSE_SCHUR)
    assert s.termination_type == g["termination_type"] and len(s.iterations) == len(g["iterations"])
    assert (s.num_successful_steps, s.num_unsuccessful_steps) == (g["num_successful_steps"], g["num_unsuccessful_steps"])
    for a, b in zip(s.iterations, g["iterations"]):
        assert (a.iteration, int(a.step_is_valid), int(a.step_is_successful)) == (b["iteration"], b["step_is_valid"], b["step_is_successful"])
        assert np.isclose(a.cost, b["cost"], rtol=1e-6) and np.isclose(a.trust_region_radius, b["trust_region_radius"], rtol=1e-4)
    assert abs(s.final_cost - g["final_cost"]) <= COST_RTOL * abs(g["final_cost"])
    assert rel_param_diff(bal.parameters.toArray(), np.array(g["params"])) <= PARAM_RTOL


MEDIUM_TRACK_CASE = dict(n_cam=300, n_pt=900, n_obs=6000, seed=4, long_tracks=(33, 64, 100, 200, 256))   # up to exactly one tile


def test_ba_medium_tracks_iterative_schur(sk, oracle):
    """Tracks longer than one warp but within one tile (33 .. 256 observations; 256 fills a tile alone): the per-point
    sums of the tile kernels walk them with one thread.  Rows must follow the oracle as everywhere else."""
    d = synth.make_bal(**MEDIUM_TRACK_CASE)
    assert np.sort(np.bincount(d.point_index))[-5:].tolist() == sorted(MEDIUM_TRACK_CASE["long_tracks"])
    p, so = oracle_ba(oracle, d, _abi.ITERATIVE_SCHUR, _abi.JACOBI, max_num_iterations=12)
    bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.JACOBI, max_num_iterations=12)
    assert_same_trajectory(s, so, row_rtol=1e-9)
    assert rel_param_diff(bal.parameters.toArray(), p.params) <= PARAM_RTOL


@pytest.mark.parametrize("sums", ["chunked", "serial"])
@pytest.mark.parametrize("case", [dict(shape="small", seed=2), LONG_TRACK_SMALL, MEDIUM_TRACK_CASE])
def test_matvec_kernels_agree_bitwise(sk, monkeypatch, case, sums):
    """The persistent TMA-prefetching implicit-Schur product (k_ba_matvec_tma) and the classic one-CTA-per-tile kernel add in
    the same order -- fixed by the point / segment lists for the serial chains (the default), by the chunk tables of the tile
    records for the two-level sums (SKERES_MATVEC_SUMS=chunked): every LM row and every parameter must be identical, not
    merely close.  Both run inside the kernel SEQUENCE of the PCG loop here (SKERES_PCG=sequence)."""
    d = synth.make_bal(**case)
    runs = []
    monkeypatch.setenv("SKERES_MATVEC_SUMS", sums)             # all read when a solver is constructed
    monkeypatch.setenv("SKERES_PCG", "sequence")
    for mode in ("classic", "tma"):
        monkeypatch.setenv("SKERES_MATVEC", mode)
        bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI)
        runs.append(([r.cost for r in s.iterations], [r.linear_solver_iterations for r in s.iterations], bal.parameters.toArray()))
    for other in runs[1:]:
        assert runs[0][0] == other[0] and runs[0][1] == other[1]
        assert np.array_equal(runs[0][2], other[2])


@pytest.mark.parametrize("case,prec,kw", [(dict(shape="small", seed=2), _abi.SCHUR_JACOBI, {}), (dict(shape="ladybug-49", seed=1), _abi.SCHUR_JACOBI, {}),
                                          (MEDIUM_TRACK_CASE, _abi.JACOBI, {}), (dict(shape="small", seed=3), _abi.IDENTITY, {}),
                                          (dict(shape="ladybug-49", seed=1), _abi.SCHUR_JACOBI, dict(eta=1e-10, max_linear_solver_iterations=3000, max_num_iterations=3)),
                                          (dict(shape="tiny", seed=1), _abi.SCHUR_JACOBI, dict(max_linear_solver_iterations=3)),
                                          # more virtual blocks of 8 cameras (313) than persistent CTAs (196 tiles): two blocks per round
                                          (dict(n_cam=2500, n_pt=8000, n_obs=50000, seed=7), _abi.SCHUR_JACOBI, dict(max_num_iterations=6))])
def test_fused_pcg_matches_kernel_sequence_bitwise(sk, monkeypatch, case, prec, kw):
    """The fused PCG solve (one persistent cooperative kernel per linear solve: products, vector phases and termination tests
    behind grid barriers, pcg_fused.cu) is built from the device functions of the kernel sequence (pcg_kernels.cu) and must
    reproduce it bit for bit: LM rows, PCG iteration counts -- through residual resets every 10th iteration, the iteration
    cap and solves of hundreds of iterations -- and parameters.  The sequence is run with one warp per camera in its
    second-level sums (SKERES_PCG_WPC=1), the order the fused kernel uses."""
    d = synth.make_bal(**case)
    runs = []
    monkeypatch.setenv("SKERES_PCG_WPC", "1")
    for mode in ("sequence", "fused"):
        monkeypatch.setenv("SKERES_PCG", mode)
        bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, prec, **kw)
        fams = s.kernel_times()
        assert (fams["pcg_solve"][1] > 0) == (mode == "fused"), fams       # the path that was asked for is the one that ran
        runs.append(([r.cost for r in s.iterations], [r.linear_solver_iterations for r in s.iterations], bal.parameters.toArray(),
                     fams["schur_matvec"][1]))
    assert runs[0][0] == runs[1][0] and runs[0][1] == runs[1][1]
    assert np.array_equal(runs[0][2], runs[1][2])
    assert runs[0][3] == runs[1][3]                                         # products executed: counted by the host / derived from the readback


@pytest.mark.parametrize("pcg", ["fused", "sequence"])
def test_l2_copy_policies_do_not_change_results(sk, monkeypatch, pcg):
    """The copies of the implicit-Schur product carry L2 cache policies (ba_product.cuh: evict_first for the Jacobian stream,
    evict_last for the first SKERES_L2_KEEP_MB megabytes of tiles; chosen by the working set of the vector phases, ba_solver.cu).
    A cache policy moves no bit: plain copies (-1), an all-stream policy (0), the default and an everything-resident policy
    give identical LM rows, PCG counts and parameters -- in the fused solve and in the kernel sequence, whose stand-alone
    product kernel issues the same copies."""
    d = synth.make_bal(n_cam=300, n_pt=60000, n_obs=400000, seed=3)
    monkeypatch.setenv("SKERES_PCG", pcg)
    runs = []
    for mb in ("-1", "0", None, "1000"):
        if mb is None:
            monkeypatch.delenv("SKERES_L2_KEEP_MB", raising=False)
        else:
            monkeypatch.setenv("SKERES_L2_KEEP_MB", mb)
        bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI, max_num_iterations=6)
        runs.append(([r.cost for r in s.iterations], [r.linear_solver_iterations for r in s.iterations], bal.parameters.toArray()))
    assert len(runs[0][0]) == 7 and sum(runs[0][1]) > 20
    for r in runs[1:]:
        assert r[0] == runs[0][0] and r[1] == runs[0][1] and np.array_equal(r[2], runs[0][2])


@pytest.mark.parametrize("sums", ["serial", "chunked"])
@pytest.mark.parametrize("case", [dict(shape="ladybug-49", seed=1), MEDIUM_TRACK_CASE, LONG_TRACK_SMALL])
def test_device_built_tile_records_equal_the_host_builder(sk, monkeypatch, case, sums):
    """The per-tile metadata records of the implicit-Schur product are packed by a kernel (k_ba_build_tile_records); the host
    builder (ba_layout.cu, checked against a direct evaluation by tests/hostcheck) stays as SKERES_TILE_REC=host.  Every field
    of a record steers the product -- slots, point / segment starts, permutations, camera ids, output rows and, for the
    two-level sums, the chunk tables -- so identical LM rows and identical parameters, bit for bit, mean identical records."""
    d = synth.make_bal(**case)
    monkeypatch.setenv("SKERES_MATVEC_SUMS", sums)
    runs = []
    for where in ("host", "device"):
        monkeypatch.setenv("SKERES_TILE_REC", where)
        bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI, max_num_iterations=5)
        runs.append(([r.cost for r in s.iterations], [r.linear_solver_iterations for r in s.iterations], bal.parameters.toArray()))
    assert runs[0][0] == runs[1][0] and runs[0][1] == runs[1][1] and np.array_equal(runs[0][2], runs[1][2])


@pytest.mark.parametrize("case", [dict(shape="tiny", seed=1), dict(shape="ladybug-49", seed=1), MEDIUM_TRACK_CASE,
                                  dict(n_cam=2500, n_pt=8000, n_obs=50000, seed=7), dict(n_cam=300, n_pt=60000, n_obs=400000, seed=3)])
def test_device_built_layout_equals_the_host_builder(sk, monkeypatch, case):
    """the layout (dense ids, point CSR, tiles, tile-local camera segments, camera -> segment lists) is built by
    kernels from the uploaded residual-block table (ba_layout_device.cu); the host builder (ba_layout.cu) stays as
    SKERES_LAYOUT=host and for what the device path hands back (unsorted input, long tracks, explicit Schur solvers).  Every
    array of the layout steers the kernels, so bit-identical LM rows and parameters mean identical layouts."""
    d = synth.make_bal(**case)
    runs = []
    for where in ("host", "device"):
        monkeypatch.setenv("SKERES_LAYOUT", where)
        bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI, max_num_iterations=5)
        runs.append(([r.cost for r in s.iterations], [r.linear_solver_iterations for r in s.iterations], bal.parameters.toArray(),
                     (s.num_residual_blocks, s.num_parameter_blocks, s.num_parameters)))
    assert runs[0][0] == runs[1][0] and runs[0][1] == runs[1][1] and np.array_equal(runs[0][2], runs[1][2]) and runs[0][3] == runs[1][3]


def test_fused_pcg_needs_no_host_polling(sk, monkeypatch):
    """north_star (4): one readback per LM iteration.  With the fused solve a linear solve is ONE launch whatever its
    iteration count, so the launches of a whole solve are a fixed number per LM iteration."""
    monkeypatch.setenv("SKERES_PCG", "fused")
    d = synth.make_bal("ladybug-49", seed=1)
    bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI)
    fams = s.kernel_times()
    assert fams["pcg_solve"][1] == s.num_linear_solves and fams["pcg_vector"][1] == 2 * s.num_linear_solves   # begin + start2 only
    assert sum(r.linear_solver_iterations for r in s.iterations) > 200


def test_two_level_sums_follow_the_serial_sums(sk, monkeypatch):
    """The chunked two-level sums only re-associate additions: the first LM rows agree with the serial chains to rounding."""
    d = synth.make_bal("ladybug-49", seed=1)
    runs = []
    for sums in ("chunked", "serial"):
        monkeypatch.setenv("SKERES_MATVEC_SUMS", sums)
        bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.JACOBI, max_num_iterations=4)
        runs.append(s)
    for a, b in zip(runs[0].iterations, runs[1].iterations):
        assert a.linear_solver_iterations == b.linear_solver_iterations and np.isclose(a.cost, b.cost, rtol=1e-12)


def residuals_at(oracle, d, params):
    """Reprojection residuals (gauge-invariant) at a parameter vector, evaluated by the oracle on the host."""
    p = oracle.OracleProblem(params)
    p.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets())
    return p.evaluate()[1]


def test_ba_tight_tolerances_reach_the_same_optimum(sk, oracle):
    """Both sides iterated to the optimum (function_tolerance 1e-12) by DIFFERENT linear solvers.
    These synthetic problems do not fix the 7-dof gauge (global similarity), so the minimiser is an orbit, not a
    point: measured on a B200, the two parameter vectors differ by 0.35 relative while the costs agree to 1e-9.
    Parameters are therefore not comparable here; the gauge-invariant quantities are: the cost and the residuals."""
    d = synth.make_bal("small", seed=4)
    p, so = oracle_ba(oracle, d, _abi.DENSE_SCHUR, function_tolerance=1e-12, max_num_iterations=60)
    bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI, function_tolerance=1e-12, max_num_iterations=60, eta=1e-3)
    assert s.termination_type == _abi.CONVERGENCE
    assert abs(s.final_cost - so.final_cost) <= 1e-9 * so.final_cost
    r_gpu, r_ora = residuals_at(oracle, d, bal.parameters.toArray()), residuals_at(oracle, d, p.params)
    worst = float(np.max(np.abs(r_gpu - r_ora)))
    # measured on a B200: 8.2e-6 px (rms residual 0.39 px) with the parameters 0.35 apart - the gauge orbit.
    # Bound 1e-4 px: ~12x above the measurement, 5000x below the observation noise.
    assert worst <= 1e-4, f"residuals at the two optima differ by {worst:.3e} px"


@pytest.mark.parametrize("lst", [_abi.DENSE_SCHUR, _abi.ITERATIVE_SCHUR])
def test_ba_unsorted_observations(sk, oracle, lst):
    """Residual blocks added in arbitrary order (the reference adds them in file order) give the same solve."""
    d = synth.make_bal("small", seed=6)
    order = np.random.default_rng(2).permutation(d.num_observations)
    p, so = oracle_ba(oracle, d, lst)
    bal, s = gpu_ba(sk, d, lst, order=order)
    assert_same_trajectory(s, so)
    assert rel_param_diff(bal.parameters.toArray(), p.params) <= PARAM_RTOL


def robust_case(sk, oracle, kind, a, **opts):
    d = synth.make_bal("small", seed=8)
    d.observations[::37] += 25.0                                  # gross outliers
    loss = sk.PredefinedLossFunctions.cauchyLoss(a) if kind == _abi.LOSS_CAUCHY else sk.PredefinedLossFunctions.huberLoss(a)
    p, so = oracle_ba(oracle, d, _abi.DENSE_SCHUR, loss=(kind, a), **opts)
    bal, s = gpu_ba(sk, d, _abi.DENSE_SCHUR, loss=loss, **opts)
    return p, so, bal, s


@pytest.mark.parametrize("kind,a", [(_abi.LOSS_CAUCHY, 2.0), (_abi.LOSS_HUBER, 1.5)])
def test_ba_robust_loss_well_conditioned(sk, oracle, kind, a):
    """Strict: with the trust region capped at 1e12 the reduced system stays well conditioned and the whole LM
    trajectory (rows, radii, costs) must equal the oracle's."""
    p, so, bal, s = robust_case(sk, oracle, kind, a, max_trust_region_radius=1e12)
    assert all(r.step_is_valid for r in so.iterations)
    assert_same_trajectory(s, so)


@pytest.mark.parametrize("kind,a", [(_abi.LOSS_CAUCHY, 2.0), (_abi.LOSS_HUBER, 1.5)])
def test_ba_robust_loss_default_radius(sk, oracle, kind, a):
    """Default max_trust_region_radius = 1e16.  KNOWN DIVERGENCE: once the radius saturates,
    the LM damping is ~1e-16 of J'J and, the gauge being free, the reduced camera matrix is singular to working
    precision.  The oracle's Cholesky then reports a non-positive pivot (invalid step, radius halved) on every other
    iteration, the device factorisation does not: Huber(1.5) takes 46 rows in the oracle and 37 on the device, same
    steps, same final cost.  Rows are compared exactly only below the saturated radius."""
    p, so, bal, s = robust_case(sk, oracle, kind, a)
    assert s.termination_type == so.termination_type
    assert abs(s.final_cost - so.final_cost) <= COST_RTOL * abs(so.final_cost)
    rmax = 1e16
    n = 0
    for g, o in zip(s.iterations, so.iterations):
        if max(g.trust_region_radius, o.trust_region_radius) >= rmax:
            break
        assert (g.step_is_valid, g.step_is_successful) == (o.step_is_valid, o.step_is_successful)
        assert np.isclose(g.cost, o.cost, rtol=1e-6)
        n += 1
    assert n >= 10, "the comparable prefix of the trajectory is suspiciously short"


def test_ba_per_block_api_equals_bulk_api(sk):
    """Problem.addResidualBlock per observation (SimpleBundleAdjuster.scala:139-145) == the bulk call."""
    d = synth.make_bal("tiny", seed=9)
    bal = sk.BalProblem.fromArrays(d)
    loss = sk.PredefinedLossFunctions.trivialLoss()
    problem = sk.Problem()
    o = d.observations
    for i in range(d.num_observations):
        cost = sk.SnavelyReprojectionError(o[2 * i], o[2 * i + 1]).toAutoDiffCostFunction()
        problem.addResidualBlock(cost, loss, bal.mutableCameraForObservation(i), bal.mutablePointForObservation(i))
    opt = sk.Solver.Options()
    opt.setLinearSolverType(_abi.DENSE_SCHUR)
    s1 = sk.Solver.Summary()
    sk.ceres.solve(opt, problem, s1)
    bal2, s2 = gpu_ba(sk, d, _abi.DENSE_SCHUR)
    assert s1.final_cost == s2.final_cost and len(s1.iterations) == len(s2.iterations)
    assert np.array_equal(bal.parameters.toArray(), bal2.parameters.toArray())
    assert "DENSE_SCHUR" in s1.fullReport()


def test_rank_local_blocks_and_declared_cameras(sk):
    """sk_solver_options.residual_blocks_are_local on one GPU: declared cameras (Problem::AddParameterBlock) enter the
    camera table even without an observation.  An unobserved all-zero camera has zero gradient and takes no step, so the
    solve must be the one of the plain problem, bit for bit; without the flag a declared block is ignored, as in Ceres."""
    d = synth.make_bal("small", seed=2)
    bal0, s0 = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI)
    n = d.parameters.size
    for local in (1, 0):
        arr = sk.DoubleArray.fromArray(np.concatenate([d.parameters, np.zeros(9)]))      # the extra camera lives behind the points
        bal = sk.BalProblem(d.num_cameras, d.num_points, d.camera_index, d.point_index, d.observations, arr)
        problem = bal.buildLocalProblem(0, 1)
        problem.addParameterBlocks(arr, [n], 9)
        assert problem.numParameterBlocks() == d.num_cameras + d.num_points + 1
        o = sk.Solver.Options()
        o.setLinearSolverType(_abi.ITERATIVE_SCHUR); o.setPreconditionerType(_abi.SCHUR_JACOBI)
        o.residual_blocks_are_local = local
        s = sk.Solver.Summary()
        sk.ceres.solve(o, problem, s)
        assert [r.cost for r in s.iterations] == [r.cost for r in s0.iterations]
        assert [r.linear_solver_iterations for r in s.iterations] == [r.linear_solver_iterations for r in s0.iterations]
        x = arr.toArray()
        assert np.array_equal(x[:n], bal0.parameters.toArray()) and np.all(x[n:] == 0.0)
        assert s.num_parameter_blocks == d.num_cameras + d.num_points + local


def test_bal_file_round_trip(sk, tmp_path):
    """BalProblem.fromFile (SimpleBundleAdjuster.scala:37-77) on a file in the BAL text layout."""
    d = synth.make_bal("tiny", seed=10)
    path = tmp_path / "problem.txt"
    synth.write_bal_text(d, path)
    bal = sk.BalProblem.fromFile(path)
    assert (bal.numCameras, bal.numPoints, bal.numObservations) == (d.num_cameras, d.num_points, d.num_observations)
    assert np.array_equal(bal.cameraIndex, d.camera_index) and np.array_equal(bal.pointIndex, d.point_index)
    assert np.array_equal(bal.observations, d.observations) and np.array_equal(bal.parameters.toArray(), d.parameters)
    with pytest.raises(sk.SkeresError) as e:
        sk.BalProblem.fromFile(tmp_path / "missing.txt")
    assert e.value.status == _abi.ERR_IO


def test_unsupported_requests_fail_loudly(sk):
    d = synth.make_bal("tiny", seed=1)
    bal = sk.BalProblem.fromArrays(d)
    problem = bal.buildProblem()
    for setter in (lambda o: o.setLinearSolverType(_abi.CGNR), lambda o: o.setMinimizerType(_abi.LINE_SEARCH),
                   lambda o: (o.setLinearSolverType(_abi.ITERATIVE_SCHUR), o.setPreconditionerType(_abi.CLUSTER_JACOBI))):
        o = sk.Solver.Options()
        o.setLinearSolverType(_abi.DENSE_SCHUR)
        setter(o)
        with pytest.raises(sk.SkeresError) as e:
            sk.ceres.solve(o, problem, sk.Solver.Summary())
        assert e.value.status == _abi.ERR_UNSUPPORTED
    # a Schur solver on a non-BA problem
    m, c = sk.DoubleArray(1), sk.DoubleArray(1)
    p2 = sk.Problem()
    p2.addResidualBlock(sk.ExponentialResidual(1.0, 2.0).toAutoDiffCostFunction(), None, m.toPointer(), c.toPointer())
    o = sk.Solver.Options()
    o.setLinearSolverType(_abi.DENSE_SCHUR)
    with pytest.raises(sk.SkeresError):
        sk.ceres.solve(o, p2, sk.Solver.Summary())


def test_solver_is_deterministic_and_restartable(sk):
    """Atomic-free fixed-order reductions: two runs are bit-identical; a prepared solver can be re-run."""
    d = synth.make_bal("ladybug-49", seed=2)
    bal = sk.BalProblem.fromArrays(d)
    problem = bal.buildProblem()
    o = sk.Solver.Options()
    o.setLinearSolverType(_abi.ITERATIVE_SCHUR)
    o.setPreconditionerType(_abi.SCHUR_JACOBI)
    solver = sk.PreparedSolver(o, problem)
    s1 = solver.minimize()
    x1 = bal.parameters.toArray()
    bal.parameters.copyFrom(d.parameters)
    s2 = solver.minimize()
    x2 = bal.parameters.toArray()
    assert np.array_equal(x1, x2)
    assert [(r.cost, r.linear_solver_iterations) for r in s1.iterations] == [(r.cost, r.linear_solver_iterations) for r in s2.iterations]
    s3 = solver.minimize(max_num_iterations=2)                        # continues from the converged point
    assert len(s3.iterations) <= 3 and s3.initial_cost <= s1.final_cost * (1 + 1e-9)
    solver.close()


# --------------------------------------------------------------------------------------------------- config 4: batched curve fits
def test_batched_curve_fits_match_per_problem_oracle(sk, oracle):
    n = 3000
    x, y, truth = synth.make_curve_fit_batch(n, seed=5)
    xa, ya = sk.DoubleArray.fromArray(x), sk.DoubleArray.fromArray(y)
    mc = sk.DoubleArray(2 * n)
    o = sk.Solver.Options()
    o.setLinearSolverType(_abi.DENSE_QR)
    o.setMaxNumIterations(25)
    summary, ic, fc, it, tt = sk.curve_fit_batch_solve(o, xa, ya, mc)
    sol = mc.toArray().reshape(2, n)
    assert np.all(tt == _abi.CONVERGENCE)
    oo = _abi.default_options()
    oo.linear_solver_type, oo.max_num_iterations = _abi.DENSE_QR, 25
    for j in range(0, n, 7):
        p = oracle.OracleProblem(np.zeros(2))
        p.add_residual_blocks(_abi.FUNCTOR_EXPONENTIAL_RESIDUAL, np.stack([x[:, j], y[:, j]], 1), np.tile([0, 1], (x.shape[0], 1)))
        so = p.solve(oo)
        assert it[j] == so.num_successful_steps + so.num_unsuccessful_steps, j
        assert tt[j] == so.termination_type
        assert abs(ic[j] - so.initial_cost) <= 1e-12 * so.initial_cost and abs(fc[j] - so.final_cost) <= COST_RTOL * so.final_cost
        assert rel_param_diff(sol[:, j], p.params, 1e-3) <= PARAM_RTOL
    assert np.median(np.abs(sol[0] - truth[0])) < 0.02                 # and the fits recover the generating parameters
    assert summary.num_kernel_launches == summary.num_iterations       # one launch per LM iteration


def test_batched_equals_single_problem_path(sk):
    """The in-tree CurveFitting data as a batch of one equals the DENSE_QR single-problem solve."""
    d = load("curve_fitting_data.json")
    x1, s1, _ = curve_fit(sk)
    xa, ya = sk.DoubleArray.fromArray(np.array(d["x"])), sk.DoubleArray.fromArray(np.array(d["y"]))
    mc = sk.DoubleArray(2)
    o = sk.Solver.Options()
    o.setLinearSolverType(_abi.DENSE_QR)
    o.setMaxNumIterations(25)
    summary, ic, fc, it, tt = sk.curve_fit_batch_solve(o, xa, ya, mc)
    assert it[0] == 14 and tt[0] == _abi.CONVERGENCE
    assert np.allclose(mc.toArray(), x1, rtol=1e-9)
    assert abs(fc[0] - s1.final_cost) <= 1e-12 * s1.final_cost


# --------------------------------------------------------------------------------------------------- full BASELINE sizes: properties
@pytest.mark.parametrize("shape", ["venice-1778"])
def test_full_size_properties(sk, shape):
    """At BASELINE.json's full size the oracle is too slow; check size-independent properties instead:
    monotone cost on accepted steps, run-to-run bit determinism, and invariance to the order in which
    the residual blocks were added (exercises the sort + tiling at scale)."""
    d = synth.make_bal(shape, seed=1)
    bal, s = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI, max_num_iterations=4)
    rows = s.iterations
    assert len(rows) == 5 and s.termination_type == _abi.NO_CONVERGENCE
    costs = [r.cost for r in rows if r.step_is_successful]
    assert all(b < a for a, b in zip(costs, costs[1:]))
    assert rows[0].cost > 10 * rows[-1].cost
    assert s.num_residual_blocks == d.num_observations
    x1 = bal.parameters.toArray()
    bal2, s2 = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI, max_num_iterations=4)
    assert np.array_equal(x1, bal2.parameters.toArray())
    assert [r.cost for r in s2.iterations] == [r.cost for r in rows]
    order = np.random.default_rng(0).permutation(d.num_observations)
    bal3, s3 = gpu_ba(sk, d, _abi.ITERATIVE_SCHUR, _abi.SCHUR_JACOBI, order=order, max_num_iterations=4)
    assert [r.linear_solver_iterations for r in s3.iterations] == [r.linear_solver_iterations for r in rows]
    assert np.array_equal(x1, bal3.parameters.toArray())               # the internal order is canonical


def test_batched_curve_fits_full_size_properties(sk):
    """BASELINE.json configs[3] at its full size: 1M independent CurveFitting-sized problems, one launch per LM iteration.
    Size-independent properties: every problem terminates with CONVERGENCE, no cost rises, the fitted (m, c) scatter
    around the truth as the noise level predicts, and the run is bit-reproducible."""
    n = 1_000_000
    x, y, truth = synth.make_curve_fit_batch(n, seed=1)
    xa, ya = sk.DoubleArray.fromArray(x), sk.DoubleArray.fromArray(y)
    o = sk.Solver.Options()
    o.setLinearSolverType(_abi.DENSE_QR)
    o.setMaxNumIterations(25)
    runs = []
    for _ in range(2):
        mc = sk.DoubleArray(2 * n)
        summary, ic, fc, it, tt = sk.curve_fit_batch_solve(o, xa, ya, mc)
        runs.append((mc.toArray(), fc.copy(), it.copy()))
        assert np.all(tt == _abi.CONVERGENCE)
        assert np.all(fc <= ic)
    sol = runs[0][0].reshape(2, n)
    err = sol - truth
    assert abs(err[0].mean()) < 1e-3 and abs(err[1].mean()) < 3e-3           # unbiased up to the nonlinearity
    assert 0.005 < err[0].std() < 0.08 and 0.01 < err[1].std() < 0.25
    assert np.median(runs[0][1]) == pytest.approx(0.5 * 67 * 0.2 ** 2, rel=0.1)   # residual cost ~ (n_obs - 2) sigma^2 / 2
    assert 1 <= runs[0][2].min() and runs[0][2].max() <= 25
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][2], runs[1][2])
