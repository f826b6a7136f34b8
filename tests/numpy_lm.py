"""An INDEPENDENT numpy statement of the solver loop the oracle restates in C++: Ceres' trust-region
minimizer with the Levenberg-Marquardt strategy, and its conjugate-gradients solver on the reduced camera system with the
SchurJacobi preconditioner -- written against dense numpy linear algebra (the full Jacobian, the explicit Schur complement
as a matrix), sharing nothing with oracle/oracle.cc but the residual / Jacobian evaluation of the functor (which the
reference's own spec vectors and finite differences pin).  Test infrastructure only; small problems only (dense J).

What it follows (Ceres 1.x, as published): TrustRegionMinimizer::Minimize (Jacobi scaling by 1 / (1 + ||J_j||) fixed at x0;
step quality rho = cost change / model cost change; accept when rho > min_relative_decrease; function / gradient / parameter
tolerance tests), LevenbergMarquardtStrategy (D^2 = clamp(diag(J'J)) / radius; on success radius /= max(1/3, 1 - (2 rho - 1)^3),
on failure radius /= decrease_factor and decrease_factor *= 2), ConjugateGradientsSolver::Solve (r recomputed from b - A x every
10th iteration; termination on zeta = i (Q_i - Q_{i-1}) / Q_i < q_tolerance = eta), SchurJacobiPreconditioner (inverse of the
block diagonal of S)."""
import numpy as np


def dense_jacobian(d, Jv):
    """oracle evaluate() returns per residual block [2 x 9 | 2 x 3] row-major; scatter into a dense 2 n_obs x n matrix."""
    n = d.parameters.size
    J = np.zeros((2 * d.num_observations, n))
    off = d.block_offsets()
    Jv = Jv.reshape(-1, 24)
    for i in range(d.num_observations):
        J[2 * i:2 * i + 2, off[i, 0]:off[i, 0] + 9] = Jv[i, :18].reshape(2, 9)
        J[2 * i:2 * i + 2, off[i, 1]:off[i, 1] + 3] = Jv[i, 18:].reshape(2, 3)
    return J


def pcg_on_schur_complement(S, b, blocks, eta, max_it=500, min_it=0, reset_period=10, preconditioner="schur_jacobi", FtF=None):
    """ConjugateGradientsSolver::Solve on S x = b.  M^-1 = inverse of the 9 x 9 diagonal blocks of S (SCHUR_JACOBI) or of
    FtF (JACOBI: block_diagonal_FtF_inverse).  Returns (x, iterations)."""
    n = b.size
    Minv = np.zeros((n, n))
    src = S if preconditioner == "schur_jacobi" else FtF
    for c in range(blocks):
        sl = slice(9 * c, 9 * c + 9)
        Minv[sl, sl] = np.linalg.inv(src[sl, sl])
    x = np.zeros(n)
    r = b.copy()
    norm_b = np.linalg.norm(b)
    if norm_b == 0.0:
        return x, 0
    rho = 1.0
    Q0 = -np.dot(x, b + r)
    p = np.zeros(n)
    it = 0
    while True:
        it += 1
        z = Minv @ r
        last_rho, rho = rho, float(r @ z)
        if it == 1:
            p = z.copy()
        else:
            p = z + (rho / last_rho) * p
        q = S @ p
        pq = float(p @ q)
        if not (pq > 0.0):
            break
        alpha = rho / pq
        x = x + alpha * p
        r = (b - S @ x) if (it % reset_period == 0) else (r - alpha * q)
        Q1 = -float(x @ (b + r))
        zeta = it * (Q1 - Q0) / Q1
        if zeta < eta and it >= min_it:
            break
        Q0 = Q1
        if it >= max_it:
            break
    return x, it


def solve(d, evaluate, linear="dense", eta=0.1, max_num_iterations=50, preconditioner="schur_jacobi"):
    """evaluate(x) -> (cost, residuals, dense J).  linear = "dense": the damped normal equations solved directly;
    "pcg": points eliminated explicitly, reduced camera system solved by pcg_on_schur_complement, points back-substituted.
    Returns the rows [(cost, radius, successful, linear_iterations)], row 0 = the start, and the final parameters."""
    x = d.parameters.copy()
    nc = 9 * d.num_cameras
    radius, decrease_factor = 1e4, 2.0
    cost, r, J = evaluate(x)
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))
    rows = [(cost, radius, True, 0)]
    x_norm = np.linalg.norm(x)
    g = J.T @ r
    if np.max(np.abs(g)) <= 1e-10:
        return rows, x
    for _ in range(max_num_iterations):
        Js = J * scale
        D2 = np.clip((Js * Js).sum(0), 1e-6, 1e32) / radius
        gs = Js.T @ r
        if linear == "dense":
            delta = np.linalg.solve(Js.T @ Js + np.diag(D2), -gs)
            lin_its = 1
        else:
            H = Js.T @ Js + np.diag(D2)
            A, B, Cm = H[:nc, :nc], H[:nc, nc:], H[nc:, nc:]            # cameras, coupling, points (block diagonal 3 x 3)
            Cinv = np.zeros_like(Cm)
            for p_ in range(d.num_points):
                sl = slice(3 * p_, 3 * p_ + 3)
                Cinv[sl, sl] = np.linalg.inv(Cm[sl, sl])
            S = A - B @ Cinv @ B.T
            rhs = -gs[:nc] + B @ (Cinv @ gs[nc:])
            yc, lin_its = pcg_on_schur_complement(S, rhs, d.num_cameras, eta, preconditioner=preconditioner, FtF=A)
            yp = Cinv @ (-gs[nc:] - B.T @ yc)
            delta = np.concatenate([yc, yp])
        model = Js @ delta
        model_cost_change = -float(model @ (r + model / 2.0))
        if not (model_cost_change > 0.0):                      # invalid step (HandleInvalidStep): shrink and try again
            radius /= decrease_factor
            decrease_factor *= 2.0
            rows.append((cost, radius, False, lin_its))
            continue
        xc = x + delta * scale
        cand_cost, rc, Jc = evaluate(xc)
        # Ceres >= 1.12 (TrustRegionMinimizer::Minimize): the parameter and function tolerance tests look at the CANDIDATE and
        # return before the step is applied or the iteration's row is recorded.
        step_norm = np.linalg.norm(delta * scale)
        if step_norm <= 1e-8 * (x_norm + 1e-8):
            break
        if abs(cost - cand_cost) <= 1e-6 * cost:
            break
        rel = (cost - cand_cost) / model_cost_change
        if rel > 1e-3:
            x, cost, r, J = xc, cand_cost, rc, Jc
            x_norm = np.linalg.norm(x)
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rel - 1.0) ** 3))
            decrease_factor = 2.0
            rows.append((cost, radius, True, lin_its))
            if np.max(np.abs(J.T @ r)) <= 1e-10:
                break
        else:
            radius /= decrease_factor
            decrease_factor *= 2.0
            rows.append((cand_cost, radius, False, lin_its))
            if radius < 1e-32:
                break
    return rows, x
