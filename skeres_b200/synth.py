"""Deterministic synthetic inputs for the hot path.

BAL-shaped bundle-adjustment problems in the exact layout `BalProblem` uses in the reference
(examples/.../SimpleBundleAdjuster.scala:18-77): observation i is (camera_index[i], point_index[i],
x, y); the parameter vector is one contiguous double[9*n_cam + 3*n_pt], cameras first (:28-33,
:49-50); a camera is [angle-axis 3, translation 3, focal, k1, k2] in the Snavely convention
(negative z in front of the camera, :99-103).

Geometry: cameras on a ring in the x-z plane looking at the axis, points inside the ring ordered by
angle, each point seen by a short run of consecutive cameras (a vehicle-sequence-like track, as in
the Ladybug BAL problems).  Camera 0 has a (numerically) zero angle-axis vector so the Taylor
branch of Rotation.angleAxisRotatePoint (Rotation.scala:492-520) is exercised.  Observations are
sorted by point, then camera.

Also: the batched CurveFitting-shaped problems of BASELINE.json configs[3]
(CurveFitting.scala:13-20 describes how the in-tree data was generated).
"""
from dataclasses import dataclass

import numpy as np

SHAPES = {
    # name: (n_cam, n_pt, n_obs)   — BASELINE.json configs
    "ladybug-49": (49, 7776, 31843),
    "venice-1778": (1778, 993923, 5001946),
    "final-13682": (13682, 4456117, 28987644),
    "tiny": (6, 60, 240),
    "small": (16, 600, 2700),
}


@dataclass
class BalData:
    num_cameras: int
    num_points: int
    camera_index: np.ndarray   # int32 [n_obs]
    point_index: np.ndarray    # int32 [n_obs]
    observations: np.ndarray   # float64 [2 * n_obs]
    parameters: np.ndarray     # float64 [9 * n_cam + 3 * n_pt]  (perturbed start point)
    ground_truth: np.ndarray   # float64, same layout

    @property
    def num_observations(self):
        return int(self.camera_index.size)

    def block_offsets(self):
        """(n_obs, 2) int64: offset of the camera block and of the point block of each observation
        (mutableCameraForObservation / mutablePointForObservation, SimpleBundleAdjuster.scala:31-33)."""
        off = np.empty((self.num_observations, 2), dtype=np.int64)
        off[:, 0] = self.camera_index.astype(np.int64) * 9
        off[:, 1] = 9 * self.num_cameras + self.point_index.astype(np.int64) * 3
        return off


def rotate_points(aa, X):
    """Rodrigues rotation, vectorised: aa (n,3), X (n,3)."""
    th2 = np.einsum("ij,ij->i", aa, aa)
    th = np.sqrt(np.maximum(th2, 1e-300))
    w = aa / th[:, None]
    c, s = np.cos(th)[:, None], np.sin(th)[:, None]
    big = (th2 > np.finfo(float).eps)[:, None]
    rod = X * c + np.cross(w, X) * s + w * np.einsum("ij,ij->i", w, X)[:, None] * (1 - c)
    tay = X + np.cross(aa, X)
    return np.where(big, rod, tay)


def project(cams, X):
    """Snavely projection (SimpleBundleAdjuster.scala:91-114) for generating observations."""
    p = rotate_points(cams[:, 0:3], X) + cams[:, 3:6]
    xp = -p[:, 0] / p[:, 2]
    yp = -p[:, 1] / p[:, 2]
    r2 = xp * xp + yp * yp
    d = 1.0 + r2 * (cams[:, 7] + cams[:, 8] * r2)
    return np.stack([cams[:, 6] * d * xp, cams[:, 6] * d * yp], axis=1), p[:, 2]


def _track_lengths(rng, n_pt, n_obs, kmax):
    mean_extra = n_obs / n_pt - 2.0
    assert mean_extra > 0, "need more than 2 observations per point on average"
    k = 2 + rng.geometric(1.0 / (1.0 + mean_extra), size=n_pt) - 1
    k = np.minimum(k, kmax).astype(np.int64)
    # adjust to hit n_obs exactly
    for _ in range(200):
        diff = n_obs - int(k.sum())
        if diff == 0:
            break
        idx = rng.integers(0, n_pt, size=min(abs(diff), n_pt))
        idx = np.unique(idx)
        if diff > 0:
            idx = idx[k[idx] < kmax][:diff]
            k[idx] += 1
        else:
            idx = idx[k[idx] > 2][:(-diff)]
            k[idx] -= 1
    assert int(k.sum()) == n_obs, (int(k.sum()), n_obs)
    return k


def make_bal(shape="ladybug-49", seed=1, noise_px=0.5, point_sigma=0.05, rot_sigma=1e-3,
             trans_sigma=1e-2, n_cam=None, n_pt=None, n_obs=None, long_tracks=()) -> BalData:
    """long_tracks: track lengths (<= n_cam) forced onto evenly spaced points, on top of n_obs -- real BAL files hold a
    few points seen by hundreds of cameras, which the device layout cuts into chunk tiles (ba_layout.h)."""
    if n_cam is None:
        n_cam, n_pt, n_obs = SHAPES[shape]
    rng = np.random.default_rng(seed)
    # --- cameras on a ring of radius 30 in the x-z plane, looking at the y axis -------------------
    theta = 2.0 * np.pi * np.arange(n_cam) / n_cam
    rho_c = 30.0 + rng.normal(0, 0.5, n_cam)
    centers = np.stack([rho_c * np.sin(theta), rng.normal(0, 0.5, n_cam), rho_c * np.cos(theta)], axis=1)
    # world->camera rotation is a rotation about y by -theta: angle-axis (0, -theta, 0); at theta = 0
    # it is the identity, i.e. a zero angle-axis vector (Taylor branch).
    aa = np.zeros((n_cam, 3))
    aa[:, 1] = -theta
    aa[theta > np.pi, 1] += 2.0 * np.pi        # keep |angle| <= pi
    aa[1:] += rng.normal(0, 0.02, (n_cam - 1, 3))
    aa[0] = rng.normal(0, 1e-10, 3)            # ||aa||^2 << ulp(1.0)
    t = -rotate_points(aa, centers)
    f = rng.uniform(500.0, 2000.0, n_cam)
    k1 = rng.normal(0, 1e-2, n_cam) * 1e-1
    k2 = rng.normal(0, 1e-3, n_cam) * 1e-1
    cams = np.concatenate([aa, t, f[:, None], k1[:, None], k2[:, None]], axis=1)
    # --- points inside the ring, ordered by angle -------------------------------------------------
    phi = 2.0 * np.pi * (np.arange(n_pt) + rng.uniform(-0.5, 0.5, n_pt)) / n_pt
    rho_p = rng.uniform(4.0, 14.0, n_pt)
    pts = np.stack([rho_p * np.sin(phi), rng.uniform(-3.0, 3.0, n_pt), rho_p * np.cos(phi)], axis=1)
    # --- visibility: a run of consecutive cameras around the nearest one --------------------------
    kmax = int(min(n_cam, max(6, np.ceil(3.0 * n_obs / n_pt))))
    k = _track_lengths(rng, n_pt, n_obs, kmax)
    for j, length in enumerate(long_tracks):
        assert 2 <= length <= n_cam
        k[(j + 1) * n_pt // (len(long_tracks) + 1)] = length
    n_obs = int(k.sum())
    nearest = np.floor(phi / (2.0 * np.pi) * n_cam + 0.5).astype(np.int64)
    base = nearest + rng.integers(-2, 3, n_pt) - k // 2
    base = np.clip(base, 0, n_cam - k)
    ptr = np.zeros(n_pt + 1, dtype=np.int64)
    np.cumsum(k, out=ptr[1:])
    point_index = np.repeat(np.arange(n_pt, dtype=np.int64), k)
    within = np.arange(n_obs, dtype=np.int64) - ptr[point_index]
    camera_index = base[point_index] + within
    # --- observations = exact projection + noise --------------------------------------------------
    obs, depth = project(cams[camera_index], pts[point_index])
    assert np.all(depth < -1.0), "a point is not in front of its camera"
    obs = obs + rng.normal(0, noise_px, obs.shape)
    # --- perturbed start ---------------------------------------------------------------------------
    cams0 = cams.copy()
    cams0[1:, 0:3] += rng.normal(0, rot_sigma, (n_cam - 1, 3))
    cams0[:, 3:6] += rng.normal(0, trans_sigma, (n_cam, 3))
    cams0[:, 6] *= 1.0 + rng.normal(0, 1e-3, n_cam)
    pts0 = pts + rng.normal(0, point_sigma, pts.shape)
    gt = np.concatenate([cams.ravel(), pts.ravel()])
    x0 = np.concatenate([cams0.ravel(), pts0.ravel()])
    return BalData(n_cam, n_pt, camera_index.astype(np.int32), point_index.astype(np.int32),
                   np.ascontiguousarray(obs.ravel()), x0, gt)


def make_scene_share(rank, world, shape="venice-1778", seed=1) -> BalData:
    """Rank `rank`'s share of a problem made of `world` DISJOINT copies of one scene (bench.py's weak-scaling workload): scene r
    owns cameras [r C, (r + 1) C) and points [r P, (r + 1) P), so the point partition (contiguous, balanced by observation count)
    gives every rank exactly one scene.  Because the copies are identical and do not share parameters, the LM trajectory and the
    PCG iteration counts of the whole problem are those of one scene: the work per step is `world` times one scene's, step by
    step -- a weak-scaling workload whose per-step work does not change with the number of GPUs.
    Returns global camera / point indices for the local observations only, and the FULL start vector (cameras of every scene,
    points of every scene), as the rank-local ingestion of the multi-GPU path expects."""
    one = make_bal(shape, seed=seed)
    C_, P_ = one.num_cameras, one.num_points
    cams = np.tile(one.parameters[:9 * C_], world)
    pts = np.tile(one.parameters[9 * C_:], world)
    gt = np.concatenate([np.tile(one.ground_truth[:9 * C_], world), np.tile(one.ground_truth[9 * C_:], world)])
    return BalData(C_ * world, P_ * world, (one.camera_index + rank * C_).astype(np.int32), (one.point_index + rank * P_).astype(np.int32),
                   one.observations, np.concatenate([cams, pts]), gt)


def write_bal_text(data: BalData, path):
    """BAL text layout parsed by BalProblem.fromFile (SimpleBundleAdjuster.scala:41-62)."""
    with open(path, "w") as fh:
        fh.write(f"{data.num_cameras} {data.num_points} {data.num_observations}\n")
        o = data.observations.reshape(-1, 2)
        for i in range(data.num_observations):
            fh.write(f"{data.camera_index[i]} {data.point_index[i]} {o[i, 0]:.17e} {o[i, 1]:.17e}\n")
        for v in data.parameters:
            fh.write(f"{v:.17e}\n")


def make_curve_fit_batch(n_problems, seed=1, n_obs=67, sigma=0.2):
    """BASELINE.json configs[3]: x = 0:0.075:4.95 (CurveFitting.scala:16), y = exp(m x + c) + noise.
    Returns x, y as [n_obs][n_problems] (observation-major SoA) and the true (m, c) [2][n]."""
    rng = np.random.default_rng(seed)
    m = rng.uniform(0.1, 0.5, n_problems)
    c = rng.uniform(-0.2, 0.4, n_problems)
    xs = 0.075 * np.arange(n_obs)
    x = np.repeat(xs[:, None], n_problems, axis=1)
    y = np.exp(m[None, :] * x + c[None, :]) + rng.normal(0, sigma, (n_obs, n_problems))
    return np.ascontiguousarray(x), np.ascontiguousarray(y), np.stack([m, c])
