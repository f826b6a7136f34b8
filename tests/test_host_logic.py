"""CPU-side checks of the product's host logic (no GPU, no compute calls on the device):
  * libskeres.so loads and exports every symbol include/skeres.h declares;
  * option defaults, functor registry, error behaviour without a device;
  * the tiled bundle-adjustment layout builder and the point partition (ba_layout.cu) against a numpy restatement;
  * the __host__ __device__ arithmetic of jet.cuh (built for the host by tests/hostcheck) against the oracle.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from skeres_b200 import _abi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# --------------------------------------------------------------------------------------------------- ABI
def declared_functions():
    src = open(os.path.join(ROOT, "include", "skeres.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from skeres_b200 import _lib
    names = declared_functions()
    assert len(names) > 60
    for n in names:
        assert hasattr(_lib.lib, n), f"libskeres.so does not export {n}"
    assert set(_lib.DECLARED_SYMBOLS) == set(names), set(_lib.DECLARED_SYMBOLS) ^ set(names)
    assert _lib.lib.sk_abi_version() == _abi.ABI_VERSION


def test_no_torch_or_oracle_in_the_product_library():
    out = subprocess.run(["ldd", os.path.join(ROOT, "skeres_b200", "libskeres.so")], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "ceres" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "skeres_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "import oracle" not in text and "oracle/" not in text, f"{f} references the oracle"


def test_option_defaults_match_ceres():
    from skeres_b200 import _lib
    o = _abi.SolverOptions()
    _lib.lib.sk_solver_options_init(C.byref(o))
    d = _abi.default_options()
    for name, _ in _abi.SolverOptions._fields_:
        assert getattr(o, name) == getattr(d, name), name
    assert (o.max_num_iterations, o.initial_trust_region_radius, o.min_relative_decrease) == (50, 1e4, 1e-3)
    assert (o.function_tolerance, o.gradient_tolerance, o.parameter_tolerance, o.eta) == (1e-6, 1e-10, 1e-8, 1e-1)
    assert (o.max_linear_solver_iterations, o.min_lm_diagonal, o.max_lm_diagonal) == (500, 1e-6, 1e32)


def test_functor_registry_and_unregistered_functor():
    from skeres_b200 import api
    assert api.functor_info(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR) == (2, [9, 3], 2)      # SimpleBundleAdjuster.scala:79
    assert api.functor_info(_abi.FUNCTOR_EXPONENTIAL_RESIDUAL) == (1, [1, 1], 2)            # CurveFitting.scala:92
    assert api.functor_info(_abi.FUNCTOR_TEST_SUM10) == (1, [1] * 10, 0)
    with pytest.raises(api.SkeresError) as e:
        api.functor_info(999)
    assert e.value.status == _abi.ERR_UNSUPPORTED and "no CPU fallback" in str(e.value)

    class Arbitrary(api.AutoDiffCostFunctor):        # an arbitrary closure cannot run on the device
        pass
    with pytest.raises(api.SkeresError):
        Arbitrary(1, 2).toAutoDiffCostFunction()
    for factory in (lambda: api.PredefinedLossFunctions.tukeyLoss(1.0), lambda: api.PredefinedLossFunctions.softLOneLoss(1.0)):
        with pytest.raises(api.SkeresError) as e:
            factory()
        assert e.value.status == _abi.ERR_UNSUPPORTED
    tol = api.PredefinedLossFunctions.tolerantLoss(1.0, 2.0)        # registered (a host-side handle: no GPU needed to create it)
    assert (tol.kind, tol.a, tol.b) == (_abi.LOSS_TOLERANT, 1.0, 2.0)
    for bad in ((1.0, 0.0), (-0.5, 1.0)):                             # Ceres CHECKs a >= 0, b > 0
        with pytest.raises(api.SkeresError) as e:
            api.PredefinedLossFunctions.tolerantLoss(*bad)
        assert e.value.status == _abi.ERR_INVALID_ARGUMENT


def test_fails_loudly_without_a_gpu():
    from skeres_b200 import api
    if api.lib.sk_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.SkeresError) as e:
        api.DoubleArray(8)
    assert e.value.status == _abi.ERR_CUDA and "no CPU fallback" in str(e.value)


# --------------------------------------------------------------------------------------------------- host check harness
@pytest.fixture(scope="module")
def hc():
    d = os.path.join(ROOT, "tests", "hostcheck")
    subprocess.run(["make", "-C", d], check=True, capture_output=True)
    L = C.CDLL(os.path.join(d, "libhostcheck.so"))
    L.hc_layout_build.restype = C.c_void_p
    L.hc_layout_build.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_char_p]
    L.hc_layout_free.argtypes = [C.c_void_p]
    L.hc_evaluate.argtypes = [C.c_int] + [C.c_void_p] * 4
    L.hc_snavely_residual.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
    L.hc_loss.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]
    L.hc_correct.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.hc_invert_spd3.argtypes = [C.c_void_p, C.c_void_p]
    L.hc_invert_spd9.argtypes = [C.c_void_p, C.c_int]
    return L


def p(a):
    return a.ctypes.data_as(C.c_void_p)


class Layout:
    FIELDS = {"perm": np.int32, "obs_cam": np.int32, "obs_pt": np.int32, "pt_ptr": np.int32, "tile_obs": np.int32, "tile_pt": np.int32,
              "tile_seg": np.int32, "obs_slot": np.uint16, "obs_ptl": np.uint16, "seg_perm": np.uint16, "seg_ptr": np.int32,
              "seg_cam": np.int32, "cam_seg_ptr": np.int32, "cam_seg": np.int32, "cam_offset": np.int64, "pt_offset": np.int64, "obs": np.float64,
              "tile_np": np.int32, "tile_chunk": np.int32, "gp_tile_begin": np.int32, "gp_tile_count": np.int32, "gp_point": np.int32}

    def __init__(self, L, cam_off, pt_off, obs, rank=0, world=1, extra_cams=None):
        cam_off = np.ascontiguousarray(cam_off, dtype=np.int64); pt_off = np.ascontiguousarray(pt_off, dtype=np.int64)
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        st = C.c_int(); err = C.create_string_buffer(256)
        if extra_cams is None:
            h = L.hc_layout_build(cam_off.size, p(cam_off), p(pt_off), p(obs), rank, world, C.byref(st), err)
        else:
            ex = np.ascontiguousarray(extra_cams, dtype=np.int64)
            L.hc_layout_build_extra.restype = C.c_void_p
            L.hc_layout_build_extra.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int), C.c_char_p]
            h = L.hc_layout_build_extra(cam_off.size, p(cam_off), p(pt_off), p(obs), ex.size, p(ex), C.byref(st), err)
        self.status, self.error = st.value, err.value.decode()
        if not h:
            return
        dims = np.zeros(8, dtype=np.int32)
        L.hc_layout_dims(C.c_void_p(h), p(dims))
        self.n_obs, self.n_pts, self.n_cams, self.n_tiles, self.n_segs, self.max_seg_tile, self.max_pt_tile, self.sorted = dims.tolist()
        L.hc_layout_n_giant.argtypes = [C.c_void_p]
        self.n_giant = L.hc_layout_n_giant(C.c_void_p(h))
        L.hc_layout_n_chunks.argtypes = [C.c_void_p]
        self.n_chunks = L.hc_layout_n_chunks(C.c_void_p(h))
        size = {"perm": self.n_obs, "obs_cam": self.n_obs, "obs_pt": self.n_obs, "pt_ptr": self.n_pts + 1, "tile_obs": self.n_tiles + 1,
                "tile_pt": self.n_tiles + 1, "tile_seg": self.n_tiles + 1, "obs_slot": self.n_obs, "obs_ptl": self.n_obs, "seg_perm": self.n_obs,
                "seg_ptr": self.n_segs + 1, "seg_cam": self.n_segs, "cam_seg_ptr": self.n_cams + 1, "cam_seg": self.n_segs,
                "cam_offset": self.n_cams, "pt_offset": self.n_pts, "obs": 2 * self.n_obs, "tile_np": self.n_tiles, "tile_chunk": self.n_tiles,
                "gp_tile_begin": self.n_giant, "gp_tile_count": self.n_giant, "gp_point": self.n_giant}
        for name, dt in self.FIELDS.items():
            a = np.zeros(size[name], dtype=dt)
            fn = getattr(L, "hc_layout_" + name)
            fn.argtypes = [C.c_void_p, C.c_void_p]
            fn(C.c_void_p(h), p(a))
            setattr(self, name, a)
        L.hc_layout_free(C.c_void_p(h))


def check_layout(lay, cam_off, pt_off, obs, pt_range=None):
    """Every invariant the tile kernels rely on."""
    cams = np.unique(cam_off); pts_all = np.unique(pt_off)
    pts = pts_all if pt_range is None else pts_all[pt_range[0]:pt_range[1]]
    assert lay.n_cams == cams.size and np.array_equal(lay.cam_offset, cams)
    assert lay.n_pts == pts.size and np.array_equal(lay.pt_offset, pts)
    # sorted by (point, camera); perm maps back to the original residual blocks
    key = lay.obs_pt.astype(np.int64) * (lay.n_cams + 1) + lay.obs_cam
    assert np.all(np.diff(key) > 0)
    assert np.array_equal(cam_off[lay.perm], cams[lay.obs_cam]) and np.array_equal(pt_off[lay.perm], pts[lay.obs_pt])
    assert np.array_equal(obs.reshape(-1, 2)[lay.perm].ravel(), lay.obs)
    assert np.array_equal(np.bincount(lay.obs_pt, minlength=lay.n_pts), np.diff(lay.pt_ptr))
    # tiles: <= 256 observations, consistent point / segment ranges; regular tiles hold whole points, a track longer
    # than 256 observations is a chain of consecutive chunk tiles that together cover exactly that point
    assert lay.tile_obs[0] == 0 and lay.tile_obs[-1] == lay.n_obs and lay.tile_pt[-1] == lay.n_pts and lay.tile_seg[-1] == lay.n_segs
    assert np.all(np.diff(lay.tile_obs) <= 256) and np.all(np.diff(lay.tile_obs) > 0)
    assert lay.max_pt_tile == lay.tile_np.max() and lay.max_seg_tile == np.diff(lay.tile_seg).max()
    track = np.diff(lay.pt_ptr)
    assert np.array_equal(np.sort(lay.gp_point), np.nonzero(track > 256)[0]) and lay.n_giant == int((track > 256).sum())
    covered = np.zeros(lay.n_pts, dtype=int)
    for t in range(lay.n_tiles):
        if lay.tile_chunk[t] >= 0:
            assert lay.tile_np[t] == 1 and track[lay.tile_pt[t]] > 256
            assert np.all(lay.obs_pt[lay.tile_obs[t]:lay.tile_obs[t + 1]] == lay.tile_pt[t])
        else:
            pts = np.arange(lay.tile_pt[t], lay.tile_pt[t] + lay.tile_np[t])
            covered[pts] += 1
            assert np.all(track[pts] <= 256)
            assert lay.pt_ptr[pts[0]] == lay.tile_obs[t] and lay.pt_ptr[pts[-1] + 1] == lay.tile_obs[t + 1]      # whole points
    for g in range(lay.n_giant):
        p_, b_, c_ = lay.gp_point[g], lay.gp_tile_begin[g], lay.gp_tile_count[g]
        covered[p_] += 1
        assert c_ == -(-track[p_] // 256) and np.all(lay.tile_chunk[b_:b_ + c_] >= 0) and np.all(lay.tile_pt[b_:b_ + c_] == p_)
        assert lay.tile_obs[b_] == lay.pt_ptr[p_] and lay.tile_obs[b_ + c_] == lay.pt_ptr[p_ + 1]                  # chunks tile the track
        assert np.diff(lay.tile_obs[b_:b_ + c_ + 1]).max() - np.diff(lay.tile_obs[b_:b_ + c_ + 1]).min() <= 1      # balanced
    chunk_ids = lay.tile_chunk[lay.tile_chunk >= 0]
    assert np.array_equal(chunk_ids, np.arange(lay.n_chunks)) and lay.n_chunks == int(lay.gp_tile_count.sum())   # compact, in tile order
    assert np.all(covered == 1)                                                                                   # every point exactly once
    for t in range(lay.n_tiles):
        ob, oe, pb, sb, se = lay.tile_obs[t], lay.tile_obs[t + 1], lay.tile_pt[t], lay.tile_seg[t], lay.tile_seg[t + 1]
        assert np.array_equal(lay.obs_ptl[ob:oe], lay.obs_pt[ob:oe] - pb)
        assert np.array_equal(lay.seg_cam[sb:se], np.unique(lay.obs_cam[ob:oe]))          # one segment per distinct camera, ascending
        assert np.array_equal(lay.seg_cam[sb + lay.obs_slot[ob:oe].astype(int)], lay.obs_cam[ob:oe])
        assert lay.seg_ptr[sb] == ob and lay.seg_ptr[se] == oe
        assert sorted(lay.seg_perm[ob:oe].tolist()) == list(range(oe - ob))              # a permutation of the tile's observations
        for s in range(sb, se):
            loc = lay.seg_perm[lay.seg_ptr[s]:lay.seg_ptr[s + 1]].astype(int)
            assert np.all(lay.obs_cam[ob + loc] == lay.seg_cam[s]) and np.all(np.diff(loc) > 0)   # fixed (point) order inside a segment
    # camera -> segments in tile order, every segment exactly once
    assert sorted(lay.cam_seg.tolist()) == list(range(lay.n_segs))
    for c in range(lay.n_cams):
        segs = lay.cam_seg[lay.cam_seg_ptr[c]:lay.cam_seg_ptr[c + 1]]
        assert np.all(lay.seg_cam[segs] == c) and np.all(np.diff(segs) > 0)


@pytest.mark.parametrize("shape,seed", [("tiny", 1), ("small", 2), ("ladybug-49", 1)])
def test_layout_invariants(hc, shape, seed):
    d = synth.make_bal(shape, seed=seed)
    off = d.block_offsets()
    lay = Layout(hc, off[:, 0], off[:, 1], d.observations)
    assert lay.status == 0 and lay.sorted == 1
    check_layout(lay, off[:, 0], off[:, 1], d.observations)


def test_layout_sorts_shuffled_input_and_handles_arbitrary_offsets(hc):
    d = synth.make_bal("small", seed=5)
    off = d.block_offsets()
    perm = np.random.default_rng(1).permutation(d.num_observations)
    cam_off = off[perm, 0] * 7 + 1000003            # sparse, non-BAL offsets -> exercises the sort + binary-search id path
    pt_off = off[perm, 1] * 5 + 900000007
    obs = d.observations.reshape(-1, 2)[perm].ravel()
    lay = Layout(hc, cam_off, pt_off, obs)
    assert lay.status == 0 and lay.sorted == 0
    check_layout(lay, cam_off, pt_off, obs)


def test_layout_rejects_bad_structure(hc):
    d = synth.make_bal("tiny", seed=1)
    off = d.block_offsets()
    dup = np.concatenate([off, off[:1]])                               # same (camera, point) twice
    lay = Layout(hc, dup[:, 0], dup[:, 1], np.concatenate([d.observations, d.observations[:2]]))
    assert lay.status == _abi.ERR_UNSUPPORTED and "twice" in lay.error
    bad = off.copy(); bad[0, 0] += 4                                   # overlapping camera blocks
    lay = Layout(hc, bad[:, 0], bad[:, 1], d.observations)
    assert lay.status == _abi.ERR_INVALID_ARGUMENT and "overlap" in lay.error


def test_long_tracks_become_chunk_tiles(hc):
    """Tracks longer than one tile (real BAL files have them) are cut into balanced chunk tiles; everything else about
    the layout (segments, slots, camera lists) holds for chunk tiles exactly as for regular ones."""
    rng = np.random.default_rng(7)
    n_cam = 700
    lens = np.concatenate([rng.integers(2, 9, 40), [257], rng.integers(2, 9, 30), [256, 700, 513], rng.integers(2, 9, 50), [300]])
    cam, pt = [], []
    for j, k in enumerate(lens):
        cam.extend(sorted(rng.choice(n_cam, size=k, replace=False))); pt.extend([j] * k)
    cam_off = np.array(cam, dtype=np.int64) * 9; pt_off = 9 * n_cam + np.array(pt, dtype=np.int64) * 3
    obs = rng.normal(size=2 * cam_off.size)
    lay = Layout(hc, cam_off, pt_off, obs)
    assert lay.status == 0 and lay.n_giant == 4                            # 256 still fits one regular tile
    assert sorted(lay.gp_tile_count.tolist()) == [2, 2, 3, 3]             # 257 -> 2, 300 -> 2, 513 -> 3, 700 -> 3
    check_layout(lay, cam_off, pt_off, obs)
    for world in (2, 3):                                                  # long tracks are never split across ranks
        total = 0
        for r in range(world):
            sub_ = Layout(hc, cam_off, pt_off, obs, rank=r, world=world)
            assert sub_.status == 0
            total += sub_.n_giant
        assert total == 4


def test_ragged_tracks_fill_tiles(hc):
    """Tracks of very different lengths (1 .. 256 observations) still give whole-point tiles."""
    rng = np.random.default_rng(3)
    n_cam = 300
    lens = np.concatenate([[256, 1, 255, 2], rng.integers(1, 40, 200), [256]])
    cam, pt = [], []
    for j, k in enumerate(lens):
        cam.extend(sorted(rng.choice(n_cam, size=k, replace=False))); pt.extend([j] * k)
    cam_off = np.array(cam, dtype=np.int64) * 9; pt_off = 9 * n_cam + np.array(pt, dtype=np.int64) * 3
    obs = rng.normal(size=2 * cam_off.size)
    lay = Layout(hc, cam_off, pt_off, obs)
    assert lay.status == 0
    check_layout(lay, cam_off, pt_off, obs)


def _ragged_problem(seed, n_cam=300):
    rng = np.random.default_rng(seed)
    lens = np.concatenate([[256, 1, 255, 2, 17, 33], rng.integers(1, 40, 200), [256, 300], rng.integers(2, 9, 300)])
    cam, pt = [], []
    for j, k in enumerate(lens):
        cam.extend(sorted(rng.choice(n_cam, size=k, replace=False))); pt.extend([j] * k)
    return np.array(cam, dtype=np.int64) * 9, 9 * n_cam + np.array(pt, dtype=np.int64) * 3, rng.normal(size=2 * len(cam))


@pytest.mark.parametrize("case", ["ladybug-49", "small", "ragged"])
def test_two_level_sums_of_the_schur_product(hc, case):
    """The per-tile records of the implicit-Schur product (ba_tile_rec.h) and the chunked two-level sums that read them:
    the SAME per-item functions the kernels call (point_chunk_sum / point_combine / seg_chunk_sum / seg_combine), run item
    by item on the host over every regular tile, give the per-point and per-(tile, camera) sums of the layout's own
    lists, every position lies in exactly one chunk of its own point / segment, and srank inverts sperm."""
    if case == "ragged":
        cam_off, pt_off, obs = _ragged_problem(11)
    else:
        d = synth.make_bal(case, seed=2)
        off = d.block_offsets()
        cam_off, pt_off, obs = np.ascontiguousarray(off[:, 0]), np.ascontiguousarray(off[:, 1]), d.observations
    st = C.c_int(); err = C.create_string_buffer(256)
    h = hc.hc_layout_build(cam_off.size, p(cam_off), p(pt_off), p(obs), 0, 1, C.byref(st), err)
    assert h and st.value == 0, err.value
    out = np.zeros(4)
    hc.hc_check_two_level_sums.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    hc.hc_check_two_level_sums(C.c_void_p(h), 12345, p(out))
    hc.hc_layout_free(C.c_void_p(h))
    assert out[3] == 1.0, "chunk tables do not partition the tile"
    assert out[2] > 1000 and out[0] <= 4e-16 * 6 and out[1] <= 4e-16 * 10, out      # a handful of roundings apart, no more


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("sparse_offsets", [False, True])
def test_rank_local_ingestion_builds_the_same_layout(hc, world, sparse_offsets):
    """sk_solver_options.residual_blocks_are_local: a rank that is handed only the residual blocks of its own points, with
    every camera declared (Problem::AddParameterBlock), must end up with exactly the layout the replicated ingestion
    cuts out of the whole problem -- same camera table, same tiles, same segments -- so both modes give the same bits."""
    from skeres_b200 import api
    d = synth.make_bal("ladybug-49", seed=1)
    off = d.block_offsets()
    cam_off, pt_off = off[:, 0].copy(), off[:, 1].copy()
    all_cams = 9 * np.arange(d.num_cameras, dtype=np.int64)
    if sparse_offsets:                                   # the sort + binary-search id path of the builder
        cam_off = cam_off * 1000003; pt_off = pt_off * 1000003 + 5; all_cams = all_cams * 1000003
    ptr = np.concatenate([[0], np.cumsum(np.bincount(d.point_index, minlength=d.num_points))]).astype(np.int64)
    begin = api.partition_points(ptr, world)
    for r in range(world):
        o0, o1 = ptr[begin[r]], ptr[begin[r + 1]]
        # drop one camera's observations from this rank so that the declared list matters
        whole = Layout(hc, cam_off, pt_off, d.observations, rank=r, world=world)
        mine = Layout(hc, cam_off[o0:o1], pt_off[o0:o1], d.observations[2 * o0:2 * o1], extra_cams=all_cams)
        assert whole.status == 0 and mine.status == 0
        for name in Layout.FIELDS:
            a, b = getattr(whole, name), getattr(mine, name)
            if name == "perm":
                a = a - o0                               # the replicated builder indexes the whole input
            assert np.array_equal(a, b), name
    # a camera without any local observation still gets its id
    keep = cam_off != cam_off.min()
    part = Layout(hc, cam_off[keep], pt_off[keep], d.observations.reshape(-1, 2)[keep].ravel(), extra_cams=all_cams)
    assert part.status == 0 and part.n_cams == d.num_cameras and np.array_equal(part.cam_offset, all_cams)
    assert part.cam_seg_ptr[1] == 0                      # ... and owns no segment
    bare = Layout(hc, cam_off[keep], pt_off[keep], d.observations.reshape(-1, 2)[keep].ravel())
    assert bare.n_cams == d.num_cameras - 1


@pytest.mark.parametrize("world", [2, 3, 8])
def test_point_partition(hc, world):
    from skeres_b200 import api
    d = synth.make_bal("ladybug-49", seed=1)
    off = d.block_offsets()
    ptr = np.concatenate([[0], np.cumsum(np.bincount(d.point_index))]).astype(np.int64)
    begin = api.partition_points(ptr, world)
    assert begin[0] == 0 and begin[-1] == d.num_points and np.all(np.diff(begin) >= 0)
    per_rank = np.diff(ptr[begin])
    assert per_rank.sum() == d.num_observations
    assert per_rank.max() - per_rank.min() <= 2 * np.diff(ptr).max()          # balanced by observation count
    seen = []
    for r in range(world):                                                     # the per-rank layouts tile the problem exactly
        lay = Layout(hc, off[:, 0], off[:, 1], d.observations, rank=r, world=world)
        assert lay.status == 0 and lay.n_cams == d.num_cameras                 # cameras are replicated
        check_layout(lay, off[:, 0], off[:, 1], d.observations, pt_range=(begin[r], begin[r + 1]))
        seen.append(lay.perm)
    assert sorted(np.concatenate(seen).tolist()) == list(range(d.num_observations))


# --------------------------------------------------------------------------------------------------- device arithmetic on the host
def test_device_functors_match_the_oracle(hc, oracle):
    """jet.cuh (two-stage 6-wide duals for Snavely, full-width duals elsewhere) vs the 12-wide spire-style oracle."""
    d = synth.make_bal("small", seed=7)
    rng = np.random.default_rng(0)
    worst = 0.0
    for i in list(rng.integers(0, d.num_observations, 300)) + list(np.nonzero(d.camera_index == 0)[0][:20]):   # camera 0 = Taylor branch
        cam = d.parameters[9 * d.camera_index[i]:9 * d.camera_index[i] + 9]
        pt = d.parameters[9 * d.num_cameras + 3 * d.point_index[i]:][:3]
        obs = d.observations[2 * i:2 * i + 2]
        ok, res, (F, E) = oracle.evaluate(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, obs, [cam, pt])
        x = np.ascontiguousarray(np.concatenate([cam, pt])); r2 = np.zeros(2); jac = np.zeros(24)
        assert hc.hc_evaluate(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, p(np.ascontiguousarray(obs)), p(x), p(r2), p(jac)) == 1
        J = np.concatenate([F, E], axis=1)
        assert np.allclose(r2, res, rtol=1e-13, atol=1e-10)
        worst = max(worst, np.max(np.abs(jac.reshape(2, 12) - J) / np.abs(J).max()))
        r3 = np.zeros(2)
        hc.hc_snavely_residual(p(np.ascontiguousarray(cam)), p(np.ascontiguousarray(pt)), obs[0], obs[1], p(r3))
        assert np.allclose(r3, res, rtol=1e-13, atol=1e-10)
    assert worst < 1e-13
    import json
    for case in json.load(open(os.path.join(ROOT, "tests", "golden", "autodiff_spec_vectors.json")))["cases"]:
        x = np.ascontiguousarray(np.concatenate(case["parameters"]), dtype=np.float64)
        consts = np.ascontiguousarray(case["consts"] + [0.0], dtype=np.float64)
        nres = len(case["residuals"]); res = np.zeros(nres); jac = np.zeros(nres * x.size)
        assert hc.hc_evaluate(case["functor"], p(consts), p(x), p(res), p(jac)) == 1
        assert res.tolist() == case["residuals"]
        jac = jac.reshape(nres, x.size); col = 0
        for blk, want in zip(case["parameters"], case["jacobians"]):
            assert jac[:, col:col + len(blk)].ravel().tolist() == want
            col += len(blk)


def test_example_functors_device_code_matches_the_oracle(hc, oracle):
    """HelloWorld.scala / Powell.scala / PowellAnalytic.scala functors: the device Jet code (host build) equals the oracle's
    restatement bit for bit, at the examples' start points and at random points."""
    rng = np.random.default_rng(4)
    fids = [_abi.FUNCTOR_HELLO_WORLD, _abi.FUNCTOR_POWELL_F1, _abi.FUNCTOR_POWELL_F2, _abi.FUNCTOR_POWELL_ANALYTIC_F2, _abi.FUNCTOR_POWELL_F3,
            _abi.FUNCTOR_POWELL_F4]
    for fid in fids:
        nres, sizes, nconst = oracle.functor_info(fid)
        assert (nres, nconst) == (1, 0) and sizes == ([1] if fid == _abi.FUNCTOR_HELLO_WORLD else [1, 1])
        for trial in range(20):
            params = [[0.5]] if (fid == _abi.FUNCTOR_HELLO_WORLD and trial == 0) else [list(rng.normal(0, 3, n)) for n in sizes]
            ok, ro, jo = oracle.evaluate(fid, [], params)
            x = np.ascontiguousarray(np.concatenate(params)); res = np.zeros(3); jac = np.zeros(36); c = np.zeros(4)
            assert ok and hc.hc_evaluate(fid, p(c), p(x), p(res), p(jac)) == 1
            assert res[0] == ro[0] and np.array_equal(jac[:x.size], np.concatenate([np.ravel(j) for j in jo]))
            res2 = np.zeros(3)
            assert hc.hc_evaluate(fid, p(c), p(x), p(res2), None) == 1 and res2[0] == ro[0]          # residual-only branch


def test_device_loss_and_corrector_match_the_oracle(hc, oracle):
    rng = np.random.default_rng(1)
    for kind, a, b in [(_abi.LOSS_TRIVIAL, 0.0, 0.0), (_abi.LOSS_HUBER, 0.7, 0.0), (_abi.LOSS_CAUCHY, 0.5, 0.0), (_abi.LOSS_TOLERANT, 1.5, 0.4),
                       (_abi.LOSS_TOLERANT, 0.0, 2.0)]:
        for s in [0.0, 0.2, 3.0, 40.0]:
            rho = np.zeros(3)
            hc.hc_loss(kind, a, b, s, p(rho))
            assert np.allclose(rho, oracle.loss(kind, a, s, b), rtol=1e-15, atol=0)
    # Corrector through a one-block solve: compare the corrected residual/Jacobian implied by the oracle's gradient
    d = synth.make_bal("tiny", seed=3)
    for kind, a, b in [(_abi.LOSS_HUBER, 1.0, 0.0), (_abi.LOSS_CAUCHY, 2.0, 0.0), (_abi.LOSS_TOLERANT, 0.3, 0.2)]:
        op = oracle.OracleProblem(d.parameters)
        op.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets(), kind, a, b)
        cost, r, g, Jv = op.evaluate()
        op0 = oracle.OracleProblem(d.parameters)
        op0.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets())
        _, r0, _, J0 = op0.evaluate()
        for i in rng.integers(0, d.num_observations, 25):
            res = r0[2 * i:2 * i + 2].copy(); J = J0[24 * i:24 * i + 18].copy()
            hc.hc_correct(kind, a, b, 2, 9, p(res), p(J))
            assert np.allclose(res, r[2 * i:2 * i + 2], rtol=1e-14, atol=1e-14)
            assert np.allclose(J, Jv[24 * i:24 * i + 18], rtol=1e-13, atol=1e-13)


def test_tolerant_loss_and_the_alpha_branch_of_the_corrector(hc, oracle):
    """TolerantLoss (ceres.i:175) is the one registered loss with rho'' > 0, i.e. the one that takes the Corrector through
    its alpha != 0 branch (DESIGN.md parity gap 3, closed by this test).  Independent checks, not oracle-vs-device only:
    rho against the closed form and its numerical derivatives; the corrected Jacobian and residual against the identities
    they are built to satisfy (Triggs):  J~' J~ = J' (rho' I + 2 rho'' r r') J   and   J~' r~ = rho' J' r."""
    a, b = 0.8, 0.35
    c = b * np.log1p(np.exp(-a / b))
    f = lambda s: b * np.log1p(np.exp((s - a) / b)) - c
    for s in [0.0, 0.05, 0.6, 0.8, 1.3, 5.0, 13.0]:
        rho = np.zeros(3)
        hc.hc_loss(_abi.LOSS_TOLERANT, a, b, s, p(rho))
        assert np.allclose(rho, oracle.loss(_abi.LOSS_TOLERANT, a, s, b), rtol=1e-15, atol=0)
        h = 1e-5
        assert np.isclose(rho[0], f(s), rtol=1e-13, atol=1e-15)
        assert np.isclose(rho[1], (f(s + h) - f(s - h)) / (2 * h), rtol=1e-7, atol=1e-10)
        h2 = 1e-3                                                                  # second difference: round-off ~ eps f / h^2
        assert np.isclose(rho[2], (f(s + h2) - 2 * f(s) + f(s - h2)) / h2 ** 2, rtol=1e-4, atol=1e-8) and (rho[2] > 0 or s > a + 36.7 * b)
    big = np.zeros(3)
    hc.hc_loss(_abi.LOSS_TOLERANT, a, b, a + 40.0 * b, p(big))                     # the overflow guard: rho -> s - a - c, rho' = 1, rho'' = 0
    assert np.allclose(big, [40.0 * b - c, 1.0, 0.0], rtol=1e-15)
    assert f(0.0) == 0.0 or abs(f(0.0)) < 1e-16                                    # rho(0) = 0
    rng = np.random.default_rng(9)
    took_alpha = 0
    for _ in range(40):
        r = rng.normal(0, 0.7, 2); J = rng.normal(size=(2, 9))
        sq = float(r @ r)
        rho = oracle.loss(_abi.LOSS_TOLERANT, a, sq, b)
        took_alpha += rho[2] > 0
        res = r.copy(); Jc = np.ascontiguousarray(J.copy())
        hc.hc_correct(_abi.LOSS_TOLERANT, a, b, 2, 9, p(res), p(Jc))
        assert np.allclose(Jc.T @ Jc, J.T @ (rho[1] * np.eye(2) + 2.0 * rho[2] * np.outer(r, r)) @ J, rtol=1e-11, atol=1e-12)
        assert np.allclose(Jc.T @ res, rho[1] * (J.T @ r), rtol=1e-11, atol=1e-12)
    assert took_alpha == 40
    # ... and the oracle's own Corrector (through a problem evaluation) satisfies the same identities
    d = synth.make_bal("tiny", seed=3)
    op = oracle.OracleProblem(d.parameters)
    op.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets(), _abi.LOSS_TOLERANT, 0.3, 0.2)
    cost, rc, g, Jv = op.evaluate()
    op0 = oracle.OracleProblem(d.parameters)
    op0.add_residual_blocks(_abi.FUNCTOR_SNAVELY_REPROJECTION_ERROR, d.observations.reshape(-1, 2), d.block_offsets())
    cost0, r0, g0, J0 = op0.evaluate()
    want_cost = 0.0
    for i in range(d.num_observations):
        blocks = lambda v: np.hstack([v[24 * i:24 * i + 18].reshape(2, 9), v[24 * i + 18:24 * i + 24].reshape(2, 3)])   # [F | E], each row-major
        r = r0[2 * i:2 * i + 2]; J = blocks(J0)
        rho = oracle.loss(_abi.LOSS_TOLERANT, 0.3, float(r @ r), 0.2)
        want_cost += 0.5 * rho[0]
        Jt = blocks(Jv); rt = rc[2 * i:2 * i + 2]
        assert np.allclose(Jt.T @ Jt, J.T @ (rho[1] * np.eye(2) + 2.0 * rho[2] * np.outer(r, r)) @ J, rtol=1e-10, atol=1e-9)
        assert np.allclose(Jt.T @ rt, rho[1] * (J.T @ r), rtol=1e-10, atol=1e-9)
    assert np.isclose(cost, want_cost, rtol=1e-13)


def test_small_spd_inverses(hc):
    rng = np.random.default_rng(2)
    for _ in range(50):
        A = rng.normal(size=(5, 3)); M = A.T @ A + 1e-3 * np.eye(3)
        m6 = np.array([M[0, 0], M[0, 1], M[0, 2], M[1, 1], M[1, 2], M[2, 2]]); inv6 = np.zeros(6)
        assert hc.hc_invert_spd3(p(m6), p(inv6)) == 1
        W = np.linalg.inv(M)
        assert np.allclose(inv6, [W[0, 0], W[0, 1], W[0, 2], W[1, 1], W[1, 2], W[2, 2]], rtol=1e-10)
        B = rng.normal(size=(20, 9)); S = np.ascontiguousarray(B.T @ B + 1e-3 * np.eye(9))
        want = np.linalg.inv(S)
        assert hc.hc_invert_spd9(p(S), 9) == 1 and np.allclose(S, want, rtol=1e-9, atol=1e-12)
    bad = np.array([1.0, 2.0, 0.0, 1.0, 0.0, 1.0]); out = np.zeros(6)
    assert hc.hc_invert_spd3(p(bad), p(out)) == 0            # not positive definite -> reported, not NaN-propagated


def test_bal_block_offsets_helper():
    """sk_bal_block_offsets == BalProblem.mutableCameraForObservation / mutablePointForObservation for every observation
    (SimpleBundleAdjuster.scala:28-33); host-only, index errors are reported with the lowest offending observation."""
    from skeres_b200 import api
    d = synth.make_bal("ladybug-49", seed=3)
    bal = api.BalProblem(d.num_cameras, d.num_points, d.camera_index, d.point_index, d.observations, None)
    off = bal.blockOffsets()
    assert np.array_equal(off, d.block_offsets())
    assert np.array_equal(off[:, 0], 9 * d.camera_index.astype(np.int64))
    assert np.array_equal(off[:, 1], 9 * d.num_cameras + 3 * d.point_index.astype(np.int64))
    assert np.array_equal(bal.blockOffsets(100, 250), off[100:250])
    assert bal.blockOffsets(5, 5).shape == (0, 2)
    bal.cameraIndex = bal.cameraIndex.copy()
    bal.cameraIndex[[11, 300]] = [-1, d.num_cameras]
    with pytest.raises(api.SkeresError) as e:
        bal.blockOffsets()
    assert e.value.status == _abi.ERR_INVALID_ARGUMENT and "observation 11" in str(e.value)


def test_graft_entry_build_runs():
    """__graft_entry__.build() is the driver's "does it build" check: incremental make + import + ABI version."""
    sys_path = list(__import__("sys").path)
    __import__("sys").path.insert(0, ROOT)
    try:
        import __graft_entry__ as g
        g.build()
    finally:
        __import__("sys").path[:] = sys_path


def test_jni_shim_type_checks():
    """bindings/jni/skeres_jni.c (the reference-side binding) cannot run here -- no JDK -- but its calls into
    include/skeres.h are type-checked against a mock jni.h; every native method the Scala facade declares has its C export."""
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I" + os.path.join(ROOT, "tests", "mock_jni"),
                        "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "bindings", "jni", "skeres_jni.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    shim = open(os.path.join(ROOT, "bindings", "jni", "skeres_jni.c")).read()
    natives = re.findall(r"@native def (\w+)\(", open(os.path.join(ROOT, "bindings", "scala", "Native.scala")).read())
    exported = set(re.findall(r"FN\((\w+)\)\(JNIEnv", shim))
    assert len(natives) > 40 and set(natives) == exported, set(natives) ^ exported
    used = set(re.findall(r"\b(sk_[a-z0-9_]+)\s*\(", shim))
    assert used <= set(declared_functions())            # the shim calls only what the header declares


@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
def test_local_range_equals_partition_points(world):
    """BalProblem.localRange (boundary scan, O(track length)) gives exactly the ranges of sk_partition_points over the point
    CSR -- including tracks that straddle a target and ranks that end up empty."""
    from skeres_b200 import api
    rng = np.random.default_rng(world)
    cases = [synth.make_bal("ladybug-49", seed=2), synth.make_bal(n_cam=40, n_pt=7, n_obs=120, seed=1, long_tracks=(40, 35))]
    for d in cases:
        bal = api.BalProblem(d.num_cameras, d.num_points, d.camera_index, d.point_index, d.observations, None)
        ptr = np.concatenate([[0], np.cumsum(np.bincount(d.point_index, minlength=d.num_points))]).astype(np.int64)
        begin = api.partition_points(ptr, world)
        for r in range(world):
            assert bal.localRange(r, world) == (int(ptr[begin[r]]), int(ptr[begin[r + 1]])), (r, world)
        assert bal.localRange(0, world)[0] == 0 and bal.localRange(world - 1, world)[1] == d.num_observations


def test_user_functor_sources_compile_without_a_gpu(tmp_path, monkeypatch):
    """sk_functor_register_source: NVRTC turns the functor source plus the library's own jet.cuh into
    sm_100a kernels; compiling needs no device (the module is loaded at first use), so the generated translation unit is
    checked here.  A source that does not compile comes back as INVALID_ARGUMENT with the compiler's log."""
    import ctypes as C
    import shutil
    import user_functor_sources as U
    from skeres_b200 import _abi
    from skeres_b200._lib import lib
    monkeypatch.setenv("SKERES_DUMP_USER_CUBIN", str(tmp_path))        # development aid of user_functor.cu: <dir>/<name>.cubin
    ids = []
    # the last two have the bundle-adjustment shape (2; 9, 3; 2 constants): their translation unit also holds the TILE evaluation
    # kernels of the Schur solvers (ba_evaluate.cuh, handed to NVRTC with the library's own ba_dev.cuh / ba_tile.cuh)
    for name, src, nres, sizes, nconsts in U.SPEC + [("UserExponentialResidual", U.EXPONENTIAL, 1, [1, 1], 2)] + U.BA_SHAPED:
        fid = C.c_int(0)
        st = lib.sk_functor_register_source(name.encode(), src.encode(), nres, len(sizes), (C.c_int * len(sizes))(*sizes), nconsts, C.byref(fid))
        assert st == _abi.OK, lib.sk_last_error().decode()
        assert fid.value >= 1000 and fid.value not in ids
        ids.append(fid.value)
        nr, nb, nc = C.c_int(), C.c_int(), C.c_int()
        bs = (C.c_int * 10)()
        assert lib.sk_functor_info(fid.value, C.byref(nr), C.byref(nb), bs, C.byref(nc)) == _abi.OK
        assert (nr.value, nb.value, nc.value, list(bs)[:len(sizes)]) == (nres, len(sizes), nconsts, sizes)
    # which kernels a translation unit holds: the two generic ones always, the tile kernels of the Schur solvers for the BA shape only
    if shutil.which("cuobjdump"):
        def kernels(name):
            out = subprocess.run(["cuobjdump", "--dump-resource-usage", str(tmp_path / f"{name}.cubin")], capture_output=True, text=True).stdout
            return {k for k in ("sk_user_evaluate_single", "sk_user_dense_evaluate", "sk_user_ba_evaluate_jac", "sk_user_ba_evaluate_cost") if k in out}
        assert kernels("BinaryScalarCost") == {"sk_user_evaluate_single", "sk_user_dense_evaluate"}
        for name, *_ in U.BA_SHAPED:
            assert kernels(name) == {"sk_user_evaluate_single", "sk_user_dense_evaluate", "sk_user_ba_evaluate_jac", "sk_user_ba_evaluate_cost"}
    fid = C.c_int(0)
    bad = "template <class T> __device__ bool Broken(const double* c, T const* const* x, T* r) { r[0] = undefined_symbol; return true; }"
    st = lib.sk_functor_register_source(b"Broken", bad.encode(), 1, 1, (C.c_int * 1)(1), 0, C.byref(fid))
    assert st == _abi.ERR_INVALID_ARGUMENT and "undefined_symbol" in lib.sk_last_error().decode()
    st = lib.sk_functor_register_source(b"not an identifier", bad.encode(), 1, 1, (C.c_int * 1)(1), 0, C.byref(fid))
    assert st == _abi.ERR_INVALID_ARGUMENT
