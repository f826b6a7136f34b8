// ba_layout.cu — host-side construction of the tiled BA layout (see ba_layout.h).
#include "ba_layout.h"
#include "ba_tile_rec.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <numeric>
#include <thread>

#include "common.cuh"
#include "host_parallel.h"

namespace sk {

// development: SKERES_TRACE_HOST prints where the builder's time goes
struct Lap {
  bool on = std::getenv("SKERES_TRACE_HOST") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void operator()(const char* what) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[skeres] layout: %-28s %.3f s\n", what, std::chrono::duration<double>(now - t).count());
    t = now;
  }
};

void partition_points(int64_t n_points, const int64_t* point_ptr, int world_size, int64_t* out_begin) {
  const int64_t total = point_ptr[n_points];
  out_begin[0] = 0;
  for (int r = 1; r < world_size; ++r) {
    // first point whose observation prefix reaches r/world of the total
    const int64_t target = (total * r) / world_size;
    const int64_t* it = std::lower_bound(point_ptr, point_ptr + n_points + 1, target);
    int64_t p = it - point_ptr;
    if (p > n_points) p = n_points;
    if (p < out_begin[r - 1]) p = out_begin[r - 1];
    out_begin[r] = p;
  }
  out_begin[world_size] = n_points;
}

// Maps arbitrary block offsets to dense ids ordered by offset.
static void dense_ids(int64_t n, const int64_t* off_base, int64_t stride, int block_size, std::vector<int64_t>* uniq,
                      std::vector<int32_t>* ids, const std::vector<int64_t>* extra = nullptr) {
  struct Strided { const int64_t* p; int64_t s; int64_t operator[](int64_t i) const { return p[i * s]; } } off{off_base, stride};
  int64_t lo = off[0], hi = off[0];
  {
    std::mutex mu;
    parallel_for(n, [&](int64_t a, int64_t b) {
      int64_t l = off[a], h = off[a];
      for (int64_t i = a + 1; i < b; ++i) { l = std::min(l, off[i]); h = std::max(h, off[i]); }
      std::lock_guard<std::mutex> g(mu);
      lo = std::min(lo, l); hi = std::max(hi, h);
    });
  }
  if (extra != nullptr) for (int64_t o : *extra) { lo = std::min(lo, o); hi = std::max(hi, o); }
  const int64_t range = hi - lo + 1;
  ids->resize(n);
  if (range <= std::max<int64_t>(64 * n, 1 << 20)) {           // direct table
    std::vector<int32_t> table((size_t)range, -1);
    // every thread stores the same value: relaxed atomic stores keep that defined
    parallel_for(n, [&](int64_t a, int64_t b) { for (int64_t i = a; i < b; ++i) __atomic_store_n(&table[off[i] - lo], 0, __ATOMIC_RELAXED); });
    if (extra != nullptr) for (int64_t o : *extra) table[o - lo] = 0;
    int32_t next = 0;
    int64_t last = -(int64_t)block_size;
    uniq->clear();
    for (int64_t k = 0; k < range; ++k) if (table[k] == 0) {
      SK_REQUIRE(k - last >= block_size, SK_ERR_INVALID_ARGUMENT, "overlapping parameter blocks at offsets %lld and %lld",
                 (long long)(last + lo), (long long)(k + lo));
      table[k] = next++; uniq->push_back(k + lo); last = k;
    }
    parallel_for(n, [&](int64_t a, int64_t b) { for (int64_t i = a; i < b; ++i) (*ids)[i] = table[off[i] - lo]; });
  } else {                                                       // sort + binary search
    std::vector<int64_t> u((size_t)n);
    for (int64_t i = 0; i < n; ++i) u[i] = off[i];
    if (extra != nullptr) u.insert(u.end(), extra->begin(), extra->end());
    std::sort(u.begin(), u.end());
    u.erase(std::unique(u.begin(), u.end()), u.end());
    for (size_t k = 1; k < u.size(); ++k)
      SK_REQUIRE(u[k] - u[k - 1] >= block_size, SK_ERR_INVALID_ARGUMENT, "overlapping parameter blocks");
    for (int64_t i = 0; i < n; ++i) (*ids)[i] = (int32_t)(std::lower_bound(u.begin(), u.end(), off[i]) - u.begin());
    uniq->swap(u);
  }
}

void build_ba_layout(int64_t n, const int64_t* cam_off, const int64_t* pt_off, const double* obs_xy,
                     int rank, int world_size, BaLayoutHost* out, int64_t offset_stride, const std::vector<int64_t>* extra_cam_off) {
  BaLayoutHost& L = *out;
  SK_REQUIRE(n > 0, SK_ERR_INVALID_ARGUMENT, "bundle adjustment problem without observations");
  SK_REQUIRE(n < (int64_t)2000000000, SK_ERR_UNSUPPORTED, "more than 2e9 observations");
  std::vector<int32_t> cam_id, pt_id;
  std::vector<int64_t> pt_offsets_all;
  Lap lap;
  dense_ids(n, cam_off, offset_stride, 9, &L.cam_offset, &cam_id, extra_cam_off);
  dense_ids(n, pt_off, offset_stride, 3, &pt_offsets_all, &pt_id);
  lap("dense ids");
  L.n_cams = (int32_t)L.cam_offset.size();
  const int64_t n_pts_all = (int64_t)pt_offsets_all.size();
  {  // camera and point blocks must not overlap each other: one merge pass over the two sorted offset lists
    size_t ic = 0, ip = 0;
    int64_t prev_end = INT64_MIN;
    while (ic < L.cam_offset.size() || ip < pt_offsets_all.size()) {
      const bool take_cam = ip >= pt_offsets_all.size() || (ic < L.cam_offset.size() && L.cam_offset[ic] <= pt_offsets_all[ip]);
      const int64_t o = take_cam ? L.cam_offset[ic++] : pt_offsets_all[ip++];
      SK_REQUIRE(o >= prev_end, SK_ERR_INVALID_ARGUMENT, "camera and point parameter blocks overlap at offset %lld", (long long)o);
      prev_end = o + (take_cam ? 9 : 3);
    }
  }
  lap("overlap check");
  // ---- sort by (point, camera) -------------------------------------------------------------
  bool sorted = true;
  {
    std::mutex mu;
    parallel_for(n, [&](int64_t a, int64_t b) {
      bool ok = true;
      for (int64_t i = std::max<int64_t>(a, 1); i < b && ok; ++i)
        ok = (pt_id[i] > pt_id[i - 1]) || (pt_id[i] == pt_id[i - 1] && cam_id[i] >= cam_id[i - 1]);
      if (!ok) { std::lock_guard<std::mutex> g(mu); sorted = false; }
    });
  }
  L.input_was_sorted = sorted;
  std::vector<int32_t> order;                                   // sorted position -> input index; empty = identity
  if (!sorted) {
    // Counting sort by point (ids are dense), then every point's few observations by (camera, input index): all host
    // threads, ~30x faster than one stable_sort over 5M keys, and the same order (ties keep their input order).
    order.resize((size_t)n);
    std::vector<int32_t> start((size_t)n_pts_all + 1, 0);
    parallel_for(n, [&](int64_t a, int64_t b) { for (int64_t i = a; i < b; ++i) __atomic_fetch_add(&start[(size_t)pt_id[i] + 1], 1, __ATOMIC_RELAXED); });
    for (int64_t p = 0; p < n_pts_all; ++p) start[p + 1] += start[p];
    std::vector<int32_t> next(start.begin(), start.end() - 1);
    parallel_for(n, [&](int64_t a, int64_t b) {
      for (int64_t i = a; i < b; ++i) order[(size_t)__atomic_fetch_add(&next[pt_id[i]], 1, __ATOMIC_RELAXED)] = (int32_t)i;
    });
    parallel_for(n_pts_all, [&](int64_t p0, int64_t p1) {
      for (int64_t p = p0; p < p1; ++p)
        std::sort(order.begin() + start[p], order.begin() + start[p + 1],
                  [&](int32_t a, int32_t b) { return cam_id[a] != cam_id[b] ? cam_id[a] < cam_id[b] : a < b; });
    });
  }
  auto src = [&](int64_t j) -> int64_t { return sorted ? j : (int64_t)order[j]; };
  lap("sort / sortedness");
  // global point CSR over the sorted list (every dense point id occurs, so the run boundaries ARE the CSR), then
  // this rank's point range
  std::vector<int64_t> gptr((size_t)n_pts_all + 1);
  gptr[0] = 0; gptr[n_pts_all] = n;
  parallel_for(n, [&](int64_t a, int64_t b) {
    for (int64_t j = std::max<int64_t>(a, 1); j < b; ++j) {
      const int32_t pj = pt_id[src(j)];
      if (pj != pt_id[src(j - 1)]) gptr[pj] = j;
    }
  });
  std::vector<int64_t> begin((size_t)world_size + 1);
  partition_points(n_pts_all, gptr.data(), world_size, begin.data());
  const int64_t p0 = begin[rank], p1 = begin[rank + 1];
  const int64_t o0 = gptr[p0], o1 = gptr[p1];
  L.n_pts = (int32_t)(p1 - p0);
  L.n_obs = (int32_t)(o1 - o0);
  L.pt_offset.assign(pt_offsets_all.begin() + p0, pt_offsets_all.begin() + p1);
  if (world_size > 1) L.all_pt_offset = pt_offsets_all;   // the caller needs them to publish the full solution
  L.perm.resize(L.n_obs);
  if (sorted && world_size == 1) {                        // the common case (BAL files): ids and observations are already in place
    L.obs_cam = std::move(cam_id); L.obs_pt = std::move(pt_id);
    L.obs.clear(); L.obs_src = obs_xy;
    parallel_for(L.n_obs, [&](int64_t j0, int64_t j1) { for (int64_t j = j0; j < j1; ++j) L.perm[j] = (int32_t)j; });
  } else {
    L.obs.resize((size_t)2 * L.n_obs); L.obs_src = L.obs.data();
    L.obs_cam.resize(L.n_obs); L.obs_pt.resize(L.n_obs);
    parallel_for(L.n_obs, [&](int64_t j0, int64_t j1) {
      for (int64_t j = j0; j < j1; ++j) {
        const int64_t i = src(o0 + j);
        L.perm[j] = (int32_t)i;
        L.obs[2 * j] = obs_xy[2 * i]; L.obs[2 * j + 1] = obs_xy[2 * i + 1];
        L.obs_cam[j] = cam_id[i];
        L.obs_pt[j] = (int32_t)(pt_id[i] - p0);
      }
    });
  }
  lap("point CSR + gather");
  L.pt_ptr.resize((size_t)L.n_pts + 1);
  for (int64_t p = 0; p <= L.n_pts; ++p) L.pt_ptr[p] = (int32_t)(gptr[p0 + p] - o0);
  {
    int64_t first_dup = -1;                                  // lowest index, so the message does not depend on the thread count
    std::mutex mu;
    parallel_for(L.n_obs, [&](int64_t a, int64_t b) {
      for (int64_t j = std::max<int64_t>(a, 1); j < b; ++j)
        if (L.obs_pt[j] == L.obs_pt[j - 1] && L.obs_cam[j] == L.obs_cam[j - 1]) {
          std::lock_guard<std::mutex> g(mu);
          if (first_dup < 0 || j < first_dup) first_dup = j;
          break;
        }
    });
    SK_REQUIRE(first_dup < 0, SK_ERR_UNSUPPORTED,
               "a camera observes the same point twice (residual blocks %d and %d): unsupported by the Schur path",
               L.perm[std::max<int64_t>(first_dup, 1) - 1], L.perm[std::max<int64_t>(first_dup, 0)]);
  }
  lap("pt_ptr + duplicate check");
  // ---- tiles: whole points, at most kTileObs observations -----------------------------------
  // tile t covers observations [tile_obs[t], tile_obs[t+1]) and points [tile_pt[t], tile_pt[t] + tile_np[t])
  L.tile_obs.clear(); L.tile_pt.clear(); L.tile_np.clear(); L.tile_chunk.clear();
  L.gp_tile_begin.clear(); L.gp_tile_count.clear(); L.gp_point.clear();
  int32_t cur = 0, cur_first_pt = 0, cur_first_obs = 0;
  L.n_chunks = 0;
  auto close_regular = [&](int32_t p_end) {          // emit the pending regular tile [cur_first_pt, p_end)
    if (cur == 0) return;
    L.tile_obs.push_back(cur_first_obs); L.tile_pt.push_back(cur_first_pt); L.tile_np.push_back(p_end - cur_first_pt); L.tile_chunk.push_back(-1);
    cur = 0;
  };
  for (int32_t p = 0; p < L.n_pts; ++p) {
    const int32_t k = L.pt_ptr[p + 1] - L.pt_ptr[p];
    if (k > kTileObs) {                              // long track: its own chain of chunk tiles
      close_regular(p);
      const int32_t nch = (k + kTileObs - 1) / kTileObs;
      L.gp_point.push_back(p); L.gp_tile_begin.push_back((int32_t)L.tile_obs.size()); L.gp_tile_count.push_back(nch);
      for (int32_t c = 0; c < nch; ++c) {
        L.tile_obs.push_back(L.pt_ptr[p] + (int32_t)((int64_t)k * c / nch));
        L.tile_pt.push_back(p); L.tile_np.push_back(1); L.tile_chunk.push_back(L.n_chunks++);
      }
      cur_first_pt = p + 1; cur_first_obs = L.pt_ptr[p + 1];
      continue;
    }
    if (cur + k > kTileObs) close_regular(p);
    if (cur == 0) { cur_first_pt = p; cur_first_obs = L.pt_ptr[p]; }
    cur += k;
  }
  close_regular(L.n_pts);
  L.n_tiles = (int32_t)L.tile_obs.size();
  L.tile_obs.push_back(L.n_obs); L.tile_pt.push_back(L.n_pts);
  L.n_giant = (int32_t)L.gp_point.size();
  lap("tiles");
  // ---- tile-local camera segments (tiles are independent: built by all host threads) -------------
  SK_REQUIRE(L.n_cams < (1 << 24), SK_ERR_UNSUPPORTED, "more than 2^24 cameras");
  L.obs_slot.resize(L.n_obs); L.obs_ptl.resize(L.n_obs); L.seg_perm.resize(L.n_obs);
  L.tile_seg.assign((size_t)L.n_tiles + 1, 0);
  // pass 1: camera-sort each tile, remember the tile-local segment structure, count segments
  std::vector<int32_t> tile_nseg((size_t)L.n_tiles, 0);
  parallel_for(L.n_tiles, [&](int64_t t0, int64_t t1) {
    uint32_t keys[kTileObs];
    for (int64_t t = t0; t < t1; ++t) {
      const int32_t ob = L.tile_obs[t], oe = L.tile_obs[t + 1], pb = L.tile_pt[t], no = oe - ob;
      for (int32_t j = 0; j < no; ++j) {
        L.obs_ptl[ob + j] = (uint16_t)(L.obs_pt[ob + j] - pb);
        keys[j] = ((uint32_t)L.obs_cam[ob + j] << 8) | (uint32_t)j;          // kTileObs <= 256, cams < 2^24
      }
      std::sort(keys, keys + no);
      int32_t prev_cam = -1, nseg = 0;
      for (int32_t q = 0; q < no; ++q) {
        const int32_t cam = (int32_t)(keys[q] >> 8), loc = (int32_t)(keys[q] & 255u);
        if (cam != prev_cam) { ++nseg; prev_cam = cam; }
        L.seg_perm[ob + q] = (uint16_t)loc;
        L.obs_slot[ob + loc] = (uint16_t)(nseg - 1);
      }
      tile_nseg[t] = nseg;
    }
  });
  L.max_seg_tile = 0; L.max_pt_tile = 0;
  for (int32_t t = 0; t < L.n_tiles; ++t) {
    L.tile_seg[t + 1] = L.tile_seg[t] + tile_nseg[t];
    L.max_seg_tile = std::max(L.max_seg_tile, tile_nseg[t]);
    L.max_pt_tile = std::max(L.max_pt_tile, L.tile_np[t]);
  }
  L.n_segs = L.tile_seg[L.n_tiles];
  L.seg_ptr.assign((size_t)L.n_segs + 1, 0); L.seg_cam.assign((size_t)L.n_segs, 0);
  // pass 2: segment starts and cameras
  parallel_for(L.n_tiles, [&](int64_t t0, int64_t t1) {
    for (int64_t t = t0; t < t1; ++t) {
      const int32_t ob = L.tile_obs[t], oe = L.tile_obs[t + 1];
      int32_t s = L.tile_seg[t] - 1, prev_cam = -1;
      for (int32_t q = ob; q < oe; ++q) {
        const int32_t cam = L.obs_cam[ob + L.seg_perm[q]];
        if (cam != prev_cam) { ++s; L.seg_ptr[s] = q; L.seg_cam[s] = cam; prev_cam = cam; }
      }
    }
  });
  L.seg_ptr[L.n_segs] = L.n_obs;
  lap("segments");
  // ---- camera -> segments (tile order) --------------------------------------------------------
  L.cam_seg_ptr.assign((size_t)L.n_cams + 1, 0);
  for (int32_t s = 0; s < L.n_segs; ++s) L.cam_seg_ptr[L.seg_cam[s] + 1]++;
  for (int32_t c = 0; c < L.n_cams; ++c) L.cam_seg_ptr[c + 1] += L.cam_seg_ptr[c];
  L.cam_seg.resize(L.n_segs);
  std::vector<int32_t> fill(L.cam_seg_ptr.begin(), L.cam_seg_ptr.end() - 1);
  for (int32_t s = 0; s < L.n_segs; ++s) L.cam_seg[fill[L.seg_cam[s]]++] = s;
  lap("camera -> segments");
}

// Per-tile metadata records of the implicit-Schur product (ba_tile_rec.h): everything a tile needs besides the Jacobian and
// (E^T E)^-1, packed so that one bulk copy brings it on chip, plus the chunk tables of the two-level sums.
void build_tile_records(const BaLayoutHost& H, TileRecDims* dims, std::vector<unsigned char>* out) {
  const int T = kTileObs;
  const TileRecDims D = tile_rec_dims(std::max(H.max_seg_tile, 1), std::max(H.max_pt_tile, 1));
  const int sp = D.sp, pp = D.pp, sc = D.sc;
  const size_t stride = (size_t)D.stride;
  out->assign((size_t)std::max(H.n_tiles, 1) * stride, 0);
  unsigned char* rec = out->data();
  std::vector<int32_t> seg_pos((size_t)H.n_segs);            // inverse of cam_seg: position of a segment in the camera-major list
  for (int32_t t = 0; t < H.n_segs; ++t) seg_pos[H.cam_seg[t]] = t;
  parallel_for(H.n_tiles, [&](int64_t t_begin, int64_t t_end) {
  for (int64_t t = t_begin; t < t_end; ++t) {
    unsigned char* base = rec + (size_t)t * stride;
    uint16_t* slot = reinterpret_cast<uint16_t*>(base); uint16_t* ptl = slot + T; uint16_t* sperm = ptl + T; uint16_t* srank = sperm + T;
    int32_t* sptr = reinterpret_cast<int32_t*>(base + 8 * T); int32_t* pptr = sptr + sp; int32_t* scam = pptr + pp; int32_t* spos = scam + sp;
    uint16_t* pchunk = reinterpret_cast<uint16_t*>(spos + sp); uint16_t* pcptr = pchunk + T; uint16_t* schunk = pcptr + pp; uint16_t* scptr = schunk + sc;
    const int ob = H.tile_obs[t], no = H.tile_obs[t + 1] - ob, pb = H.tile_pt[t], np = H.tile_np[t];
    const int sb = H.tile_seg[t], ns = H.tile_seg[t + 1] - sb;
    for (int j = 0; j < no; ++j) {
      slot[j] = H.obs_slot[ob + j]; ptl[j] = H.obs_ptl[ob + j]; sperm[j] = H.seg_perm[ob + j];
      srank[H.seg_perm[ob + j]] = (uint16_t)j;
    }
    for (int s = 0; s <= ns; ++s) sptr[s] = H.seg_ptr[sb + s] - ob;
    for (int s = 0; s < ns; ++s) { scam[s] = H.seg_cam[sb + s]; spos[s] = seg_pos[sb + s]; }
    int nc = 0;
    for (int s = 0; s < ns; ++s) {                           // segment s = positions [sptr[s], sptr[s + 1]) of the camera-sorted order
      scptr[s] = (uint16_t)nc;
      for (int b = sptr[s]; b < sptr[s + 1]; b += kSegChunk) schunk[nc++] = (uint16_t)(b | ((std::min(kSegChunk, sptr[s + 1] - b) - 1) << 8));
    }
    scptr[ns] = (uint16_t)nc;
    if (H.tile_chunk[t] < 0) {                               // chunk tiles of long tracks have no per-point phase in the tile kernels
      for (int q = 0; q <= np; ++q) pptr[q] = H.pt_ptr[pb + q] - ob;
      nc = 0;
      for (int q = 0; q < np; ++q) {
        pcptr[q] = (uint16_t)nc;
        for (int b = pptr[q]; b < pptr[q + 1]; b += kPtChunk) pchunk[nc++] = (uint16_t)(b | ((std::min(kPtChunk, pptr[q + 1] - b) - 1) << 8));
      }
      pcptr[np] = (uint16_t)nc;
    }
  }
  });
  *dims = D;
}

}  // namespace sk
