// ba_layout_device.cu — the bundle-adjustment layout (ba_layout.h) built ON THE DEVICE from the uploaded residual-block table
// ("BAL reader -> SoA loader with on-device CSR").
//
// The host builder (ba_layout.cu) needs 0.06-0.075 s for the 5.0 M observations of the Venice shape on 16 host threads -- and
// 0.2 s when eight ranks share those threads -- which was the largest part of what the end-to-end path spends outside the LM
// iterations.  Here the (camera offset, point offset) pairs and the observations are uploaded as they are and everything that
// is per observation or per segment happens in kernels:
//   offsets -> dense ids      min / max, mark tables, exclusive scan (= id of every block, ordered by offset), look-up
//   checks                    sorted by (point, camera)?  a camera observing a point twice?  blocks overlapping?  a track longer
//                             than a tile?  -- any "no" hands the problem to the HOST builder, which sorts, chunks long tracks
//                             and produces the error messages; the device path covers the case BAL files and the reference's
//                             SimpleBundleAdjuster produce (SimpleBundleAdjuster.scala:52-58, :139-145): observations grouped
//                             by point, cameras ascending
//   point CSR                 run boundaries of the point ids
//   tiles                     greedy packing of whole points into tiles of <= 256 observations: one sequential pass over the
//                             point CSR, done on the host (2 ms for 1 M points; the CSR comes back in one 4 MB copy)
//   tile-local segments       one CTA per tile: bitonic sort of (camera << 8 | local index) in shared memory -> permutation,
//                             slots, segment count; scan of the counts; second pass writes segment starts and cameras
//   camera -> segments        stable radix sort of the segments by camera (tile order within a camera), CSR from run boundaries
// The result is, array for array, what the host builder produces (tests: SKERES_LAYOUT=host against the default, bitwise
// identical solves).  cub (part of the CUDA toolkit) supplies the scans and the radix sort: ingestion plumbing, not the hot path.
#include "ba_layout_device.cuh"

#include <algorithm>
#include <climits>
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "host_parallel.h"

namespace sk {

namespace {

constexpr int T = kTileObs;

__global__ void k_minmax(int64_t n, const long long* __restrict__ off2, long long* out) {
  long long lc = LLONG_MAX, hc = LLONG_MIN, lp = LLONG_MAX, hp = LLONG_MIN;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const long long c = off2[2 * i], p = off2[2 * i + 1];
    lc = min(lc, c); hc = max(hc, c); lp = min(lp, p); hp = max(hp, p);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lc = min(lc, __shfl_xor_sync(0xffffffffu, lc, o)); hc = max(hc, __shfl_xor_sync(0xffffffffu, hc, o));
    lp = min(lp, __shfl_xor_sync(0xffffffffu, lp, o)); hp = max(hp, __shfl_xor_sync(0xffffffffu, hp, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(out, lc); atomicMax(out + 1, hc); atomicMin(out + 2, lp); atomicMax(out + 3, hp); }
}
__global__ void k_mark(int64_t n, const long long* __restrict__ off2, long long lo_c, long long lo_p, int* __restrict__ cam_tab,
                       int* __restrict__ pt_tab) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  cam_tab[off2[2 * i] - lo_c] = 1;                 // every thread stores the same value
  pt_tab[off2[2 * i + 1] - lo_p] = 1;
}
__global__ void k_mark_extra(int m, const long long* __restrict__ extra, long long lo_c, int* __restrict__ cam_tab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) cam_tab[extra[i] - lo_c] = 1;
}
// Marked offsets must be at least `block` apart (blocks must not overlap); offsets[id] = lo + k.
__global__ void k_blocks_from_marks(int64_t range, const int* __restrict__ tab, const int* __restrict__ scan, long long lo, int block,
                                    long long* __restrict__ offsets, int* flags) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= range || tab[k] == 0) return;
  offsets[scan[k]] = lo + k;
  for (int d = 1; d < block && k + d < range; ++d) if (tab[k + d] != 0) flags[0] = 1;
}
__global__ void k_ids(int64_t n, const long long* __restrict__ off2, long long lo_c, long long lo_p, const int* __restrict__ cam_scan,
                      const int* __restrict__ pt_scan, int* __restrict__ cam_id, int* __restrict__ pt_id) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  cam_id[i] = cam_scan[off2[2 * i] - lo_c];
  pt_id[i] = pt_scan[off2[2 * i + 1] - lo_p];
}
// flags[1]: not sorted by (point, camera); flags[2]: a camera observes a point twice; pt_ptr = run boundaries of the point ids.
__global__ void k_order_and_csr(int64_t n, int n_pts, const int* __restrict__ cam_id, const int* __restrict__ pt_id, int* __restrict__ pt_ptr,
                                int* flags) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0) { pt_ptr[0] = 0; pt_ptr[n_pts] = (int)n; return; }
  const int p = pt_id[i], pp = pt_id[i - 1];
  if (p != pp) { pt_ptr[p] = (int)i; if (p < pp) flags[1] = 1; return; }
  const int c = cam_id[i], cp = cam_id[i - 1];
  if (c < cp) flags[1] = 1;
  if (c == cp) flags[2] = 1;
}
__global__ void k_max_track(int n_pts, const int* __restrict__ pt_ptr, int* flags) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_pts) { const int k = pt_ptr[p + 1] - pt_ptr[p]; if (k > T) flags[3] = 1; if (k <= 0) flags[1] = 1; }
}

// One CTA per tile: camera-sort the tile (keys are unique: camera << 8 | local index), tile-local points, slots, segment count.
__global__ void __launch_bounds__(T) k_tile_segments(int n_tiles, const int* __restrict__ tile_obs, const int* __restrict__ tile_pt,
                                                     const int* __restrict__ cam_id, const int* __restrict__ pt_id,
                                                     unsigned short* __restrict__ obs_ptl, unsigned short* __restrict__ seg_perm,
                                                     unsigned short* __restrict__ obs_slot, int* __restrict__ tile_nseg, int* max_seg) {
  __shared__ unsigned int keys[T];
  __shared__ int scan[T];
  const int t = blockIdx.x, j = threadIdx.x;
  const int ob = tile_obs[t], no = tile_obs[t + 1] - ob, pb = tile_pt[t];
  keys[j] = (j < no) ? (((unsigned)cam_id[ob + j] << 8) | (unsigned)j) : 0xffffffffu;
  if (j < no) obs_ptl[ob + j] = (unsigned short)(pt_id[ob + j] - pb);
  __syncthreads();
  for (int k = 2; k <= T; k <<= 1)
    for (int d = k >> 1; d > 0; d >>= 1) {
      const int x = j ^ d;
      if (x > j) {
        const unsigned a = keys[j], b = keys[x];
        const bool up = (j & k) == 0;
        if ((a > b) == up) { keys[j] = b; keys[x] = a; }
      }
      __syncthreads();
    }
  const bool valid = j < no;
  const int start = (valid && (j == 0 || (keys[j] >> 8) != (keys[j - 1] >> 8))) ? 1 : 0;
  scan[j] = start;
  __syncthreads();
  for (int d = 1; d < T; d <<= 1) {                       // inclusive scan of the segment starts
    const int v = (j >= d) ? scan[j - d] : 0;
    __syncthreads();
    scan[j] += v;
    __syncthreads();
  }
  if (valid) {
    const int loc = (int)(keys[j] & 255u);
    seg_perm[ob + j] = (unsigned short)loc;
    obs_slot[ob + loc] = (unsigned short)(scan[j] - 1);
  }
  if (j == T - 1) { const int ns = scan[T - 1]; tile_nseg[t] = ns; atomicMax(max_seg, ns); }
}
// Segment starts and cameras: position q of the camera-sorted order starts segment tile_seg[t] + slot when its camera differs
// from the one before it.
__global__ void __launch_bounds__(T) k_tile_segments2(const int* __restrict__ tile_obs, const int* __restrict__ tile_seg,
                                                      const int* __restrict__ cam_id, const unsigned short* __restrict__ seg_perm,
                                                      const unsigned short* __restrict__ obs_slot, int* __restrict__ seg_ptr,
                                                      int* __restrict__ seg_cam) {
  const int t = blockIdx.x, q = threadIdx.x;
  const int ob = tile_obs[t], no = tile_obs[t + 1] - ob;
  if (q >= no) return;
  const int loc = seg_perm[ob + q];
  const int cam = cam_id[ob + loc];
  if (q == 0 || cam_id[ob + seg_perm[ob + q - 1]] != cam) {
    const int s = tile_seg[t] + obs_slot[ob + loc];
    seg_ptr[s] = ob + q; seg_cam[s] = cam;
  }
}
__global__ void k_iota(int n, int* out) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = i; }
// cam_seg (segments sorted by camera, stable) -> seg_pos (its inverse) and the camera CSR from the run boundaries of the keys.
__global__ void k_cam_csr(int n_segs, int n_cams, const int* __restrict__ sorted_cam, const int* __restrict__ cam_seg, int* __restrict__ seg_pos,
                          int* __restrict__ cam_seg_ptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_segs) return;
  seg_pos[cam_seg[i]] = i;
  const int c = sorted_cam[i], cp = (i == 0) ? -1 : sorted_cam[i - 1];
  for (int k = cp + 1; k <= c; ++k) cam_seg_ptr[k] = i;     // cameras without segments on this rank get empty ranges
  if (i == n_segs - 1) for (int k = c + 1; k <= n_cams; ++k) cam_seg_ptr[k] = n_segs;
}

inline int blocks(int64_t n, int per = 256) { return (int)((n + per - 1) / per); }

struct Scan {                                               // exclusive sum of ints with cub, temp storage grown on demand
  DBuf<unsigned char> tmp;
  void operator()(const int* in, int* out, int64_t n, cudaStream_t s) {
    size_t bytes = 0;
    SK_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, s));
    if (tmp.n < bytes) tmp.alloc(bytes);
    SK_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int)n, s));
  }
};

}  // namespace

bool build_ba_layout_device(int64_t n, const int64_t* offsets2, const double* obs_xy, const std::vector<int64_t>* extra_cam_off,
                            cudaStream_t s, BaLayoutHost* Hout, BaLayoutDevice* D) {
  { const char* e = std::getenv("SKERES_LAYOUT"); if (e != nullptr && e[0] == 'h') return false; }   // development / tests
  { const char* e = std::getenv("SKERES_TILE_REC"); if (e != nullptr && e[0] == 'h') return false; } // the host record builder reads the host layout
  if (n <= 0 || n >= (int64_t)2000000000) return false;
  BaLayoutHost& H = *Hout;
  // ---- upload ------------------------------------------------------------------------------------------------
  DBuf<long long> d_off((size_t)2 * n);
  static_assert(sizeof(long long) == sizeof(int64_t), "");
  SK_CUDA(cudaMemcpyAsync(d_off.p, offsets2, sizeof(int64_t) * 2 * (size_t)n, cudaMemcpyHostToDevice, s));
  D->obs.alloc((size_t)2 * n);
  SK_CUDA(cudaMemcpyAsync(D->obs.p, obs_xy, sizeof(double) * 2 * (size_t)n, cudaMemcpyHostToDevice, s));
  // ---- ranges ------------------------------------------------------------------------------------------------
  DBuf<long long> d_mm(4);
  HBuf<long long> h_mm(4);
  h_mm.p[0] = LLONG_MAX; h_mm.p[1] = LLONG_MIN; h_mm.p[2] = LLONG_MAX; h_mm.p[3] = LLONG_MIN;
  SK_CUDA(cudaMemcpyAsync(d_mm.p, h_mm.p, 4 * sizeof(long long), cudaMemcpyHostToDevice, s));
  k_minmax<<<std::min(blocks(n), 1184), 256, 0, s>>>(n, d_off.p, d_mm.p);
  SK_CUDA(cudaMemcpyAsync(h_mm.p, d_mm.p, 4 * sizeof(long long), cudaMemcpyDeviceToHost, s));
  SK_CUDA(cudaStreamSynchronize(s));
  long long lo_c = h_mm.p[0], hi_c = h_mm.p[1];
  const long long lo_p = h_mm.p[2], hi_p = h_mm.p[3];
  if (extra_cam_off != nullptr) for (int64_t o : *extra_cam_off) { lo_c = std::min<long long>(lo_c, o); hi_c = std::max<long long>(hi_c, o); }
  const int64_t range_c = hi_c - lo_c + 1, range_p = hi_p - lo_p + 1;
  const int64_t table_cap = std::max<int64_t>(64 * n, 1 << 20);
  if (range_c > table_cap || range_p > table_cap) return false;             // sparse offsets: the host builder sorts them
  if (!(hi_c + 9 <= lo_p || hi_p + 3 <= lo_c)) return false;                // camera and point blocks interleaved: host checks overlap
  // ---- dense ids -----------------------------------------------------------------------------------------------
  DBuf<int> cam_tab((size_t)range_c), pt_tab((size_t)range_p), cam_scan((size_t)range_c + 1), pt_scan((size_t)range_p + 1);
  cam_tab.zero(s); pt_tab.zero(s);
  k_mark<<<blocks(n), 256, 0, s>>>(n, d_off.p, lo_c, lo_p, cam_tab.p, pt_tab.p);
  DBuf<long long> d_extra;
  if (extra_cam_off != nullptr && !extra_cam_off->empty()) {
    std::vector<long long> ex(extra_cam_off->begin(), extra_cam_off->end());
    d_extra.upload(ex, s);
    k_mark_extra<<<blocks((int64_t)ex.size()), 256, 0, s>>>((int)ex.size(), d_extra.p, lo_c, cam_tab.p);
    SK_CUDA(cudaStreamSynchronize(s));                                       // ex is a temporary
  }
  Scan scan;
  scan(cam_tab.p, cam_scan.p, range_c, s);
  scan(pt_tab.p, pt_scan.p, range_p, s);
  DBuf<int> flags(8);
  flags.zero(s);
  HBuf<int> h_cnt(12);
  // counts = last scan entry + last mark (both tables end with a marked entry: hi is an offset that occurs)
  SK_CUDA(cudaMemcpyAsync(h_cnt.p, cam_scan.p + (range_c - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
  SK_CUDA(cudaMemcpyAsync(h_cnt.p + 1, pt_scan.p + (range_p - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
  SK_CUDA(cudaStreamSynchronize(s));
  const int n_cams = h_cnt.p[0] + 1, n_pts = h_cnt.p[1] + 1;
  if (n_cams >= (1 << 24)) return false;
  D->cam_off.alloc((size_t)n_cams); D->pt_off.alloc((size_t)n_pts);
  k_blocks_from_marks<<<blocks(range_c), 256, 0, s>>>(range_c, cam_tab.p, cam_scan.p, lo_c, 9, D->cam_off.p, flags.p);
  k_blocks_from_marks<<<blocks(range_p), 256, 0, s>>>(range_p, pt_tab.p, pt_scan.p, lo_p, 3, D->pt_off.p, flags.p);
  DBuf<int> cam_id((size_t)n), pt_id((size_t)n);
  k_ids<<<blocks(n), 256, 0, s>>>(n, d_off.p, lo_c, lo_p, cam_scan.p, pt_scan.p, cam_id.p, pt_id.p);
  // ---- order, duplicates, point CSR, long tracks ---------------------------------------------------------------------
  D->pt_ptr.alloc((size_t)n_pts + 1);
  SK_CUDA(cudaMemsetAsync(D->pt_ptr.p, 0xff, sizeof(int) * ((size_t)n_pts + 1), s));    // -1: a point id that never starts a run
  k_order_and_csr<<<blocks(n), 256, 0, s>>>(n, n_pts, cam_id.p, pt_id.p, D->pt_ptr.p, flags.p);
  k_max_track<<<blocks(n_pts), 256, 0, s>>>(n_pts, D->pt_ptr.p, flags.p);
  std::vector<int32_t> pt_ptr((size_t)n_pts + 1);
  SK_CUDA(cudaMemcpyAsync(h_cnt.p + 4, flags.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
  SK_CUDA(cudaMemcpyAsync(pt_ptr.data(), D->pt_ptr.p, sizeof(int) * ((size_t)n_pts + 1), cudaMemcpyDeviceToHost, s));
  SK_CUDA(cudaStreamSynchronize(s));
  const int* fl = h_cnt.p + 4;
  if (fl[0] || fl[1] || fl[2] || fl[3]) return false;    // overlapping blocks / unsorted / duplicates / long tracks: host builder
  // ---- tiles (host: one sequential pass over the point CSR) ------------------------------------------------------------
  std::vector<int32_t> tile_obs, tile_pt, tile_np;
  tile_obs.reserve((size_t)n / 200 + 16); tile_pt.reserve((size_t)n / 200 + 16); tile_np.reserve((size_t)n / 200 + 16);
  int32_t cur = 0, first_pt = 0, max_pt_tile = 0;
  for (int32_t p = 0; p < n_pts; ++p) {
    const int32_t k = pt_ptr[p + 1] - pt_ptr[p];
    if (cur + k > kTileObs) {
      tile_obs.push_back(pt_ptr[first_pt]); tile_pt.push_back(first_pt); tile_np.push_back(p - first_pt);
      max_pt_tile = std::max(max_pt_tile, p - first_pt);
      cur = 0;
    }
    if (cur == 0) first_pt = p;
    cur += k;
  }
  if (cur > 0) { tile_obs.push_back(pt_ptr[first_pt]); tile_pt.push_back(first_pt); tile_np.push_back(n_pts - first_pt); max_pt_tile = std::max(max_pt_tile, n_pts - first_pt); }
  const int n_tiles = (int)tile_obs.size();
  tile_obs.push_back((int32_t)n); tile_pt.push_back(n_pts);
  D->tile_obs.upload(tile_obs, s); D->tile_pt.upload(tile_pt, s); D->tile_np.upload(tile_np, s);   // regular tiles only: enc = np
  // ---- tile-local segments ---------------------------------------------------------------------------------------------
  D->obs_ptl.alloc((size_t)n); D->seg_perm.alloc((size_t)n); D->obs_slot.alloc((size_t)n);
  DBuf<int> tile_nseg((size_t)n_tiles + 1);
  tile_nseg.zero(s);
  D->tile_seg.alloc((size_t)n_tiles + 1);
  int* max_seg = flags.p + 4;
  k_tile_segments<<<n_tiles, T, 0, s>>>(n_tiles, D->tile_obs.p, D->tile_pt.p, cam_id.p, pt_id.p, D->obs_ptl.p, D->seg_perm.p, D->obs_slot.p,
                                        tile_nseg.p, max_seg);
  scan(tile_nseg.p, D->tile_seg.p, (int64_t)n_tiles + 1, s);
  SK_CUDA(cudaMemcpyAsync(h_cnt.p + 2, D->tile_seg.p + n_tiles, sizeof(int), cudaMemcpyDeviceToHost, s));
  SK_CUDA(cudaMemcpyAsync(h_cnt.p + 3, max_seg, sizeof(int), cudaMemcpyDeviceToHost, s));
  SK_CUDA(cudaStreamSynchronize(s));                        // also: tile_obs / tile_pt / tile_np (host temporaries) are uploaded
  const int n_segs = h_cnt.p[2], max_seg_tile = h_cnt.p[3];
  D->seg_ptr.alloc((size_t)n_segs + 1); D->seg_cam.alloc((size_t)n_segs);
  k_tile_segments2<<<n_tiles, T, 0, s>>>(D->tile_obs.p, D->tile_seg.p, cam_id.p, D->seg_perm.p, D->obs_slot.p, D->seg_ptr.p, D->seg_cam.p);
  const int n32 = (int)n;
  SK_CUDA(cudaMemcpyAsync(D->seg_ptr.p + n_segs, &n32, sizeof(int), cudaMemcpyHostToDevice, s));
  // ---- camera -> segments -----------------------------------------------------------------------------------------------
  D->cam_seg.alloc((size_t)n_segs); D->seg_pos.alloc((size_t)n_segs); D->cam_seg_ptr.alloc((size_t)n_cams + 1);
  {
    DBuf<int> iota((size_t)n_segs), sorted_cam((size_t)n_segs);
    k_iota<<<blocks(n_segs), 256, 0, s>>>(n_segs, iota.p);
    int bits = 1;
    while ((1 << bits) < n_cams) ++bits;
    size_t bytes = 0;
    SK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, D->seg_cam.p, sorted_cam.p, iota.p, D->cam_seg.p, n_segs, 0, bits, s));
    DBuf<unsigned char> tmp(bytes);
    SK_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, D->seg_cam.p, sorted_cam.p, iota.p, D->cam_seg.p, n_segs, 0, bits, s));
    k_cam_csr<<<blocks(n_segs), 256, 0, s>>>(n_segs, n_cams, sorted_cam.p, D->cam_seg.p, D->seg_pos.p, D->cam_seg_ptr.p);
    // what the host still needs: the camera table (offsets) and, for the multi-GPU contribution masks, the camera CSR
    H.cam_offset.resize((size_t)n_cams); H.cam_seg_ptr.resize((size_t)n_cams + 1);
    std::vector<long long> co((size_t)n_cams);
    SK_CUDA(cudaMemcpyAsync(co.data(), D->cam_off.p, sizeof(long long) * (size_t)n_cams, cudaMemcpyDeviceToHost, s));
    SK_CUDA(cudaMemcpyAsync(H.cam_seg_ptr.data(), D->cam_seg_ptr.p, sizeof(int) * ((size_t)n_cams + 1), cudaMemcpyDeviceToHost, s));
    SK_CUDA(cudaStreamSynchronize(s));                      // temporaries of this scope are released after the sort has run
    for (int c = 0; c < n_cams; ++c) H.cam_offset[c] = co[c];
  }
  check_launch("build_ba_layout_device");
  H.n_obs = (int32_t)n; H.n_pts = n_pts; H.n_cams = n_cams; H.n_tiles = n_tiles; H.n_segs = n_segs;
  H.max_seg_tile = max_seg_tile; H.max_pt_tile = max_pt_tile; H.n_giant = 0; H.n_chunks = 0; H.input_was_sorted = true;
  D->valid = true;
  return true;
}

}  // namespace sk
