// pcg_device.cuh — device side of the preconditioned conjugate gradients on the reduced camera system
// (ConjugateGradientsSolver::Solve), shared by the kernel sequence (pcg_kernels.cu) and the fused solve
// (pcg_fused.cu: the whole loop in one persistent kernel).  Per translation unit (anonymous namespace); every rounding of
// the vector updates is spelled out (__fma_rn / __dmul_rn / __dadd_rn) so that the two forms produce the same bits
// whatever the compiler would contract in either context.
//
// Work is organised in VIRTUAL BLOCKS of WPB = 8 cameras: block vb owns cameras [8 vb, 8 vb + 8) and one slot of every
// partial-sum array (p.q, x.(b + r), r.z).  The kernel sequence runs one CTA per virtual block; the fused solve lets each of
// its persistent CTAs walk the blocks vb = blockIdx.x, blockIdx.x + gridDim.x, ...  Sums over the slots are taken in slot
// order by fixed trees (sum_fixed), so the result does not depend on who computed a slot.
#pragma once
#include "ba_kernels.cuh"
#include "comm.cuh"
#include "lm_kernels.cuh"

namespace sk {
namespace {

constexpr int WPB = 8;                    // cameras per virtual block

__device__ __forceinline__ double block_sum_fixed(double x, double* red) {   // 256 threads
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = x;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < (int)(blockDim.x >> 5)) ? red[l] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
  }
  return r;                               // valid in thread 0
}

// N block sums at once: the same per-thread order and the same trees as N calls of block_sum_fixed, one barrier pair.
// red: [N][8].  Valid in thread 0.
template <int N>
__device__ __forceinline__ void block_sum_fixed_n(double (&x)[N], double* red) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int n = 0; n < N; ++n) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x[n] += __shfl_down_sync(0xffffffffu, x[n], o);
  }
  __syncthreads();
  if (l == 0) {
#pragma unroll
    for (int n = 0; n < N; ++n) red[n * 8 + w] = x[n];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int n = 0; n < N; ++n) {
      double r = (l < (int)(blockDim.x >> 5)) ? red[n * 8 + l] : 0.0;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
      x[n] = r;
    }
  }
}

__device__ __forceinline__ double sum_fixed(const double* part, int n, double* red) {
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += __ldcg(part + i);   // L2: partials may come from other CTAs of this launch
  return block_sum_fixed(a, red);
}

// Same value in every thread of the CTA (used where each CTA needs the scalar itself).
__device__ __forceinline__ double sum_fixed_all(const double* part, int n, double* red, double* bcast) {
  const double s = sum_fixed(part, n, red);
  if (threadIdx.x == 0) *bcast = s;
  __syncthreads();
  return *bcast;
}

__device__ __forceinline__ bool zero_or_inf(double x) { return x == 0.0 || isinf(x); }

// ---- peer window (comm.cuh): system-scope flag / data accesses over NVLink peer memory ----------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];\n" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
// Consumer side: every CTA waits until all ranks have published exchange `seq` (one polling thread per rank).  The wait is
// bounded by wall-clock time (win.timeout_ns on %globaltimer; default 60 s, SKERES_PEER_TIMEOUT_S): ranks are separate
// processes and a peer may legitimately stall for a while (lazy module load, a paused host thread), but a rank that died must
// not hang the GPU.  Returns false -- in every thread of the CTA -- when this exchange, or an earlier one, timed out: the error
// word is sticky, the caller then marks the solve LIN_FATAL instead of consuming unpublished data, and the flag reaches the
// other ranks with the next scalar allreduce (lm_kernels.cuh: SB_FLAG_LIN).
__device__ __forceinline__ bool peer_wait(const PeerWindow& win, int parity, unsigned long long seq) {
  __shared__ int failed;
  if ((int)threadIdx.x < win.world) {
    const unsigned long long* f = win.flags[win.rank] + threadIdx.x * 2 + parity;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(f) < seq) {
      __nanosleep(40);
      if (*(volatile int*)win.error != 0) break;
      if (global_ns() - t0 > win.timeout_ns) { atomicExch(win.error, 1); break; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) failed = *(volatile int*)win.error;
  __syncthreads();
  return failed == 0;
}
// y[e] of camera c summed over the contributions of the ranks that hold observations of c, in rank order (the same bits on
// every rank).  Points are partitioned, so a camera is seen by the few ranks whose points it observes: reading only those
// windows (cam_mask) keeps the gather's NVLink volume at ~(cameras touched per rank) instead of (ranks x all cameras) --
// measured round 1, N = 8: the all-windows gather was the part of the PCG iteration that grew with N (25 -> 94 us).
__device__ __forceinline__ double peer_gather(const PeerWindow& win, int parity, int c, size_t e) {
  const unsigned mask = win.cam_mask != nullptr ? (unsigned)win.cam_mask[c] : 0xffu;
  double x[kMaxPeers];
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r)                     // all loads in flight before the first add (remote latency once, not world times)
    x[r] = (r < win.world && ((mask >> r) & 1u)) ? ld_relaxed_sys(win.data[r] + (size_t)parity * win.stride + e) : 0.0;
  double acc = 0.0;
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r) if (r < win.world && ((mask >> r) & 1u)) acc += x[r];
  return acc;
}
// This rank's contribution to exchange `seq` is complete (the caller has ordered the writes of all its CTAs before this call):
// threads 0 .. world-1 of ONE CTA store the sequence number into the flag word of every rank.
__device__ __forceinline__ void peer_store_flags(const PeerWindow& win, int parity, unsigned long long seq) {
  if ((int)threadIdx.x < win.world) {
    __threadfence_system();
    st_release_sys(win.flags[threadIdx.x] + win.rank * 2 + parity, seq);
  }
}

// Finishes iteration st->iter (if not done yet) and, unless finish_only, opens the next one.  Runs on one whole CTA of 256
// threads: as its own kernel (first iteration of a solve), at the tail of k_pcg_update / k_pcg_resid2 (executed by the CTA
// that publishes its partial sums last), or by EVERY CTA of the fused solve on its own copy of the state (same inputs, same
// fixed-order sums: the same decision everywhere).  Thread 0 writes *st; callers synchronise before reading it.
__device__ __forceinline__ void pcg_head_step(PcgDev* st, const double* part_rho, const double* part_pq, const double* part_Q, int nparts,
                                              PcgParams prm, int finish_only) {
  if (st->active == 0) return;
  const int it = st->iter;
  const bool need_finish = it >= 1 && st->pad_ != it;
  // p.q and x.(b + r) of the iteration being finished, r.z of the one being opened: one pass, one barrier pair (each sum
  // keeps the order sum_fixed gives it)
  __shared__ double red3[3 * 8];
  double acc[3] = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
    if (need_finish) { acc[0] += __ldcg(part_pq + i); acc[1] += __ldcg(part_Q + i); }
    if (!finish_only) acc[2] += __ldcg(part_rho + i);
  }
  block_sum_fixed_n<3>(acc, red3);
  const double pq = acc[0], xbr = acc[1], rho = acc[2];
  if (threadIdx.x != 0) return;
  if (need_finish) {
    st->pad_ = it;
    st->pq = pq;
    if (!(pq > 0.0) || isinf(pq)) { st->active = 0; st->termination = LIN_NO_CONVERGENCE; return; }   // indefinite: x was not updated
    st->alpha = st->rho / pq;
    if (isinf(st->alpha)) { st->active = 0; st->termination = LIN_FAILURE; return; }
    const double Q1 = -1.0 * xbr;
    st->Q1 = Q1;
    const double zeta = it * (Q1 - st->Q0) / Q1;
    if (zeta < prm.q_tolerance && it >= prm.min_iterations) { st->active = 0; st->termination = LIN_SUCCESS; return; }
    st->Q0 = Q1;
    if (it >= prm.max_iterations) { st->active = 0; st->termination = LIN_NO_CONVERGENCE; return; }
  }
  if (finish_only) return;
  st->last_rho = st->rho;
  st->rho = rho;
  st->iter = it + 1;
  if (zero_or_inf(rho) || !(rho == rho)) { st->active = 0; st->termination = LIN_FAILURE; return; }
  if (it + 1 > 1) {
    st->beta = rho / st->last_rho;
    if (zero_or_inf(st->beta)) { st->active = 0; st->termination = LIN_FAILURE; return; }
  }
}

// acc = seg_y[t][k] + seg_y[t + 3 WPC][k] + ... in that order, sixteen partials in flight (the adds stay one chain in t order:
// only the loads are batched -- a camera of the Venice shape owns ~190 partials, 63 per segment lane, and every batch costs
// one L2 round trip).  The product kernels store a segment's partial at its camera-major position (BaDev::seg_pos), so a
// camera's partials are the contiguous rows [cam_seg_ptr[c], cam_seg_ptr[c + 1]) in tile order: no index is read on the way.
// L2 loads: the partials come from other CTAs.
template <int WPC>
__device__ __forceinline__ double walk_segments(const double* seg_y, int t, int e, int k) {
  constexpr int S = 3 * WPC;
  constexpr int B = 16;
  double acc = 0.0;
  for (; t + (B - 1) * S < e; t += B * S) {
    double x[B];
#pragma unroll
    for (int i = 0; i < B; ++i) x[i] = __ldcg(seg_y + (size_t)(t + i * S) * 9 + k);
#pragma unroll
    for (int i = 0; i < B; ++i) acc += x[i];
  }
  if (t + 3 * S < e) {
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = __ldcg(seg_y + (size_t)(t + i * S) * 9 + k);
#pragma unroll
    for (int i = 0; i < 4; ++i) acc += x[i];
    t += 4 * S;
  }
  for (; t < e; t += S) acc += __ldcg(seg_y + (size_t)t * 9 + k);
  return acc;
}

// WPC warps per camera (blockDim.x == WPB * WPC * 32): y[c] = fixed-order sum of the camera's segment partials -- each of the
// camera's WPC warps takes every WPC-th group of three segments (lanes: 3 segment lanes x 9 components), the warp sums are
// added in warp order.  A camera owns ~190 (tile, camera) partials on the Venice shape; WPC is chosen per problem from the
// average number of partials per camera (pcg_wpc).  Result in lanes 0..8 of the camera's FIRST warp (sub == 0); ends with a
// CTA barrier inside.  part: [WPB][WPC][9] shared.
template <int WPC>
__device__ __forceinline__ double camera_partial_sum(const BaDev& L, const double* seg_y, int c, double (*part)[WPC][9]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cl = warp / WPC, sub = warp % WPC;
  const int k = lane % 9, j = lane / 9;               // 3 segment lanes x 9 components; lanes 27..31 idle
  double acc = 0.0;
  if (c < L.n_cams) {
    if (lane < 27) acc = walk_segments<WPC>(seg_y, L.cam_seg_ptr[c] + sub * 3 + j, L.cam_seg_ptr[c + 1], k);
    const double a1 = __shfl_down_sync(0xffffffffu, acc, 9), a2 = __shfl_down_sync(0xffffffffu, acc, 18);
    acc = (acc + a1) + a2;
    if (WPC > 1 && lane < 9) part[cl][sub][lane] = acc;
  }
  if (WPC > 1) {
    __syncthreads();
    if (c < L.n_cams && sub == 0 && lane < 9) {
      acc = part[cl][0][lane];
#pragma unroll
      for (int w = 1; w < WPC; ++w) acc += part[cl][w][lane];
    }
    __syncthreads();                                   // part may be rewritten by the next virtual block
  }
  return acc;
}

// Virtual block vb of "reduce": p = z + beta p_old (iteration 1: p = z); q = y + D^2 p (stored in z, as Ceres does); the
// block's p.q into part_pq[vb].  y = the camera's summed segment partials, or y_in (already reduced / allreduced), or --
// peer -- the ranks' window contributions added in rank order.
template <int WPC>
__device__ __forceinline__ void pcg_reduce_block(const BaDev& L, int vb, const double* seg_y, const double* y_in, const double* D,
                                                 double* z, double* p, double* part_pq, int iter, double beta,
                                                 const PeerWindow& win, int parity) {
  __shared__ double part[WPB][WPC][9];
  __shared__ double red[WPB];
  const bool peer = win.world > 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cl = warp / WPC, sub = warp % WPC;
  const int c = vb * WPB + cl;
  double acc = 0.0;
  if (y_in == nullptr && !peer) acc = camera_partial_sum<WPC>(L, seg_y, c, part);
  double pq = 0.0;
  if (c < L.n_cams && sub == 0 && lane < 9) {
    const size_t e = (size_t)c * 9 + lane;
    if (peer) acc = peer_gather(win, parity, c, e);
    else if (y_in != nullptr) acc = __ldcg(y_in + e);
    const double zk = __ldcg(z + e);
    const double pk = (iter == 1) ? zk : __fma_rn(beta, __ldcg(p + e), zk);
    const double d = D[e];
    const double qk = __fma_rn(__dmul_rn(d, d), pk, acc);
    p[e] = pk; z[e] = qk;
    pq = __dmul_rn(pk, qk);
  }
  if (sub == 0) {                                      // camera leaders: p.q of the camera into red[], then the block's 8 in order
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) pq += __shfl_down_sync(0xffffffffu, pq, o);   // lanes 0..8 hold data: 16-wide tree covers them
    if (lane == 0) red[cl] = pq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < WPB; ++w) s += red[w];
    part_pq[vb] = s;
  }
  __syncthreads();                                     // red may be rewritten by the next virtual block
}

// z = M^-1 r for one camera: lanes 0..8 hold r, every lane of the warp takes part in the shuffles.
__device__ __forceinline__ double precondition(const double* Minv, int c, int lane, double rk) {
  if (Minv == nullptr) return rk;
  double zk = 0.0;
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const double rj = __shfl_sync(0xffffffffu, rk, j);
    if (lane < 9) zk = __fma_rn(Minv[(size_t)c * 81 + lane * 9 + j], rj, zk);
  }
  return zk;
}

// Virtual block vb of "update" (256 threads, warp per camera): x += alpha p; then, unless `recompute` (r is rebuilt from
// b - S x after one more product), r -= alpha q (q lives in z), the block's x.(b + r) into part_Q[vb], z = M^-1 r and the
// block's r.z into part_rho[vb].  go == false (p.q not positive, alpha infinite): Ceres breaks before touching x; the slots
// are still written (zeros) so that the head step can record why.
__device__ __forceinline__ void pcg_update_block(int n_cams, int vb, const double* Minv, const double* b, double* x, const double* p,
                                                 double* r, double* z, double alpha, bool go, int recompute, double* part_Q,
                                                 double* part_rho) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = vb * WPB + warp;
  double qsum = 0.0, rz = 0.0;
  if (go && c < n_cams) {
    const size_t e = (size_t)c * 9 + (lane < 9 ? lane : 0);
    double rk = 0.0;
    if (lane < 9) {
      const double xk = __fma_rn(alpha, __ldcg(p + e), __ldcg(x + e));
      x[e] = xk;
      if (!recompute) {
        rk = __fma_rn(-alpha, __ldcg(z + e), __ldcg(r + e));
        r[e] = rk;
        qsum = __dmul_rn(xk, __dadd_rn(b[e], rk));
      }
    }
    if (!recompute) {
      const double zk = precondition(Minv, c, lane, rk);
      if (lane < 9) { z[e] = zk; rz = __dmul_rn(rk, zk); }
    }
  }
  if (!recompute) {
    __shared__ double red2[2 * 8];
    double s12[2] = {qsum, rz};
    block_sum_fixed_n<2>(s12, red2);
    if (threadIdx.x == 0) { part_Q[vb] = s12[0]; part_rho[vb] = s12[1]; }
    __syncthreads();                                   // red2 may be rewritten by the next virtual block
  }
}

// Virtual block vb of the residual reset (256 threads, warp per camera): r = b - (y + D^2 x) with y = the summed segment
// partials of S_local x (or y_in, or the peer windows); then the x.(b + r) slot, z = M^-1 r and the r.z slot.
__device__ __forceinline__ void pcg_resid_block(const BaDev& L, int vb, const double* seg_y, const double* y_in, const double* D,
                                                const double* Minv, const double* b, const double* x, double* r, double* z,
                                                double* part_Q, double* part_rho, const PeerWindow& win, int parity) {
  __shared__ double red2[2 * 8];
  const bool peer = win.world > 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = vb * WPB + warp;
  double qsum = 0.0, rz = 0.0;
  double acc = 0.0;
  if (y_in == nullptr && !peer) acc = camera_partial_sum<1>(L, seg_y, c, nullptr);
  if (c < L.n_cams) {
    const size_t e = (size_t)c * 9 + (lane < 9 ? lane : 0);
    double rk = 0.0;
    if (lane < 9) {
      if (peer) acc = peer_gather(win, parity, c, e);
      else if (y_in != nullptr) acc = __ldcg(y_in + e);
      const double d = D[e], xk = __ldcg(x + e);
      rk = __dsub_rn(b[e], __fma_rn(__dmul_rn(d, d), xk, acc));
      r[e] = rk;
      qsum = __dmul_rn(xk, __dadd_rn(b[e], rk));
    }
    const double zk = precondition(Minv, c, lane, rk);
    if (lane < 9) { z[e] = zk; rz = __dmul_rn(rk, zk); }
  }
  double s12[2] = {qsum, rz};
  block_sum_fixed_n<2>(s12, red2);
  if (threadIdx.x == 0) { part_Q[vb] = s12[0]; part_rho[vb] = s12[1]; }
  __syncthreads();
}

// Virtual block vb: y[c] = the camera's summed segment partials, in exactly the order pcg_reduce_block uses.
template <int WPC>
__device__ __forceinline__ void cam_reduce9_block(const BaDev& L, int vb, const double* seg_y, double* y) {
  __shared__ double part[WPB][WPC][9];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cl = warp / WPC, sub = warp % WPC;
  const int c = vb * WPB + cl;
  const double acc = camera_partial_sum<WPC>(L, seg_y, c, part);
  if (c < L.n_cams && sub == 0 && lane < 9) y[(size_t)c * 9 + lane] = acc;
}

// ---- three virtual blocks per round, by lane groups ---------------------------------------------------------------------------
// At many cameras (13,682 of the Final shape, 14,224 of eight Venice scenes) the fused solve's persistent grid needs seven rounds
// of virtual blocks per vector phase, each a chain of dependent L2 round trips, while a warp uses 9 of its 32 lanes for a camera.
// Here lanes 0-8, 9-17 and 18-26 of warp w serve camera w of THREE blocks (va, va + step, va + 2 step; a block >= nparts is
// skipped): the same arithmetic per camera, the same trees (a shuffle that would reach into the neighbouring group is masked
// -- the one-block form adds the zeros of its idle lanes there), the same slots: the same bits, a third of the rounds, no extra
// registers.  256 threads.  Used where no walk over segment partials is involved: the update, and the reduce when the camera
// sums come from the peer windows.
__device__ __forceinline__ double group_sum9(double v, int i) {       // sum over the 9 lanes of a group, valid in its first lane
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) { const double t = __shfl_down_sync(0xffffffffu, v, o); if (i + o < 9) v += t; }
  return v;
}
// slot[vb_g] = the 8 warps' group sums of block g combined as block_sum_fixed combines its 8 warp totals.  red: [3][N][8].
template <int N>
__device__ __forceinline__ void block3_slots(const double (&v)[N], int g, int i, double (*red)[N][8], const int (&vb)[3], int nparts,
                                             double* const (&out)[N]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (g < 3 && i == 0) {
#pragma unroll
    for (int n = 0; n < N; ++n) red[g][n][warp] = v[n];
  }
  __syncthreads();
  if (warp < 3) {
#pragma unroll
    for (int n = 0; n < N; ++n) {
      double r = (lane < 8) ? red[warp][n][lane] : 0.0;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
      if (lane == 0 && vb[warp] < nparts) out[n][vb[warp]] = r;
    }
  }
  __syncthreads();                                      // red may be rewritten by the next round
}

__device__ __forceinline__ void pcg_reduce_block3_peer(const BaDev& L, int va, int step, int nparts, const double* D, double* z, double* p,
                                                       double* part_pq, int iter, double beta, const PeerWindow& win, int parity) {
  __shared__ double red[3][1][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / 9, i = lane - 9 * g;
  const int vb[3] = {va, va + step, va + 2 * step};
  const int v = g < 3 ? vb[g] : nparts;
  const int c = v * WPB + warp;
  double pq[1] = {0.0};
  if (g < 3 && v < nparts && c < L.n_cams) {
    const size_t e = (size_t)c * 9 + i;
    const double acc = peer_gather(win, parity, c, e);
    const double zk = __ldcg(z + e);
    const double pk = (iter == 1) ? zk : __fma_rn(beta, __ldcg(p + e), zk);
    const double d = D[e];
    const double qk = __fma_rn(__dmul_rn(d, d), pk, acc);
    p[e] = pk; z[e] = qk;
    pq[0] = __dmul_rn(pk, qk);
  }
  pq[0] = group_sum9(pq[0], i);
  double* const out[1] = {part_pq};
  block3_slots<1>(pq, g, i, red, vb, nparts, out);
}

__device__ __forceinline__ void pcg_update_block3(int n_cams, int va, int step, int nparts, const double* Minv, const double* b, double* x,
                                                  const double* p, double* r, double* z, double alpha, bool go, int recompute,
                                                  double* part_Q, double* part_rho) {
  __shared__ double red[3][2][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / 9, i = lane - 9 * g;
  const int vb[3] = {va, va + step, va + 2 * step};
  const int v = g < 3 ? vb[g] : nparts;
  const int c = v * WPB + warp;
  const bool on = go && g < 3 && v < nparts && c < n_cams;
  double s12[2] = {0.0, 0.0};
  double rk = 0.0;
  const size_t e = on ? (size_t)c * 9 + i : 0;
  if (on) {
    const double xk = __fma_rn(alpha, __ldcg(p + e), __ldcg(x + e));
    x[e] = xk;
    if (!recompute) {
      rk = __fma_rn(-alpha, __ldcg(z + e), __ldcg(r + e));
      r[e] = rk;
      s12[0] = __dmul_rn(xk, __dadd_rn(b[e], rk));
    }
  }
  if (!recompute) {
    double zk = rk;
    if (Minv != nullptr) {
      zk = 0.0;
      const int base = (g < 3 ? g : 0) * 9;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const double rj = __shfl_sync(0xffffffffu, rk, base + j);
        if (on) zk = __fma_rn(Minv[(size_t)c * 81 + i * 9 + j], rj, zk);
      }
    }
    if (on) { z[e] = zk; s12[1] = __dmul_rn(rk, zk); }
    s12[0] = group_sum9(s12[0], i); s12[1] = group_sum9(s12[1], i);
    double* const out[2] = {part_Q, part_rho};
    block3_slots<2>(s12, g, i, red, vb, nparts, out);
  }
}

}  // namespace
}  // namespace sk
