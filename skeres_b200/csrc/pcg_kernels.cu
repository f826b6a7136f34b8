// pcg_kernels.cu — fused preconditioned conjugate gradients on the reduced camera system
// (ConjugateGradientsSolver::Solve) with every scalar kept on the device.
//
// One PCG iteration is three launches (the first iteration of a solve has k_pcg_head in front):
//   [head step]   (1 CTA)  finishes the previous iteration (Q-based termination test, max iterations,
//                          indefiniteness), then rho = r.z, beta, iteration counter; executed at the tail of
//                          k_pcg_update / k_pcg_resid2 by the CTA that publishes its partial sums last
//   k_ba_matvec   (tiles)  implicit Schur product of p = z + beta p_old, formed on the fly (ba_kernels.cu)
//   k_pcg_reduce  (warp per camera) p = z + beta p_old (stored), q = sum of the camera's segment partials
//                          + D^2 p, and the p.q partials
//   k_pcg_update  (warp per camera) alpha = rho / p.q; x += alpha p; r -= alpha q; Q partials;
//                          z = M^-1 r (SchurJacobi block) and the r.z partials of the NEXT iteration
// Every residual_reset_period-th iteration r is recomputed as b - S x (one more matvec + k_pcg_resid2).
// All sums are fixed-order (per-warp partials reduced by one CTA), hence bit-reproducible.
// Multi-GPU (points partitioned, cameras replicated): a fourth launch, k_cam_reduce9_warp, reduces this rank's partials per
// camera into its NVLink peer window and publishes them; k_pcg_reduce / k_pcg_resid2 then wait for every rank's window and
// add them in rank order (comm.cuh: PeerWindow) -- the exchange is part of the kernels on either side, not a collective call
// between them.  Without peer mapping the reduced vector goes through an NCCL allreduce instead (y_in).
#include "pcg_kernels.cuh"

#include <cstdlib>

namespace sk {

namespace {

constexpr int WPB = 8;                    // warps (= cameras) per CTA

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

__device__ double sum_fixed(const double* part, int n, double* red) {
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += __ldcg(part + i);   // L2: partials may come from other CTAs of this launch
  return block_sum_fixed(a, red);
}

// Same value in every thread of the CTA (used where each CTA needs the scalar itself).
__device__ double sum_fixed_all(const double* part, int n, double* red, double* bcast) {
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
// Consumer side: every CTA waits until all ranks have published exchange `seq` (one polling thread per rank).  The wait is
// bounded by wall-clock time (win.timeout_ns on %globaltimer; default 60 s, SKERES_PEER_TIMEOUT_S): ranks are separate
// processes and a peer may legitimately stall for a while (lazy module load, a paused host thread), but a rank that died must
// not hang the GPU.  Returns false -- in every thread of the CTA -- when this exchange, or an earlier one, timed out: the error
// word is sticky, the caller then marks the solve LIN_FATAL instead of consuming unpublished data, and the flag reaches the
// other ranks with the next scalar allreduce (lm_kernels.cuh: SB_FLAG_LIN).
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
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
__device__ __forceinline__ void pcg_mark_fatal(PcgDev* st) {
  if (blockIdx.x == 0 && threadIdx.x == 0) { st->active = 0; st->termination = LIN_FATAL; }
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
// Producer side, at the end of the kernel that wrote this rank's contribution: the CTA that finishes last publishes `seq` to
// every rank.  Executed unconditionally (also by a kernel whose guard says "skip") so that the ranks' exchanges stay paired.
__device__ __forceinline__ void peer_publish(const PeerWindow& win, int parity, unsigned long long seq) {
  __shared__ int last;
  __syncthreads();                                          // the CTA's writes happen before thread 0's fence (cumulativity)
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(win.done_count, 1u);
    last = (t == gridDim.x - 1) ? 1 : 0;
    if (last) *win.done_count = 0;
  }
  __syncthreads();
  if (last && (int)threadIdx.x < win.world) {
    __threadfence_system();
    st_release_sys(win.flags[threadIdx.x] + win.rank * 2 + parity, seq);
  }
}

__global__ void k_pcg_begin(int n_cams, const double* __restrict__ rhs, const double* __restrict__ Minv, double* __restrict__ x,
                            double* __restrict__ r, double* __restrict__ z, double* __restrict__ part_bb, double* __restrict__ part_rho) {
  // x0 = 0 => r = b - S*0 = b;  z = M^-1 r;  partials of |b|^2 and r.z
  __shared__ double red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * WPB + warp;
  double bb = 0.0, rz = 0.0;
  if (c < n_cams) {
    const double rk = (lane < 9) ? rhs[(size_t)c * 9 + lane] : 0.0;
    double zk = 0.0;
    if (Minv != nullptr) {
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const double rj = __shfl_sync(0xffffffffu, rk, j);
        if (lane < 9) zk += Minv[(size_t)c * 81 + lane * 9 + j] * rj;
      }
    } else zk = rk;
    if (lane < 9) { x[(size_t)c * 9 + lane] = 0.0; r[(size_t)c * 9 + lane] = rk; z[(size_t)c * 9 + lane] = zk; bb = rk * rk; rz = rk * zk; }
  }
  const double s1 = block_sum_fixed(bb, red), s2 = block_sum_fixed(rz, red);
  if (threadIdx.x == 0) { part_bb[blockIdx.x] = s1; part_rho[blockIdx.x] = s2; }
}

__global__ void k_pcg_start2(PcgDev* st, const double* part_bb, int nparts, int* lin_error, const double* global_lin_flag) {
  __shared__ double red[8];
  const double bb = sum_fixed(part_bb, nparts, red);
  if (threadIdx.x != 0) return;
  st->norm_b = sqrt(bb);
  st->rho = 1.0; st->last_rho = 1.0; st->pq = 0.0; st->alpha = 0.0; st->beta = 0.0;
  st->Q0 = 0.0; st->Q1 = 0.0;                                   // Q0 = -x.(b + r) with x = 0
  st->iter = 0; st->active = 1; st->termination = LIN_NO_CONVERGENCE; st->pad_ = 0;   // pad_ = last finished iteration
  st->done_count = 0; st->pad2_ = 0;
  if (global_lin_flag != nullptr && *global_lin_flag != 0.0) *lin_error |= 1;     // failed on some rank = failed for all
  if (*lin_error) { st->active = 0; st->termination = LIN_FAILURE; }
  else if (st->norm_b == 0.0) { st->active = 0; st->termination = LIN_SUCCESS; }
}

// Finishes iteration st->iter (if not done yet) and, unless finish_only, opens the next one.  Runs on one whole CTA of 256
// threads: as its own kernel (first iteration of a solve) or at the tail of k_pcg_update / k_pcg_resid2, executed by the CTA
// that publishes its partial sums last (pcg_last_block) -- same inputs, same fixed-order sums, one launch less per iteration.
__device__ void pcg_head_step(PcgDev* st, const double* part_rho, const double* part_pq, const double* part_Q, int nparts,
                              PcgParams prm, int finish_only, double* red) {
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

__global__ void k_pcg_head(PcgDev* st, const double* part_rho, const double* part_pq, const double* part_Q, int nparts,
                           PcgParams prm, int finish_only) {
  __shared__ double red[8];
  pcg_head_step(st, part_rho, part_pq, part_Q, nparts, prm, finish_only, red);
}

// True in every thread of exactly one CTA of the grid: the one whose thread 0 arrives last.  Callers publish their results
// before the call; the fences make them visible to the last CTA.
__device__ bool pcg_last_block(PcgDev* st) {
  __shared__ int last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&st->done_count, 1u);
    last = (t == gridDim.x - 1) ? 1 : 0;
    if (last) st->done_count = 0;
  }
  __syncthreads();
  if (last) __threadfence();
  return last != 0;
}

// WPC warps per camera: y = fixed-order sum of the camera's segment partials (or y_in when already reduced / allreduced);
// p = z + beta p_old (iteration 1: p = z); q = y + D^2 p (stored in z, as Ceres does); p.q.
// A camera owns ~190 (tile, camera) partials on the Venice shape; with one warp (three segment lanes) the walk over them is
// a 60-step chain of dependent L2 loads and was the longest of the small PCG kernels.  Each of the camera's WPC warps takes
// every WPC-th group of three segments, the warp sums are added in warp order.
// WPC is chosen per problem from the average number of partials per camera (pcg_wpc): 4 at ~190 (one Venice-sized share on
// one GPU), 2 or 1 when the cameras of a larger problem are spread over more ranks and each rank holds few partials per camera
// -- four warps walking two partials each made these kernels six waves of near-empty CTAs at 14,224 cameras.

// acc = seg_y[t][k] + seg_y[t + 3 WPC][k] + ... in that order, four partials in flight.  The product kernels store a segment's
// partial at its camera-major position (BaDev::seg_pos), so a camera's partials are the contiguous rows [cam_seg_ptr[c],
// cam_seg_ptr[c + 1]) in tile order: no index is read on the way.
template <int WPC>
__device__ __forceinline__ double walk_segments(const BaDev& L, const double* __restrict__ seg_y, int t, int e, int k) {
  constexpr int S = 3 * WPC;
  double acc = 0.0;
  for (; t + 3 * S < e; t += 4 * S) {
    const double x0 = seg_y[(size_t)t * 9 + k], x1 = seg_y[(size_t)(t + S) * 9 + k], x2 = seg_y[(size_t)(t + 2 * S) * 9 + k], x3 = seg_y[(size_t)(t + 3 * S) * 9 + k];
    acc += x0; acc += x1; acc += x2; acc += x3;
  }
  for (; t < e; t += S) acc += seg_y[(size_t)t * 9 + k];
  return acc;
}

template <int WPC>
__global__ void __launch_bounds__(WPB * WPC * 32) k_pcg_reduce(BaDev L, const double* __restrict__ seg_y, const double* __restrict__ y_in,
                                                                const double* __restrict__ D, double* __restrict__ z, double* __restrict__ p,
                                                                double* __restrict__ part_pq, const PcgDev* st, PeerWindow win, int parity,
                                                                unsigned long long seq) {
  if (st->active == 0) return;
  const bool peer = win.world > 1;                     // y = sum over the ranks' windows (k_cam_reduce9_warp published them)
  if (peer && !peer_wait(win, parity, seq)) { pcg_mark_fatal(const_cast<PcgDev*>(st)); return; }
  __shared__ double part[WPB][WPC][9];
  __shared__ double red[WPB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cl = warp / WPC, sub = warp % WPC;        // camera within the CTA, warp within the camera
  const int c = blockIdx.x * WPB + cl;
  const int k = lane % 9, j = lane / 9;               // 3 segment lanes x 9 components; lanes 27..31 idle
  if (c < L.n_cams && y_in == nullptr && !peer) {
    double acc = 0.0;
    if (lane < 27)
      acc = walk_segments<WPC>(L, seg_y, L.cam_seg_ptr[c] + sub * 3 + j, L.cam_seg_ptr[c + 1], k);
    const double a1 = __shfl_down_sync(0xffffffffu, acc, 9), a2 = __shfl_down_sync(0xffffffffu, acc, 18);
    acc = (acc + a1) + a2;
    if (lane < 9) part[cl][sub][lane] = acc;
  }
  __syncthreads();
  double pq = 0.0;
  if (c < L.n_cams && sub == 0 && lane < 9) {
    double acc;
    if (peer) acc = peer_gather(win, parity, c, (size_t)c * 9 + lane);
    else if (y_in == nullptr) {
      acc = part[cl][0][lane];
#pragma unroll
      for (int w = 1; w < WPC; ++w) acc += part[cl][w][lane];
    } else acc = y_in[(size_t)c * 9 + lane];
    const size_t e = (size_t)c * 9 + lane;
    const double zk = z[e];
    const double pk = (st->iter == 1) ? zk : __fma_rn(st->beta, p[e], zk);
    const double d = D[e];
    const double qk = (d * d) * pk + acc;
    p[e] = pk; z[e] = qk;
    pq = pk * qk;
  }
  if (sub == 0) {                                      // camera leaders: p.q of the camera into red[], then the CTA's 8 in order
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) pq += __shfl_down_sync(0xffffffffu, pq, o);   // lanes 0..8 hold data: 16-wide tree covers them
    if (lane == 0) red[cl] = pq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < WPB; ++w) s += red[w];
    part_pq[blockIdx.x] = s;
  }
}

// Warp per camera: x += alpha p; then either r -= alpha q (q lives in z) or, when `recompute`, nothing more
// (r is rebuilt by k_pcg_resid after the extra matvec).  Without recompute also: Q partials, z = M^-1 r, r.z.
__global__ void k_pcg_update(int n_cams, const double* __restrict__ Minv, const double* __restrict__ b, double* __restrict__ x,
                             const double* __restrict__ p, double* __restrict__ r, double* __restrict__ z, const double* __restrict__ part_pq,
                             int nparts, int recompute, double* __restrict__ part_Q, double* __restrict__ part_rho, PcgDev* st,
                             PcgParams prm) {
  if (st->active == 0) return;
  __shared__ double red[8];
  __shared__ double bc;
  const double pq = sum_fixed_all(part_pq, nparts, red, &bc);
  const bool ok = (pq > 0.0) && !isinf(pq);
  const double alpha = st->rho / pq;
  const bool go = ok && !isinf(alpha);               // otherwise Ceres breaks before touching x; k_pcg_head records why
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * WPB + warp;
  double qsum = 0.0, rz = 0.0;
  if (go && c < n_cams) {
    const size_t e = (size_t)c * 9 + (lane < 9 ? lane : 0);
    double rk = 0.0;
    if (lane < 9) {
      const double xk = x[e] + alpha * p[e];
      x[e] = xk;
      if (!recompute) {
        rk = r[e] - alpha * z[e];
        r[e] = rk;
        qsum = xk * (b[e] + rk);
      }
    }
    if (!recompute) {
      double zk = 0.0;
      if (Minv != nullptr) {
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const double rj = __shfl_sync(0xffffffffu, rk, j);
          if (lane < 9) zk += Minv[(size_t)c * 81 + lane * 9 + j] * rj;
        }
      } else zk = rk;
      if (lane < 9) { z[e] = zk; rz = rk * zk; }
    }
  }
  if (!recompute) {
    __shared__ double red2[2 * 8];
    double s12[2] = {qsum, rz};
    block_sum_fixed_n<2>(s12, red2);
    if (threadIdx.x == 0) { part_Q[blockIdx.x] = s12[0]; part_rho[blockIdx.x] = s12[1]; }
    // the CTA that publishes last finishes this iteration and opens the next one (what k_pcg_head would do next)
    if (pcg_last_block(st)) pcg_head_step(st, part_rho, part_pq, part_Q, nparts, prm, 0, red);
  }
}

// Residual reset: r = b - (y + D^2 x) with y = sum of segment partials of S_local x (or y_in); then the
// Q partials, z = M^-1 r and the r.z partials.
__global__ void k_pcg_resid2(BaDev L, const double* __restrict__ seg_y, const double* __restrict__ y_in, const double* __restrict__ D,
                             const double* __restrict__ Minv, const double* __restrict__ b, const double* __restrict__ x,
                             double* __restrict__ r, double* __restrict__ z, double* __restrict__ part_Q, double* __restrict__ part_rho,
                             PcgDev* st, const double* __restrict__ part_pq, PcgParams prm, PeerWindow win, int parity,
                             unsigned long long seq) {
  if (st->active == 0) return;
  const bool peer = win.world > 1;
  if (peer && !peer_wait(win, parity, seq)) { pcg_mark_fatal(st); return; }
  __shared__ double red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * WPB + warp;
  double qsum = 0.0, rz = 0.0;
  if (c < L.n_cams) {
    const int k = lane % 9, j = lane / 9;
    double acc = 0.0;
    if (peer) { if (lane < 9) acc = peer_gather(win, parity, c, (size_t)c * 9 + lane); }
    else if (y_in == nullptr) {
      if (lane < 27)
        for (int t = L.cam_seg_ptr[c] + j; t < L.cam_seg_ptr[c + 1]; t += 3) acc += seg_y[(size_t)t * 9 + k];
      const double a1 = __shfl_down_sync(0xffffffffu, acc, 9), a2 = __shfl_down_sync(0xffffffffu, acc, 18);
      acc = (acc + a1) + a2;
    } else if (lane < 9) acc = y_in[(size_t)c * 9 + lane];
    const size_t e = (size_t)c * 9 + (lane < 9 ? lane : 0);
    double rk = 0.0;
    if (lane < 9) {
      const double d = D[e], xk = x[e];
      rk = b[e] - ((d * d) * xk + acc);
      r[e] = rk;
      qsum = xk * (b[e] + rk);
    }
    double zk = 0.0;
    if (Minv != nullptr) {
#pragma unroll
      for (int jj = 0; jj < 9; ++jj) {
        const double rj = __shfl_sync(0xffffffffu, rk, jj);
        if (lane < 9) zk += Minv[(size_t)c * 81 + lane * 9 + jj] * rj;
      }
    } else zk = rk;
    if (lane < 9) { z[e] = zk; rz = rk * zk; }
  }
  const double s1 = block_sum_fixed(qsum, red), s2 = block_sum_fixed(rz, red);
  if (threadIdx.x == 0) { part_Q[blockIdx.x] = s1; part_rho[blockIdx.x] = s2; }
  if (pcg_last_block(st)) pcg_head_step(st, part_rho, part_pq, part_Q, (int)gridDim.x, prm, 0, red);
}

// y[c] = fixed-order sum of the camera's segment partials, in exactly the order k_pcg_reduce uses (WPC warps per camera,
// three interleaved segment lanes each, warp sums added in warp order).  Used before the allreduce of the multi-GPU path.
template <int WPC>
__global__ void __launch_bounds__(WPB * WPC * 32) k_cam_reduce9_warp(BaDev L, const double* __restrict__ seg_y, double* __restrict__ y,
                                                                      const int* guard, PeerWindow win, int parity, unsigned long long seq) {
  const bool run = guard == nullptr || *guard != 0;
  __shared__ double part[WPB][WPC][9];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cl = warp / WPC, sub = warp % WPC;
  const int c = blockIdx.x * WPB + cl;
  const int k = lane % 9, j = lane / 9;
  if (run && c < L.n_cams) {
    double acc = 0.0;
    if (lane < 27)
      acc = walk_segments<WPC>(L, seg_y, L.cam_seg_ptr[c] + sub * 3 + j, L.cam_seg_ptr[c + 1], k);
    const double a1 = __shfl_down_sync(0xffffffffu, acc, 9), a2 = __shfl_down_sync(0xffffffffu, acc, 18);
    acc = (acc + a1) + a2;
    if (lane < 9) part[cl][sub][lane] = acc;
  }
  __syncthreads();
  if (run && c < L.n_cams && sub == 0 && lane < 9) {
    double acc = part[cl][0][lane];
#pragma unroll
    for (int w = 1; w < WPC; ++w) acc += part[cl][w][lane];
    y[(size_t)c * 9 + lane] = acc;
  }
  if (win.world > 1) peer_publish(win, parity, seq);       // y is this rank's window slot
}

}  // namespace

int pcg_blocks(int n_cams) { return cdiv(n_cams, WPB); }

// Warps per camera of the segment walks (k_pcg_reduce, k_cam_reduce9_warp): each warp holds three partials per step.
static int pcg_wpc(const BaDev& L) {
  static const int forced = [] { const char* e = getenv("SKERES_PCG_WPC"); return e ? atoi(e) : 0; }();   // development / tests
  if (forced == 1 || forced == 2 || forced == 4) return forced;
  if (pcg_blocks(L.n_cams) <= 296) return 4;               // one wave of 1024-thread CTAs (2 per SM x 148 SMs): nothing to gain
  const double per_cam = (double)L.n_segs / (double)(L.n_cams > 0 ? L.n_cams : 1);
  return per_cam >= 96.0 ? 4 : (per_cam >= 36.0 ? 2 : 1);
}

void launch_cam_reduce9_warp(const BaDev& L, const double* seg_y, double* y, const int* guard, cudaStream_t s, const PeerWindow* win,
                             unsigned long long seq) {
  PeerWindow w{};
  if (win) w = *win;
  const int parity = (int)(seq & 1);
  if (w.world > 1) y = w.data[w.rank] + (size_t)parity * w.stride;
  switch (pcg_wpc(L)) {
    case 4: k_cam_reduce9_warp<4><<<pcg_blocks(L.n_cams), WPB * 4 * 32, 0, s>>>(L, seg_y, y, guard, w, parity, seq); break;
    case 2: k_cam_reduce9_warp<2><<<pcg_blocks(L.n_cams), WPB * 2 * 32, 0, s>>>(L, seg_y, y, guard, w, parity, seq); break;
    default: k_cam_reduce9_warp<1><<<pcg_blocks(L.n_cams), WPB * 32, 0, s>>>(L, seg_y, y, guard, w, parity, seq); break;
  }
  check_launch("k_cam_reduce9_warp");
}

void launch_pcg_begin(int n_cams, const double* rhs, const double* Minv, double* x, double* r, double* z, double* part_bb,
                      double* part_rho, PcgDev* st, int* lin_error, const double* global_lin_flag, cudaStream_t s) {
  const int nb = pcg_blocks(n_cams);
  k_pcg_begin<<<nb, WPB * 32, 0, s>>>(n_cams, rhs, Minv, x, r, z, part_bb, part_rho);
  k_pcg_start2<<<1, 256, 0, s>>>(st, part_bb, nb, lin_error, global_lin_flag);
  check_launch("k_pcg_begin");
}
void launch_pcg_head(PcgDev* st, const double* part_rho, const double* part_pq, const double* part_Q, int nparts, PcgParams prm,
                     int finish_only, cudaStream_t s) {
  k_pcg_head<<<1, 256, 0, s>>>(st, part_rho, part_pq, part_Q, nparts, prm, finish_only);
  check_launch("k_pcg_head");
}
void launch_pcg_reduce(const BaDev& L, const double* seg_y, const double* y_in, const double* D, double* z, double* p, double* part_pq,
                       const PcgDev* st, cudaStream_t s, const PeerWindow* win, unsigned long long seq) {
  PeerWindow w{};
  if (win) w = *win;
  const int parity = (int)(seq & 1);
  switch (pcg_wpc(L)) {
    case 4: k_pcg_reduce<4><<<pcg_blocks(L.n_cams), WPB * 4 * 32, 0, s>>>(L, seg_y, y_in, D, z, p, part_pq, st, w, parity, seq); break;
    case 2: k_pcg_reduce<2><<<pcg_blocks(L.n_cams), WPB * 2 * 32, 0, s>>>(L, seg_y, y_in, D, z, p, part_pq, st, w, parity, seq); break;
    default: k_pcg_reduce<1><<<pcg_blocks(L.n_cams), WPB * 32, 0, s>>>(L, seg_y, y_in, D, z, p, part_pq, st, w, parity, seq); break;
  }
  check_launch("k_pcg_reduce");
}
void launch_pcg_update(int n_cams, const double* Minv, const double* b, double* x, const double* p, double* r, double* z,
                       const double* part_pq, int recompute, double* part_Q, double* part_rho, PcgDev* st, PcgParams prm, cudaStream_t s) {
  const int nb = pcg_blocks(n_cams);
  k_pcg_update<<<nb, WPB * 32, 0, s>>>(n_cams, Minv, b, x, p, r, z, part_pq, nb, recompute, part_Q, part_rho, st, prm);
  check_launch("k_pcg_update");
}
void launch_pcg_resid2(const BaDev& L, const double* seg_y, const double* y_in, const double* D, const double* Minv, const double* b,
                       const double* x, double* r, double* z, double* part_Q, double* part_rho, PcgDev* st, const double* part_pq,
                       PcgParams prm, cudaStream_t s, const PeerWindow* win, unsigned long long seq) {
  PeerWindow w{};
  if (win) w = *win;
  k_pcg_resid2<<<pcg_blocks(L.n_cams), WPB * 32, 0, s>>>(L, seg_y, y_in, D, Minv, b, x, r, z, part_Q, part_rho, st, part_pq, prm, w,
                                                          (int)(seq & 1), seq);
  check_launch("k_pcg_resid2");
}

}  // namespace sk
