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

#include "pcg_device.cuh"

namespace sk {

namespace {

__device__ __forceinline__ void pcg_mark_fatal(PcgDev* st) {
  if (blockIdx.x == 0 && threadIdx.x == 0) { st->active = 0; st->termination = LIN_FATAL; }
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
  if (last) peer_store_flags(win, parity, seq);
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
    const double zk = precondition(Minv, c, lane, rk);
    if (lane < 9) { x[(size_t)c * 9 + lane] = 0.0; r[(size_t)c * 9 + lane] = rk; z[(size_t)c * 9 + lane] = zk; bb = __dmul_rn(rk, rk); rz = __dmul_rn(rk, zk); }
  }
  const double s1 = block_sum_fixed(bb, red), s2 = block_sum_fixed(rz, red);
  if (threadIdx.x == 0) { part_bb[blockIdx.x] = s1; part_rho[blockIdx.x] = s2; }
}

__global__ void k_pcg_start2(PcgDev* st, const double* part_bb, int nparts, int* lin_error, const double* global_lin_flag,
                             unsigned int* grid_bar) {
  __shared__ double red[8];
  const double bb = sum_fixed(part_bb, nparts, red);
  if (threadIdx.x != 0) return;
  st->norm_b = sqrt(bb);
  st->rho = 1.0; st->last_rho = 1.0; st->pq = 0.0; st->alpha = 0.0; st->beta = 0.0;
  st->Q0 = 0.0; st->Q1 = 0.0;                                   // Q0 = -x.(b + r) with x = 0
  st->iter = 0; st->active = 1; st->termination = LIN_NO_CONVERGENCE; st->pad_ = 0;   // pad_ = last finished iteration
  st->done_count = 0; st->pad2_ = 0;
  if (grid_bar != nullptr) *grid_bar = 0u;                      // grid barrier of the fused solve (pcg_fused.cu)
  if (global_lin_flag != nullptr && *global_lin_flag != 0.0) *lin_error |= 1;     // failed on some rank = failed for all
  if (*lin_error) { st->active = 0; st->termination = LIN_FAILURE; }
  else if (st->norm_b == 0.0) { st->active = 0; st->termination = LIN_SUCCESS; }
}

__global__ void k_pcg_head(PcgDev* st, const double* part_rho, const double* part_pq, const double* part_Q, int nparts,
                           PcgParams prm, int finish_only) {
  pcg_head_step(st, part_rho, part_pq, part_Q, nparts, prm, finish_only);
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

// One CTA per virtual block (pcg_device.cuh); WPC warps per camera.
template <int WPC>
__global__ void __launch_bounds__(WPB * WPC * 32) k_pcg_reduce(BaDev L, const double* __restrict__ seg_y, const double* __restrict__ y_in,
                                                                const double* __restrict__ D, double* z, double* p,
                                                                double* __restrict__ part_pq, const PcgDev* st, PeerWindow win, int parity,
                                                                unsigned long long seq) {
  if (st->active == 0) return;
  if (win.world > 1 && !peer_wait(win, parity, seq)) { pcg_mark_fatal(const_cast<PcgDev*>(st)); return; }   // y = sum over the ranks' windows
  pcg_reduce_block<WPC>(L, blockIdx.x, seg_y, y_in, D, z, p, part_pq, st->iter, st->beta, win, parity);
}

__global__ void k_pcg_update(int n_cams, const double* __restrict__ Minv, const double* __restrict__ b, double* x,
                             const double* p, double* r, double* z, const double* __restrict__ part_pq,
                             int nparts, int recompute, double* __restrict__ part_Q, double* __restrict__ part_rho, PcgDev* st,
                             PcgParams prm) {
  if (st->active == 0) return;
  __shared__ double red[8];
  __shared__ double bc;
  const double pq = sum_fixed_all(part_pq, nparts, red, &bc);
  const bool ok = (pq > 0.0) && !isinf(pq);
  const double alpha = st->rho / pq;
  const bool go = ok && !isinf(alpha);               // otherwise Ceres breaks before touching x; the head step records why
  pcg_update_block(n_cams, blockIdx.x, Minv, b, x, p, r, z, alpha, go, recompute, part_Q, part_rho);
  // the CTA that publishes last finishes this iteration and opens the next one (what k_pcg_head would do next)
  if (!recompute && pcg_last_block(st)) pcg_head_step(st, part_rho, part_pq, part_Q, nparts, prm, 0);
}

__global__ void k_pcg_resid2(BaDev L, const double* __restrict__ seg_y, const double* __restrict__ y_in, const double* __restrict__ D,
                             const double* __restrict__ Minv, const double* __restrict__ b, const double* x,
                             double* r, double* z, double* __restrict__ part_Q, double* __restrict__ part_rho,
                             PcgDev* st, const double* __restrict__ part_pq, PcgParams prm, PeerWindow win, int parity,
                             unsigned long long seq) {
  if (st->active == 0) return;
  if (win.world > 1 && !peer_wait(win, parity, seq)) { pcg_mark_fatal(st); return; }
  pcg_resid_block(L, blockIdx.x, seg_y, y_in, D, Minv, b, x, r, z, part_Q, part_rho, win, parity);
  if (pcg_last_block(st)) pcg_head_step(st, part_rho, part_pq, part_Q, (int)gridDim.x, prm, 0);
}

// y[c] = fixed-order sum of the camera's segment partials, in exactly the order k_pcg_reduce uses.  Used before the allreduce
// of the multi-GPU path: into ybuf (NCCL) or into this rank's peer-window slot, published by the last CTA.
template <int WPC>
__global__ void __launch_bounds__(WPB * WPC * 32) k_cam_reduce9_warp(BaDev L, const double* __restrict__ seg_y, double* __restrict__ y,
                                                                      const int* guard, PeerWindow win, int parity, unsigned long long seq) {
  const bool run = guard == nullptr || *guard != 0;
  if (run) cam_reduce9_block<WPC>(L, blockIdx.x, seg_y, y);
  if (win.world > 1) peer_publish(win, parity, seq);       // y is this rank's window slot
}

}  // namespace

int pcg_blocks(int n_cams) { return cdiv(n_cams, WPB); }

// Warps per camera of the segment walks (k_pcg_reduce, k_cam_reduce9_warp): each warp holds three partials per step.
static int pcg_wpc(const BaDev& L) {
  const char* env = getenv("SKERES_PCG_WPC");              // development / tests
  const int forced = env ? atoi(env) : 0;
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
                      double* part_rho, PcgDev* st, int* lin_error, const double* global_lin_flag, unsigned int* grid_bar, cudaStream_t s) {
  const int nb = pcg_blocks(n_cams);
  k_pcg_begin<<<nb, WPB * 32, 0, s>>>(n_cams, rhs, Minv, x, r, z, part_bb, part_rho);
  k_pcg_start2<<<1, 256, 0, s>>>(st, part_bb, nb, lin_error, global_lin_flag, grid_bar);
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
