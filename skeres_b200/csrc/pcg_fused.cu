// pcg_fused.cu — ConjugateGradientsSolver::Solve on the implicit Schur complement as ONE persistent kernel.
//
// The kernel sequence of pcg_kernels.cu spends ~26 us of every 230 us PCG iteration (Venice-1778 shape, one B200) in two
// small vector kernels and the launch gaps around them, restarts the product's copy pipeline from cold once per iteration,
// and needs the host to poll the iteration state every few iterations.  Here the grid that runs the implicit-Schur product
// (one wave of co-resident CTAs: cooperative launch) stays resident for the whole linear solve:
//
//   loop   product of the direction z + beta p      every CTA walks its tiles (ba_product.cuh: ProductPass); the LAST tile of a
//                                                   pass already starts the TMA copies of the FIRST tile of the next pass -- the
//                                                   Jacobian does not change during a solve, only the input vector does
//          grid barrier
//          reduce   (virtual blocks of 8 cameras)   p = z + beta p, q = sum of the camera's segment partials + D^2 p, p.q slots
//          grid barrier
//          update                                   alpha = rho / p.q (every CTA sums the slots itself: same order, same bits);
//                                                   x += alpha p, r -= alpha q, z = M^-1 r, the x.(b + r) and r.z slots
//          [every residual_reset_period-th iteration: grid barrier, product of x, grid barrier, r = b - S x ...]
//          grid barrier
//          head step                                Q-based termination test, rho, beta, iteration counter -- taken by EVERY CTA
//                                                   on its own copy of the scalar state (pcg_head_step), so no broadcast is needed
//
// Three grid barriers per iteration, no launch, no host involvement: the host enqueues one kernel per linear solve and reads
// the outcome with the LM iteration's state block.  The phases are the device functions the kernel sequence is built from
// (pcg_device.cuh), hence the two forms give bit-identical solves (tests: test_fused_pcg_matches_kernel_sequence_bitwise).
//
// Multi-GPU (points partitioned, cameras replicated): the per-iteration exchange of the camera-sized product runs inside the
// kernel over the NVLink peer window (comm.cuh) -- own contribution into the window, grid barrier, flags to every rank, wait for
// every rank's flag, gather in rank order.  A wait that times out does not leave the loop on its own (the other CTAs would wait
// at the next grid barrier forever): it raises the window's error word, which all CTAs read after the next barrier.
#include "pcg_fused.cuh"

#include <algorithm>
#include <cstdlib>

#include "ba_product.cuh"
#include "pcg_device.cuh"

namespace sk {

namespace {

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Barrier over all CTAs of the (co-resident) grid.  `*target` (shared memory, thread 0's) counts the arrivals expected so far;
// the counter only grows.
__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int* target) {
  __syncthreads();                                   // the CTA's writes happen before thread 0's fence (cumulativity)
  if (threadIdx.x == 0) {
    const unsigned int t = *target + gridDim.x;
    *target = t;
    __threadfence();
    atomicAdd(bar, 1u);                                // (arriving with red.release.gpu instead of fence + atomic: measured, no difference)
    while (ld_acquire_gpu(bar) < t) { }
  }
  __syncthreads();
}

// Development only (never defined in the product build): timing ablations of k_pcg_solve.
#ifndef SK_FUSED_COHERENT
#define SK_FUSED_COHERENT 1      // 0: input gathers through L1 (WRONG results: stale lines)
#endif
#ifndef SK_FUSED_CHAIN
#define SK_FUSED_CHAIN 1         // 0: every pass starts its copy pipeline from cold
#endif

// Everything the loop keeps between its phases lives in shared memory, not in registers: the product pass needs every register
// of the 128 that two CTAs per SM allow (see ProductPass).
struct SolveCtl {
  PcgDev s;                                          // this CTA's copy of the scalar state: identical in every CTA at every barrier
  unsigned long long seq;                            // peer-window exchanges so far
  unsigned long long t_last, ns_product, ns_vector;  // CTA 0, thread 0: device-clock accounting of the phases
  unsigned int bar_target;
  int peer_failed;
  int reset_pass;                                    // the coming pass multiplies x (residual reset), not the direction
  double red[8];
  double bc;
};

// ONE call site of the product pass.  With two (direction product / residual-reset product of x) ptxas ran out of registers at
// the 128 that two CTAs per SM allow and issued every shared-memory load of the per-point / per-segment sums right in front of
// the add that consumes it instead of a batch of loads ahead (SASS: "LDS DADD LDS DADD ..." against "LDS x 11, then LDS / DADD
// interleaved" in the stand-alone product kernel): 0.316 ms per product instead of 0.206 ms (profiles/r02_v2_*).  The loop
// below therefore runs one pass per trip and selects its operands from the state: a residual reset is a trip of its own.
template <bool TMAP>
__global__ void __launch_bounds__(T, 2) k_pcg_solve(const __grid_constant__ CUtensorMap tmapJ, const PcgSolveArgs A) {
  extern __shared__ __align__(128) double sm[];
  __shared__ SolveCtl ctl;
  const int tid = threadIdx.x;
  if (tid == 0) {
    ctl.s = *A.st; ctl.seq = A.seq_base; ctl.bar_target = 0; ctl.ns_product = 0; ctl.ns_vector = 0; ctl.t_last = global_ns();
    ctl.reset_pass = 0;
  }
  __syncthreads();
  if (ctl.s.active == 0) return;                     // the set-up failed or b == 0 (k_pcg_start2): the same decision in every CTA
  const BaDev& L = A.L;
  const int nparts = (L.n_cams + WPB - 1) / WPB;
  const bool peer = A.win.world > 1;
  const bool packed = nparts > (int)gridDim.x;       // more virtual blocks than CTAs: three blocks per round by lane groups (pcg_device.cuh)
  const bool stamps = A.phase_ns != nullptr && blockIdx.x == 0 && tid == 0;
  ProductPass<TMAP, false, SK_FUSED_COHERENT != 0> P;
  P.init(L, sm);

  pcg_head_step(&ctl.s, A.part_rho, A.part_pq, A.part_Q, nparts, A.prm, 0);   // opens iteration 1 (what k_pcg_head does in the sequence)
  __syncthreads();
  while (ctl.s.active) {
    // ---- one implicit-Schur product: of the direction z + beta p (p = z in iteration 1), or -- residual reset -- of x
    const bool reset_pass = ctl.reset_pass != 0;
    if (stamps) { const unsigned long long t = global_ns(); ctl.ns_vector += t - ctl.t_last; ctl.t_last = t; }
    P.run(&tmapJ, L, A.J2, reset_pass ? A.x : A.z, A.p, ctl.s.beta, !reset_pass && ctl.s.iter != 1, A.einv, A.seg_y, SK_FUSED_CHAIN != 0, sm);
    grid_sync(A.grid_bar, &ctl.bar_target);
    int parity = 0;
    if (peer) {                                      // exchange of the camera-sized result over the NVLink peer window
      if (tid == 0) ++ctl.seq;
      __syncthreads();
      const unsigned long long seq = ctl.seq;
      parity = (int)(seq & 1ull);
      double* y = A.win.data[A.win.rank] + (size_t)parity * A.win.stride;
      // only the virtual blocks this rank holds observations for (the peers read a camera's slot from the ranks in its mask only);
      // the flags of all rounds are fetched together
      unsigned own = 0xffffffffu;
      if (A.win.vb_own != nullptr && nparts <= 32 * (int)gridDim.x) {
        own = 0u;
        for (int vb = blockIdx.x, k = 0; vb < nparts; vb += gridDim.x, ++k) own |= (A.win.vb_own[vb] != 0 ? 1u : 0u) << k;
      }
      for (int vb = blockIdx.x, k = 0; vb < nparts; vb += gridDim.x, ++k)
        if ((own >> (k & 31)) & 1u) cam_reduce9_block<1>(L, vb, A.seg_y, y);
      grid_sync(A.grid_bar, &ctl.bar_target);        // this rank's contribution is complete
      if (blockIdx.x == 0) peer_store_flags(A.win, parity, seq);
      peer_wait(A.win, parity, seq);                 // a time-out raises win.error; checked by all CTAs after the next barrier
    }
    if (stamps) { const unsigned long long t = global_ns(); ctl.ns_product += t - ctl.t_last; ctl.t_last = t; }
    // ---- vector phases
    bool finish = true;                              // this trip ends the iteration (head step)
    if (!reset_pass) {
      const int it = ctl.s.iter; const double beta = ctl.s.beta;
      if (packed && peer)
        for (int vb = blockIdx.x; vb < nparts; vb += 3 * gridDim.x)
          pcg_reduce_block3_peer(L, vb, (int)gridDim.x, nparts, A.D, A.z, A.p, A.part_pq, it, beta, A.win, parity);
      else
        for (int vb = blockIdx.x; vb < nparts; vb += gridDim.x)
          pcg_reduce_block<1>(L, vb, A.seg_y, nullptr, A.D, A.z, A.p, A.part_pq, it, beta, A.win, parity);
      grid_sync(A.grid_bar, &ctl.bar_target);
      const int recompute = (it % A.reset_period == 0) ? 1 : 0;
      const double pq = sum_fixed_all(A.part_pq, nparts, ctl.red, &ctl.bc);
      const bool ok = (pq > 0.0) && !isinf(pq);
      const double alpha = ctl.s.rho / pq;
      const bool go = ok && !isinf(alpha);
      if (packed)
        for (int vb = blockIdx.x; vb < nparts; vb += 3 * gridDim.x)
          pcg_update_block3(L.n_cams, vb, (int)gridDim.x, nparts, A.Minv, A.b, A.x, A.p, A.r, A.z, alpha, go, recompute, A.part_Q, A.part_rho);
      else
        for (int vb = blockIdx.x; vb < nparts; vb += gridDim.x)
          pcg_update_block(L.n_cams, vb, A.Minv, A.b, A.x, A.p, A.r, A.z, alpha, go, recompute, A.part_Q, A.part_rho);
      if (recompute) finish = false;                 // r = b - S x from one more product: the next trip
    } else {
      for (int vb = blockIdx.x; vb < nparts; vb += gridDim.x)
        pcg_resid_block(L, vb, A.seg_y, nullptr, A.D, A.Minv, A.b, A.x, A.r, A.z, A.part_Q, A.part_rho, A.win, parity);
    }
    grid_sync(A.grid_bar, &ctl.bar_target);
    if (peer) {                                      // did any CTA's wait time out?  (same answer in every CTA: read after the barrier)
      if (tid == 0) ctl.peer_failed = *(volatile int*)A.win.error;
      __syncthreads();
      if (ctl.peer_failed != 0) {
        __syncthreads();
        if (tid == 0) { ctl.s.active = 0; ctl.s.termination = LIN_FATAL; }
        __syncthreads();
        break;
      }
    }
    if (tid == 0) ctl.reset_pass = finish ? 0 : 1;
    if (finish) pcg_head_step(&ctl.s, A.part_rho, A.part_pq, A.part_Q, nparts, A.prm, 0);   // finishes this iteration, opens the next
    __syncthreads();
  }
  P.drain(L, sm);
  if (blockIdx.x == 0 && tid == 0) {
    ctl.s.done_count = 0;
    *A.st = ctl.s;
    if (A.phase_ns != nullptr) {
      ctl.ns_vector += global_ns() - ctl.t_last;
      atomicAdd(A.phase_ns, ctl.ns_product); atomicAdd(A.phase_ns + 1, ctl.ns_vector);
    }
  }
}

}  // namespace

namespace {
using SolveKernel = void (*)(const CUtensorMap, const PcgSolveArgs);
struct SolveCfg { SolveKernel kernel; size_t smem; int grid; };
// Kernel variant, shared memory and grid (one wave of co-resident CTAs) for a problem; grid == 0: not supported.
SolveCfg solve_config(const BaDev& L, bool have_tmap) {
  static int sms = 0, smem_max = 0, coop = 0;
  if (sms == 0) {
    int dev = 0; SK_CUDA(cudaGetDevice(&dev));
    SK_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    SK_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    SK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int v = have_tmap ? 1 : 0;
  static const SolveKernel kernels[2] = {k_pcg_solve<false>, k_pcg_solve<true>};
  SolveCfg c{kernels[v], ProductPass<true, false, true>::smem_bytes(L), 0};
  if (L.n_tiles == 0 || L.tile_rec == nullptr || L.n_giant != 0 || !coop || c.smem > (size_t)smem_max) return c;
  static size_t cfg_smem[2] = {0}; static int per_sm[2] = {0};
  if (c.smem != cfg_smem[v]) {
    if (c.smem > 48 * 1024) SK_CUDA(cudaFuncSetAttribute(c.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    SK_CUDA(cudaFuncSetAttribute(c.kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    SK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[v], c.kernel, T, c.smem));
    cfg_smem[v] = c.smem;
  }
  if (per_sm[v] > 0) c.grid = std::max(1, std::min(L.n_tiles, per_sm[v] * sms));   // every CTA resident: the grid barriers need it
  return c;
}
}  // namespace

bool pcg_solve_supported(const BaDev& L, bool have_tmap) { return solve_config(L, have_tmap).grid > 0; }

void launch_pcg_solve(const PcgSolveArgs& args, const CUtensorMap* tmapJ, cudaStream_t s) {
  const SolveCfg c = solve_config(args.L, tmapJ != nullptr);
  SK_REQUIRE(c.grid > 0, SK_ERR_INTERNAL, "fused PCG solve launched on a problem it does not support");
  static const CUtensorMap no_map{};
  CUtensorMap map = tmapJ != nullptr ? *tmapJ : no_map;
  PcgSolveArgs a = args;
  void* params[2] = {&map, &a};
  SK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(c.kernel), dim3(c.grid), dim3(T), params, c.smem, s));
}

}  // namespace sk
