// ba_kernels.cu — sm_100a tile kernels of the bundle-adjustment hot path.
//
// One CTA (256 threads) per tile of <= 256 observations holding whole points (ba_layout.h).
// Thread i of the CTA owns observation i of the tile: it reads its 12 Jacobian planes as coalesced
// 16-byte vectors, keeps them in registers and reuses them for every phase of the kernel, so each
// kernel reads the stored Jacobian exactly once (algorithmic bytes == DRAM bytes).
// Per-point sums are formed by one thread per point in observation order (the order Ceres'
// SchurEliminator walks a chunk); per-camera sums go through the tile-local camera segments
// (seg_reduce9) and a second-level fixed-order reduction (k_cam_reduce): no atomics anywhere.
//
// What each kernel restates (Ceres sources are NOT in the reference tree):
//   k_ba_evaluate         ProgramEvaluator / ResidualBlock::Evaluate / Corrector  (A.2) on top of
//                         AutodiffCostFunction.scala:74-134 + SimpleBundleAdjuster.scala:79-119
//   k_ba_schur_setup      SchurEliminator::Eliminate  (A.5) restricted to rhs + diagonal blocks
//                         (= SchurJacobiPreconditioner::UpdateImpl, A.7) + ImplicitSchurComplement::Init
//   k_ba_matvec           ImplicitSchurComplement::RightMultiply (A.6), the four passes fused
//   k_ba_back_substitute  ImplicitSchurComplement::BackSubstitute / SchurEliminator::BackSubstitute
//                         + the model-cost-change product of TrustRegionMinimizer (A.3 step 3)
#include "ba_kernels.cuh"

#include <algorithm>
#include <cstdlib>
#include <utility>

#include "ba_evaluate.cuh"
#include "ba_product.cuh"
#include "lm_kernels.cuh"
#include "user_functor.cuh"

namespace sk {

namespace {

// Sums of N per-thread values over the CTA, returned to EVERY thread (fixed shuffle tree, then the 8 warp totals in
// warp order): used by the long-track kernels, which accumulate per-point sums chunk by chunk.  red: [N][8] doubles.
template <int N>
__device__ __forceinline__ void block_sum_all(double (&x)[N], double* red) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int n = 0; n < N; ++n) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x[n] += __shfl_down_sync(0xffffffffu, x[n], o);
    if (l == 0) red[n * (T / 32) + w] = x[n];
  }
  __syncthreads();
#pragma unroll
  for (int n = 0; n < N; ++n) {
    double r = 0.0;
#pragma unroll
    for (int k = 0; k < T / 32; ++k) r += red[n * (T / 32) + k];
    x[n] = r;
  }
  __syncthreads();   // red may be rewritten by the next call
}

// out[(sb+s)*ostride + ooff + k] = sum over segment s of v[k][local obs], fixed (point) order.
// pos_map != nullptr: row pos_map[segment] instead of row `segment` (the camera-major rows of the implicit-Schur product).
__device__ __forceinline__ void seg_reduce9(const BaDev& L, const Tile& q, const double* v, double* out,
                                            int ostride, int ooff, const int* pos_map = nullptr) {
  for (int idx = threadIdx.x; idx < q.ns * 9; idx += blockDim.x) {
    const int s = idx / 9, k = idx - s * 9;
    const int b = L.seg_ptr[q.sb + s], e = L.seg_ptr[q.sb + s + 1];
    double sum = 0.0;
    for (int pos = b; pos < e; ++pos) sum += v[k * VLD + L.seg_perm[pos]];
    const int row = pos_map != nullptr ? pos_map[q.sb + s] : q.sb + s;
    out[(size_t)row * ostride + ooff + k] = sum;
  }
}

__device__ __forceinline__ double2 ldcs2(const double2* p) { return __ldcs(p); }

// ------------------------------------------------------------------------------------------------
// Residuals (+ Jacobian) of one tile: ba_evaluate.cuh (the body is shared with the kernels NVRTC compiles for functors given
// as source); here with the built-in SnavelyReprojectionError (jet.cuh: two 6-wide dual stages).
// (Register caps for 3 CTAs per SM -- here and in k_ba_schur_setup -- were measured and buy nothing: profiles/r02_v11_setup_kernels.md.)
template <bool JAC>
__global__ void __launch_bounds__(T) k_ba_evaluate(BaDev L, const double* __restrict__ x, const double* __restrict__ scale,
                                                   LossSpec loss, int write_j, double2* __restrict__ J2,
                                                   double2* __restrict__ r2, double* __restrict__ grad,
                                                   double* __restrict__ cnorm2, double* __restrict__ seg_g,
                                                   double* __restrict__ seg_n, double* __restrict__ tile_cost,
                                                   double* __restrict__ chunk_pt, int* fail_flag, const int* guard) {
  ba_evaluate_tile<JAC, SnavelyBuiltin>(L, x, scale, loss, write_j, J2, r2, grad, cnorm2, seg_g, seg_n, tile_cost, chunk_pt, fail_flag, guard);
}

// ------------------------------------------------------------------------------------------------
// Second level of the per-camera sums: out[c][k] = sum over the camera's (tile, camera) partials seg[s][k], tile order.
// One CTA per camera.  The partial list (~190 entries on the Venice shape) is dealt round-robin to R = 288 / K sub-ranges
// (32 for K = 9, 6 for K = 45), each summed by K threads with four loads in flight; the R sub-sums are then added in
// sub-range order.  Fixed order, no atomics.  (One thread per (camera, k) walking the whole list cost 37-41 us per launch.)
constexpr int kCamReduceThreads = 288;
__global__ void __launch_bounds__(kCamReduceThreads) k_cam_reduce(BaDev L, int K, const double* __restrict__ seg, double* __restrict__ out,
                                                                 const int* guard) {
  if (guard != nullptr && *guard == 0) return;
  __shared__ double part[kCamReduceThreads];
  const int c = blockIdx.x;
  const int R = kCamReduceThreads / K;
  const int r = threadIdx.x / K, k = threadIdx.x - r * K;
  const int e = L.cam_seg_ptr[c + 1];
  if (r < R) {
    double sum = 0.0;
    int t = L.cam_seg_ptr[c] + r;
    for (; t + 3 * R < e; t += 4 * R) {
      const int s0 = L.cam_seg[t], s1 = L.cam_seg[t + R], s2 = L.cam_seg[t + 2 * R], s3 = L.cam_seg[t + 3 * R];
      const double x0 = seg[(size_t)s0 * K + k], x1 = seg[(size_t)s1 * K + k], x2 = seg[(size_t)s2 * K + k], x3 = seg[(size_t)s3 * K + k];
      sum += x0; sum += x1; sum += x2; sum += x3;
    }
    for (; t < e; t += R) sum += seg[(size_t)L.cam_seg[t] * K + k];
    part[threadIdx.x] = sum;                       // == part[r * K + k]
  }
  __syncthreads();
  if ((int)threadIdx.x < K) {
    double a = part[threadIdx.x];
    for (int q = 1; q < R; ++q) a += part[q * K + threadIdx.x];
    out[(size_t)c * K + threadIdx.x] = a;
  }
}

// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int upper_row(int idx) {   // row k of the idx-th entry of a packed 9x9 upper triangle
  int k = 0, start = 0;
  while (idx >= start + (9 - k)) { start += 9 - k; ++k; }
  return k;
}
__host__ __device__ constexpr int upper_col(int idx) {
  int k = 0, start = 0;
  while (idx >= start + (9 - k)) { start += 9 - k; ++k; }
  return k + (idx - start);
}

// Entry IDX (0..44) of the packed upper 9x9 block one observation contributes: F^T F - G^T (E^T E)^-1 G, indices folded at
// compile time so that F, G, H stay in registers.
// Entry (k, l) of an observation's contribution to its camera's diagonal block of the reduced matrix,
//   F^T F - G^T (E^T E)^-1 G  with  G = E^T F      (SchurEliminator)
// written as F^T P F with the 2 x 2 projector P = I - E (E^T E)^-1 E^T of the observation: 18 values (Pf = P F) instead of the 54
// of G and H = (E^T E)^-1 G stay alive across the three staging rounds -- the kernel spilled 196 bytes at the 128 registers that
// two CTAs per SM allow -- and an entry costs two multiply-adds instead of five.  JACOBI (ftf_only): P = I.
template <int IDX>
__device__ __forceinline__ double block_entry(const double2 (&Fv)[9], const double2 (&Pf)[9]) {
  constexpr int k = upper_row(IDX), l = upper_col(IDX);
  return Fv[k].x * Pf[l].x + Fv[k].y * Pf[l].y;
}
template <int FIRST, int... Is>
__device__ __forceinline__ void stage_block_entries(double* vt, const double2 (&Fv)[9], const double2 (&Pf)[9], std::integer_sequence<int, Is...>) {
  ((vt[Is * VLD] = block_entry<FIRST + Is>(Fv, Pf)), ...);
}

constexpr int kSetupPlanes = 18;     // staged planes per round of k_ba_schur_setup: 9 rhs + 45 block entries = 3 rounds of 18
__global__ void __launch_bounds__(T, 2) k_ba_schur_setup(BaDev L, const double2* __restrict__ J2, const double2* __restrict__ r2,
                                                      const double* __restrict__ D, double* __restrict__ einv,
                                                      double* __restrict__ seg_rhs, double* __restrict__ seg_M,
                                                      int* error_flag, int ftf_only) {
  // ftf_only: the diagonal blocks keep only F^T F (JACOBI preconditioner = block_diagonal_FtF_inverse of the implicit
  // Schur complement) instead of F^T F - G^T (E^T E)^-1 G (SCHUR_JACOBI / the explicit reduced matrix).
  // The 54 per-observation values that go to a camera (9 of the reduced right-hand side, 45 of the packed upper block)
  // are staged 18 planes at a time, so that every thread of the CTA has a segment sum to form; the tile's segment
  // structure is fetched into shared memory once, behind the Jacobian loads.
  extern __shared__ double sm[];
  const Tile q = load_tile(L, blockIdx.x);
  if (q.chunk >= 0) return;            // long tracks: k_ba_schur_setup_giant
  const int tid = threadIdx.x;
  double* v = sm;                      // [18][VLD]
  double* pinv = v + kSetupPlanes * VLD;   // [max_pt][9]: inverse (6, upper) + inverse * g (3)
  TileMetaSmem meta;
  meta.sptr = reinterpret_cast<int*>(pinv + L.max_pt_tile * 9);      // [max_seg + 1]
  meta.pptr = meta.sptr + L.max_seg_tile + 1;                        // [max_pt + 1]
  meta.sperm = reinterpret_cast<unsigned short*>(meta.pptr + L.max_pt_tile + 1);   // [T]
  const bool active = tid < q.no;
  const int i = q.ob + tid;
  const size_t O = (size_t)L.n_obs;
  double2 Fv[9], Ev[3], r = make_double2(0.0, 0.0);
  int ptl = 0;
  if (active) {
#pragma unroll
    for (int k = 0; k < 9; ++k) Fv[k] = ldcs2(J2 + k * O + i);
#pragma unroll
    for (int k = 0; k < 3; ++k) Ev[k] = ldcs2(J2 + (9 + k) * O + i);
    r = r2[i];
    ptl = L.obs_ptl[i];
  }
  stage_tile_meta(L, q, meta);
  if (active) {
    // E^T E (upper: 00 01 02 11 12 22) and E^T r
    v[0 * VLD + tid] = Ev[0].x * Ev[0].x + Ev[0].y * Ev[0].y;
    v[1 * VLD + tid] = Ev[0].x * Ev[1].x + Ev[0].y * Ev[1].y;
    v[2 * VLD + tid] = Ev[0].x * Ev[2].x + Ev[0].y * Ev[2].y;
    v[3 * VLD + tid] = Ev[1].x * Ev[1].x + Ev[1].y * Ev[1].y;
    v[4 * VLD + tid] = Ev[1].x * Ev[2].x + Ev[1].y * Ev[2].y;
    v[5 * VLD + tid] = Ev[2].x * Ev[2].x + Ev[2].y * Ev[2].y;
#pragma unroll
    for (int k = 0; k < 3; ++k) v[(6 + k) * VLD + tid] = Ev[k].x * r.x + Ev[k].y * r.y;
  }
  __syncthreads();
  if (tid < q.np) {
    const int p = q.pb + tid;
    const int b = meta.pptr[tid], e = meta.pptr[tid + 1];
    const double* Dp = D + (size_t)9 * L.n_cams + (size_t)p * 3;
    double m[6] = {Dp[0] * Dp[0], 0.0, 0.0, Dp[1] * Dp[1], 0.0, Dp[2] * Dp[2]};
    double g[3] = {0.0, 0.0, 0.0};
    for (int j = b; j < e; ++j) {
#pragma unroll
      for (int k = 0; k < 6; ++k) m[k] += v[k * VLD + j];
#pragma unroll
      for (int k = 0; k < 3; ++k) g[k] += v[(6 + k) * VLD + j];
    }
    double inv[6];
    if (!invert_spd3(m, inv)) {
      atomicOr(error_flag, 1);
#pragma unroll
      for (int k = 0; k < 6; ++k) inv[k] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) { einv[(size_t)p * 6 + k] = inv[k]; pinv[tid * 9 + k] = inv[k]; }
    pinv[tid * 9 + 6] = inv[0] * g[0] + inv[1] * g[1] + inv[2] * g[2];
    pinv[tid * 9 + 7] = inv[1] * g[0] + inv[3] * g[1] + inv[4] * g[2];
    pinv[tid * 9 + 8] = inv[2] * g[0] + inv[4] * g[1] + inv[5] * g[2];
  }
  __syncthreads();
  double2 Pf[9];
  if (active) {
    const double* pi = pinv + ptl * 9;
    // reduced rhs: F^T (r - E (E^T E)^-1 E^T r)
    const double q0 = r.x - (Ev[0].x * pi[6] + Ev[1].x * pi[7] + Ev[2].x * pi[8]);
    const double q1 = r.y - (Ev[0].y * pi[6] + Ev[1].y * pi[7] + Ev[2].y * pi[8]);
#pragma unroll
    for (int k = 0; k < 9; ++k) v[k * VLD + tid] = Fv[k].x * q0 + Fv[k].y * q1;
    // P = I - E (E^T E)^-1 E^T (2 x 2, symmetric): T_a = sum_b inv[a][b] e_b, M = sum_a e_a T_a^T
    double p00 = 1.0, p01 = 0.0, p11 = 1.0;
    if (!ftf_only) {
      const double t0x = pi[0] * Ev[0].x + pi[1] * Ev[1].x + pi[2] * Ev[2].x, t0y = pi[0] * Ev[0].y + pi[1] * Ev[1].y + pi[2] * Ev[2].y;
      const double t1x = pi[1] * Ev[0].x + pi[3] * Ev[1].x + pi[4] * Ev[2].x, t1y = pi[1] * Ev[0].y + pi[3] * Ev[1].y + pi[4] * Ev[2].y;
      const double t2x = pi[2] * Ev[0].x + pi[4] * Ev[1].x + pi[5] * Ev[2].x, t2y = pi[2] * Ev[0].y + pi[4] * Ev[1].y + pi[5] * Ev[2].y;
      p00 = 1.0 - (Ev[0].x * t0x + Ev[1].x * t1x + Ev[2].x * t2x);
      p01 = -(Ev[0].x * t0y + Ev[1].x * t1y + Ev[2].x * t2y);
      p11 = 1.0 - (Ev[0].y * t0y + Ev[1].y * t1y + Ev[2].y * t2y);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) Pf[k] = make_double2(p00 * Fv[k].x + p01 * Fv[k].y, p01 * Fv[k].x + p11 * Fv[k].y);
    // round 0: planes 0..8 = reduced rhs (above), planes 9..17 = block entries 0..8
    stage_block_entries<0>(v + 9 * VLD + tid, Fv, Pf, std::make_integer_sequence<int, 9>());
  }
  __syncthreads();
  seg_reduce_planes<kSetupPlanes>(q, meta, v, 9, seg_rhs, 9, seg_M, 45, 0);
  // rounds 1, 2: block entries 9..26 and 27..44
  __syncthreads();
  if (active) stage_block_entries<9>(v + tid, Fv, Pf, std::make_integer_sequence<int, kSetupPlanes>());
  __syncthreads();
  seg_reduce_planes<kSetupPlanes>(q, meta, v, 0, seg_M, 45, seg_M, 45, 9);
  __syncthreads();
  if (active) stage_block_entries<9 + kSetupPlanes>(v + tid, Fv, Pf, std::make_integer_sequence<int, kSetupPlanes>());
  __syncthreads();
  seg_reduce_planes<kSetupPlanes>(q, meta, v, 0, seg_M, 45, seg_M, 45, 9 + kSetupPlanes);
}

__global__ void k_ba_precond_invert(BaDev L, const double* __restrict__ M45, const double* __restrict__ D,
                                    double* __restrict__ Minv, int* error_flag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= L.n_cams) return;
  double A[81];
  int idx = 0;
  for (int k = 0; k < 9; ++k)
    for (int l = k; l < 9; ++l) { const double m = M45[(size_t)c * 45 + idx++]; A[k * 9 + l] = m; A[l * 9 + k] = m; }
  for (int k = 0; k < 9; ++k) { const double d = D[(size_t)c * 9 + k]; A[k * 9 + k] += d * d; }
  if (!invert_spd<9>(A, 9)) {
    atomicOr(error_flag, 2);
    for (int k = 0; k < 81; ++k) A[k] = 0.0;
  }
  for (int k = 0; k < 81; ++k) Minv[(size_t)c * 81 + k] = A[k];
}

// ------------------------------------------------------------------------------------------------
// seg_reduce9 with the metadata already in shared memory.
__device__ __forceinline__ void seg_reduce9_s(const Tile& q, const TileMetaSmem& m, const double* v, double* out, const int* spos) {
  for (int idx = threadIdx.x; idx < q.ns * 9; idx += T) {
    const int s = idx / 9, k = idx - s * 9;
    const int b = m.sptr[s], e = m.sptr[s + 1];
    double sum = 0.0;
    for (int pos = b; pos < e; ++pos) sum += v[k * VLD + m.sperm[pos]];
    out[(size_t)spos[s] * 9 + k] = sum;
  }
}

template <bool CHUNKED>
__global__ void __launch_bounds__(T, 768 / T) k_ba_matvec(BaDev L, const double2* __restrict__ J2, const double* __restrict__ p,
                                                    const double* __restrict__ zdir, const PcgDev* pcg,
                                                    const double* __restrict__ einv, double* __restrict__ seg_y,
                                                    const int* guard, int pf_dist) {
  if (guard != nullptr && *guard == 0) return;
  extern __shared__ double sm[];
  const Tile q = load_tile(L, blockIdx.x);
  if (q.chunk >= 0) return;                  // long tracks: k_ba_matvec_giant
  const int tid = threadIdx.x;
  double* xs = sm;                           // [max_seg][9]
  double* v = xs + L.max_seg_tile * 9;       // [9][VLD]
  double* w = v + 9 * VLD;                   // [3][T]  E^T t per observation.  (Aliasing w onto planes 0..2 of v saves 6 KB per
                                             // CTA and is bitwise-equivalent, but measured 5 % SLOWER on a B200: profiles/r01_v3_*.)
  double* u = w + 3 * T;                     // [3][T]  (E^T E)^-1 w per point
  double* ei = u + 3 * T;                    // [max_pt][6]
  TileMetaSmem meta;
  meta.sptr = reinterpret_cast<int*>(ei + L.max_pt_tile * 6);        // [max_seg + 1]
  meta.pptr = meta.sptr + L.max_seg_tile + 1;                        // [max_pt + 1]
  meta.sperm = reinterpret_cast<unsigned short*>(meta.pptr + L.max_pt_tile + 1);   // [T]
  double* ps = reinterpret_cast<double*>((reinterpret_cast<size_t>(meta.sperm + T) + 7) & ~(size_t)7);   // [seg_chunk_scratch] (CHUNKED)
  const RecView R = rec_view(L, L.tile_rec + (size_t)blockIdx.x * L.rec_stride);   // CHUNKED: chunk tables straight from global memory
  const bool active = tid < q.no;
  const int i = q.ob + tid;
  const size_t O = (size_t)L.n_obs;
  double2 Fv[9], Ev[3];
  int slot = 0, ptl = 0;
  if (active) {                              // issue the streaming loads first
#pragma unroll
    for (int k = 0; k < 9; ++k) Fv[k] = ldcs2(J2 + k * O + i);
#pragma unroll
    for (int k = 0; k < 3; ++k) Ev[k] = ldcs2(J2 + (9 + k) * O + i);
    slot = L.obs_slot[i]; ptl = L.obs_ptl[i];
  }
  // The tile that a CTA `pf_dist` places further down the grid will stream (about one wave of resident CTAs ahead) is
  // pulled into L2 now: 12 lanes issue one bulk prefetch each.  That CTA's loads then see L2 latency, not DRAM latency.
  if (pf_dist > 0 && tid < kJPlanes && (int)blockIdx.x + pf_dist < L.n_tiles) {
    const int tn = blockIdx.x + pf_dist;
    const int obn = L.tile_obs[tn], non = L.tile_obs[tn + 1] - obn;
    if (non > 0) l2_prefetch(J2 + tid * O + obn, (unsigned)non * 16u);
  }
  // everything the later phases need from global memory is requested now, behind the Jacobian loads
  stage_tile_meta(L, q, meta);
  for (int idx = tid; idx < q.np * 6; idx += T) ei[idx] = einv[(size_t)q.pb * 6 + idx];
  for (int idx = tid; idx < q.ns * 9; idx += T) {
    const int s = idx / 9, k = idx - s * 9;
    const size_t e = (size_t)L.seg_cam[q.sb + s] * 9 + k;
    xs[idx] = (pcg == nullptr) ? p[e] : ((pcg->iter == 1) ? zdir[e] : __fma_rn(pcg->beta, p[e], zdir[e]));
  }
  __syncthreads();
  double t0 = 0.0, t1 = 0.0;
  if (active) {
#pragma unroll
    for (int k = 0; k < 9; ++k) { const double xk = xs[slot * 9 + k]; t0 = __fma_rn(Fv[k].x, xk, t0); t1 = __fma_rn(Fv[k].y, xk, t1); }
#pragma unroll
    for (int k = 0; k < 3; ++k) w[CHUNKED ? tid * 3 + k : k * T + tid] = two_rows(Ev[k].x, t0, Ev[k].y, t1);
  }
  __syncthreads();
  if (CHUNKED) point_sums_chunked(R, q.np, w, v, ei, u, T);
  else {
    if (tid < q.np) {
      const int b = meta.pptr[tid], e = meta.pptr[tid + 1];
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      for (int j = b; j < e; ++j) { a0 += w[j]; a1 += w[T + j]; a2 += w[2 * T + j]; }
      const double* m = ei + tid * 6;
      u[tid] = m[0] * a0 + m[1] * a1 + m[2] * a2;
      u[T + tid] = m[1] * a0 + m[3] * a1 + m[4] * a2;
      u[2 * T + tid] = m[2] * a0 + m[4] * a1 + m[5] * a2;
    }
    __syncthreads();
  }
  if (active) {
    const double u0 = u[ptl], u1 = u[T + ptl], u2 = u[2 * T + ptl];
    const double s0 = __dsub_rn(t0, __fma_rn(Ev[2].x, u2, __fma_rn(Ev[1].x, u1, __dmul_rn(Ev[0].x, u0))));
    const double s1 = __dsub_rn(t1, __fma_rn(Ev[2].y, u2, __fma_rn(Ev[1].y, u1, __dmul_rn(Ev[0].y, u0))));
    double* vt = CHUNKED ? v + (int)R.srank[tid] * VS : v + tid;
#pragma unroll
    for (int k = 0; k < 9; ++k) vt[CHUNKED ? k : k * VLD] = two_rows(Fv[k].x, s0, Fv[k].y, s1);
  }
  __syncthreads();
  if (CHUNKED) seg_sums_chunked(R, q.ns, q.sb, v, ps, seg_y);
  else seg_reduce9_s(q, meta, v, seg_y, R.spos);
}

// ------------------------------------------------------------------------------------------------
// Persistent, fully prefetching variant of k_ba_matvec (ba_product.cuh: ProductPass).  Same arithmetic in the same order:
// the two kernels give bit-identical solves (test_matvec_kernels_agree_bitwise).
// TMAP: the Jacobian of a tile arrives as two [12 planes][128 observations] boxes of a 2-D tensor map over the plane-major
// array (2 TMA instructions per tile) instead of 12 one-plane bulk copies.
#ifndef SK_TMA_CTAS
#define SK_TMA_CTAS (512 / T)
#endif
template <bool TMAP, bool CHUNKED>
__global__ void __launch_bounds__(T, SK_TMA_CTAS) k_ba_matvec_tma(const __grid_constant__ CUtensorMap tmapJ, BaDev L, const double2* __restrict__ J2, const double* __restrict__ p,
                                                              const double* __restrict__ zdir, const PcgDev* pcg,
                                                              const double* __restrict__ einv, double* __restrict__ seg_y,
                                                              const int* guard) {
  if (guard != nullptr && *guard == 0) return;
  extern __shared__ __align__(128) double sm[];
  ProductPass<TMAP, CHUNKED, false> P;
  P.init(L, sm);
  // PCG direction z + beta p formed on the fly (p = z in iteration 1); the two scalars are read once, not per tile
  const bool dir_is_z = pcg != nullptr && pcg->iter == 1;
  const double beta = (pcg != nullptr && !dir_is_z) ? pcg->beta : 0.0;
  P.run(&tmapJ, L, J2, (pcg == nullptr) ? p : zdir, p, beta, pcg != nullptr && !dir_is_z, einv, seg_y, false, sm);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(T) k_ba_back_substitute(BaDev L, const double2* __restrict__ J2,
                                                          const double2* __restrict__ r2, const double* __restrict__ z,
                                                          const double* __restrict__ einv, double* __restrict__ step,
                                                          double* __restrict__ tile_mcc) {
  extern __shared__ double sm[];
  const Tile q = load_tile(L, blockIdx.x);
  if (q.chunk >= 0) return;                  // long tracks: k_ba_back_substitute_giant
  const int tid = threadIdx.x;
  double* zs = sm;                           // [max_seg][9]
  double* w = zs + L.max_seg_tile * 9;       // [3][T]
  double* u = w + 3 * T;                     // [3][T]
  double* red = u + 3 * T;                   // [8]
  const bool active = tid < q.no;
  const int i = q.ob + tid;
  const size_t O = (size_t)L.n_obs;
  double2 Fv[9], Ev[3], r = make_double2(0.0, 0.0);
  int slot = 0, ptl = 0;
  if (active) {
#pragma unroll
    for (int k = 0; k < 9; ++k) Fv[k] = ldcs2(J2 + k * O + i);
#pragma unroll
    for (int k = 0; k < 3; ++k) Ev[k] = ldcs2(J2 + (9 + k) * O + i);
    r = r2[i];
    slot = L.obs_slot[i]; ptl = L.obs_ptl[i];
  }
  for (int idx = tid; idx < q.ns * 9; idx += T) {
    const int s = idx / 9, k = idx - s * 9;
    zs[idx] = z[(size_t)L.seg_cam[q.sb + s] * 9 + k];
  }
  __syncthreads();
  double s0 = 0.0, s1 = 0.0;
  if (active) {
    double f0 = 0.0, f1 = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) { const double zk = zs[slot * 9 + k]; f0 += Fv[k].x * zk; f1 += Fv[k].y * zk; }
    s0 = r.x - f0; s1 = r.y - f1;             // b - F z
#pragma unroll
    for (int k = 0; k < 3; ++k) w[k * T + tid] = Ev[k].x * s0 + Ev[k].y * s1;
  }
  __syncthreads();
  if (tid < q.np) {
    const int pnt = q.pb + tid;
    const int b = L.pt_ptr[pnt] - q.ob, e = L.pt_ptr[pnt + 1] - q.ob;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int j = b; j < e; ++j) { a0 += w[j]; a1 += w[T + j]; a2 += w[2 * T + j]; }
    const double* m = einv + (size_t)pnt * 6;
    const double y0 = m[0] * a0 + m[1] * a1 + m[2] * a2;
    const double y1 = m[1] * a0 + m[3] * a1 + m[4] * a2;
    const double y2 = m[2] * a0 + m[4] * a1 + m[5] * a2;
    double* sp = step + (size_t)9 * L.n_cams + (size_t)pnt * 3;
    sp[0] = -y0; sp[1] = -y1; sp[2] = -y2;      // LM solves J y = r and steps by -y (A.4)
    u[tid] = -y0; u[T + tid] = -y1; u[2 * T + tid] = -y2;
  }
  __syncthreads();
  double mc = 0.0;
  if (active) {
    const double u0 = u[ptl], u1 = u[T + ptl], u2 = u[2 * T + ptl];
    // model residual m = J * step = F * (-z) + E * step_p = (s - r) + E * step_p
    const double m0 = (s0 - r.x) + (Ev[0].x * u0 + Ev[1].x * u1 + Ev[2].x * u2);
    const double m1 = (s1 - r.y) + (Ev[0].y * u0 + Ev[1].y * u1 + Ev[2].y * u2);
    mc = m0 * (r.x + m0 / 2.0) + m1 * (r.y + m1 / 2.0);
  }
  const double tot = block_sum(mc, red);
  if (tid == 0) tile_mcc[blockIdx.x] = tot;
}

// ------------------------------------------------------------------------------------------------
// Long tracks (more observations than one tile holds).  One CTA per long track g walks the chunk tiles
// [gp_tile_begin[g], +gp_tile_count[g]) of point gp_point[g] twice: pass 0 accumulates the per-point sums chunk by
// chunk (fixed order: block tree inside a chunk, chunks in tile order), pass 1 re-reads the chunk (48 KB, an L2 hit)
// and finishes exactly like the second half of the regular kernel.  The camera side (segments, k_cam_reduce) is the
// same as for regular tiles, so results stay atomic-free and bit-reproducible.  Real BAL files have a handful of
// such tracks; these kernels are about correctness on them, not about the roofline.
struct Giant { int pnt, tb, nc; };
__device__ __forceinline__ Giant load_giant(const BaDev& L, int g) { return {L.gp_point[g], L.gp_tile_begin[g], L.gp_tile_count[g]}; }

__global__ void k_ba_giant_point_combine(BaDev L, const double* __restrict__ chunk_pt, double* __restrict__ grad,
                                         double* __restrict__ cnorm2, const int* guard) {
  if (guard != nullptr && *guard == 0) return;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= L.n_giant) return;
  const Giant G = load_giant(L, g);
  const int c0 = -L.tile_np[G.tb] - 1;
  double s6[6] = {0, 0, 0, 0, 0, 0};
  for (int c = 0; c < G.nc; ++c)
#pragma unroll
    for (int k = 0; k < 6; ++k) s6[k] += chunk_pt[(size_t)(c0 + c) * 6 + k];
  const size_t o = (size_t)9 * L.n_cams + (size_t)G.pnt * 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) { grad[o + k] = s6[k]; cnorm2[o + k] = s6[3 + k]; }
}

__global__ void __launch_bounds__(T) k_ba_schur_setup_giant(BaDev L, const double2* __restrict__ J2, const double2* __restrict__ r2,
                                                            const double* __restrict__ D, double* __restrict__ einv,
                                                            double* __restrict__ seg_rhs, double* __restrict__ seg_M,
                                                            int* error_flag, int ftf_only) {
  extern __shared__ double sm[];
  const Giant G = load_giant(L, blockIdx.x);
  const int tid = threadIdx.x;
  double* v = sm;                      // [9][VLD]
  double* red = v + 9 * VLD;           // [9][8]
  const size_t O = (size_t)L.n_obs;
  // pass 0: E^T E (upper: 00 01 02 11 12 22) and E^T r over the whole track
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int c = 0; c < G.nc; ++c) {
    const Tile q = load_tile(L, G.tb + c);
    double x[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (tid < q.no) {
      const int i = q.ob + tid;
      double2 Ev[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) Ev[k] = J2[(9 + k) * O + i];
      const double2 r = r2[i];
      x[0] = Ev[0].x * Ev[0].x + Ev[0].y * Ev[0].y;
      x[1] = Ev[0].x * Ev[1].x + Ev[0].y * Ev[1].y;
      x[2] = Ev[0].x * Ev[2].x + Ev[0].y * Ev[2].y;
      x[3] = Ev[1].x * Ev[1].x + Ev[1].y * Ev[1].y;
      x[4] = Ev[1].x * Ev[2].x + Ev[1].y * Ev[2].y;
      x[5] = Ev[2].x * Ev[2].x + Ev[2].y * Ev[2].y;
#pragma unroll
      for (int k = 0; k < 3; ++k) x[6 + k] = Ev[k].x * r.x + Ev[k].y * r.y;
    }
    block_sum_all<9>(x, red);
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] += x[k];
  }
  // every thread holds the same sums: invert redundantly, thread 0 publishes
  const double* Dp = D + (size_t)9 * L.n_cams + (size_t)G.pnt * 3;
  double m[6] = {acc[0] + Dp[0] * Dp[0], acc[1], acc[2], acc[3] + Dp[1] * Dp[1], acc[4], acc[5] + Dp[2] * Dp[2]};
  double pi[9];
  if (!invert_spd3(m, pi)) {
    if (tid == 0) atomicOr(error_flag, 1);
#pragma unroll
    for (int k = 0; k < 6; ++k) pi[k] = 0.0;
  }
  if (tid < 6) einv[(size_t)G.pnt * 6 + tid] = pi[tid];
  pi[6] = pi[0] * acc[6] + pi[1] * acc[7] + pi[2] * acc[8];
  pi[7] = pi[1] * acc[6] + pi[3] * acc[7] + pi[4] * acc[8];
  pi[8] = pi[2] * acc[6] + pi[4] * acc[7] + pi[5] * acc[8];
  // pass 1: reduced rhs and diagonal-block partials of every chunk
  for (int c = 0; c < G.nc; ++c) {
    const Tile q = load_tile(L, G.tb + c);
    const bool active = tid < q.no;
    const int i = q.ob + tid;
    double2 Fv[9], Ev[3], r = make_double2(0.0, 0.0);
    double Gm[3][9], H[3][9];
    if (active) {
#pragma unroll
      for (int k = 0; k < 9; ++k) Fv[k] = J2[k * O + i];
#pragma unroll
      for (int k = 0; k < 3; ++k) Ev[k] = J2[(9 + k) * O + i];
      r = r2[i];
      const double q0 = r.x - (Ev[0].x * pi[6] + Ev[1].x * pi[7] + Ev[2].x * pi[8]);
      const double q1 = r.y - (Ev[0].y * pi[6] + Ev[1].y * pi[7] + Ev[2].y * pi[8]);
#pragma unroll
      for (int k = 0; k < 9; ++k) v[k * VLD + tid] = Fv[k].x * q0 + Fv[k].y * q1;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
#pragma unroll
        for (int a = 0; a < 3; ++a) Gm[a][k] = Ev[a].x * Fv[k].x + Ev[a].y * Fv[k].y;
        H[0][k] = pi[0] * Gm[0][k] + pi[1] * Gm[1][k] + pi[2] * Gm[2][k];
        H[1][k] = pi[1] * Gm[0][k] + pi[3] * Gm[1][k] + pi[4] * Gm[2][k];
        H[2][k] = pi[2] * Gm[0][k] + pi[4] * Gm[1][k] + pi[5] * Gm[2][k];
      }
    }
    __syncthreads();
    seg_reduce9(L, q, v, seg_rhs, 9, 0);
#pragma unroll
    for (int round = 0; round < 5; ++round) {
      __syncthreads();
      if (active) {
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const int k = upper_row(round * 9 + j), l = upper_col(round * 9 + j);
          const double ftf = Fv[k].x * Fv[l].x + Fv[k].y * Fv[l].y;
          v[j * VLD + tid] = ftf_only ? ftf : (ftf - (Gm[0][k] * H[0][l] + Gm[1][k] * H[1][l] + Gm[2][k] * H[2][l]));
        }
      }
      __syncthreads();
      seg_reduce9(L, q, v, seg_M, 45, round * 9);
    }
    __syncthreads();                   // v is rewritten by the next chunk
  }
}

// Input vector exactly as in k_ba_matvec (p, or the PCG direction z + beta p).
__global__ void __launch_bounds__(T) k_ba_matvec_giant(BaDev L, const double2* __restrict__ J2, const double* __restrict__ p,
                                                       const double* __restrict__ zdir, const PcgDev* pcg,
                                                       const double* __restrict__ einv, double* __restrict__ seg_y,
                                                       const int* guard) {
  if (guard != nullptr && *guard == 0) return;
  extern __shared__ double sm[];
  const Giant G = load_giant(L, blockIdx.x);
  const int tid = threadIdx.x;
  double* xs = sm;                           // [max_seg][9]
  double* v = xs + L.max_seg_tile * 9;       // [9][VLD]
  double* red = v + 9 * VLD;                 // [3][8]
  const size_t O = (size_t)L.n_obs;
  double a[3] = {0, 0, 0}, u[3] = {0, 0, 0};
  for (int pass = 0; pass < 2; ++pass) {
    for (int c = 0; c < G.nc; ++c) {
      const Tile q = load_tile(L, G.tb + c);
      const bool active = tid < q.no;
      const int i = q.ob + tid;
      double2 Fv[9], Ev[3];
      int slot = 0;
      if (active) {
#pragma unroll
        for (int k = 0; k < 9; ++k) Fv[k] = J2[k * O + i];
#pragma unroll
        for (int k = 0; k < 3; ++k) Ev[k] = J2[(9 + k) * O + i];
        slot = L.obs_slot[i];
      }
      for (int idx = tid; idx < q.ns * 9; idx += T) {
        const int s = idx / 9, k = idx - s * 9;
        const size_t e = (size_t)L.seg_cam[q.sb + s] * 9 + k;
        xs[idx] = (pcg == nullptr) ? p[e] : ((pcg->iter == 1) ? zdir[e] : __fma_rn(pcg->beta, p[e], zdir[e]));
      }
      __syncthreads();
      double t0 = 0.0, t1 = 0.0;
      if (active) {
#pragma unroll
        for (int k = 0; k < 9; ++k) { const double xk = xs[slot * 9 + k]; t0 += Fv[k].x * xk; t1 += Fv[k].y * xk; }
      }
      if (pass == 0) {
        double w[3] = {0, 0, 0};
        if (active) {
#pragma unroll
          for (int k = 0; k < 3; ++k) w[k] = Ev[k].x * t0 + Ev[k].y * t1;
        }
        block_sum_all<3>(w, red);            // ends with a barrier: xs may be restaged
#pragma unroll
        for (int k = 0; k < 3; ++k) a[k] += w[k];
      } else {
        if (active) {
          const double s0 = t0 - (Ev[0].x * u[0] + Ev[1].x * u[1] + Ev[2].x * u[2]);
          const double s1 = t1 - (Ev[0].y * u[0] + Ev[1].y * u[1] + Ev[2].y * u[2]);
#pragma unroll
          for (int k = 0; k < 9; ++k) v[k * VLD + tid] = Fv[k].x * s0 + Fv[k].y * s1;
        }
        __syncthreads();
        seg_reduce9(L, q, v, seg_y, 9, 0, L.seg_pos);
        __syncthreads();
      }
    }
    if (pass == 0) {
      const double* m = einv + (size_t)G.pnt * 6;
      u[0] = m[0] * a[0] + m[1] * a[1] + m[2] * a[2];
      u[1] = m[1] * a[0] + m[3] * a[1] + m[4] * a[2];
      u[2] = m[2] * a[0] + m[4] * a[1] + m[5] * a[2];
    }
  }
}

__global__ void __launch_bounds__(T) k_ba_back_substitute_giant(BaDev L, const double2* __restrict__ J2, const double2* __restrict__ r2,
                                                                const double* __restrict__ z, const double* __restrict__ einv,
                                                                double* __restrict__ step, double* __restrict__ tile_mcc) {
  extern __shared__ double sm[];
  const Giant G = load_giant(L, blockIdx.x);
  const int tid = threadIdx.x;
  double* zs = sm;                           // [max_seg][9]
  double* red = zs + L.max_seg_tile * 9;     // [3][8]
  const size_t O = (size_t)L.n_obs;
  double a[3] = {0, 0, 0}, u[3] = {0, 0, 0};
  for (int pass = 0; pass < 2; ++pass) {
    for (int c = 0; c < G.nc; ++c) {
      const Tile q = load_tile(L, G.tb + c);
      const bool active = tid < q.no;
      const int i = q.ob + tid;
      double2 Fv[9], Ev[3], r = make_double2(0.0, 0.0);
      int slot = 0;
      if (active) {
#pragma unroll
        for (int k = 0; k < 9; ++k) Fv[k] = J2[k * O + i];
#pragma unroll
        for (int k = 0; k < 3; ++k) Ev[k] = J2[(9 + k) * O + i];
        r = r2[i];
        slot = L.obs_slot[i];
      }
      for (int idx = tid; idx < q.ns * 9; idx += T) {
        const int s = idx / 9, k = idx - s * 9;
        zs[idx] = z[(size_t)L.seg_cam[q.sb + s] * 9 + k];
      }
      __syncthreads();
      double s0 = 0.0, s1 = 0.0;
      if (active) {
        double f0 = 0.0, f1 = 0.0;
#pragma unroll
        for (int k = 0; k < 9; ++k) { const double zk = zs[slot * 9 + k]; f0 += Fv[k].x * zk; f1 += Fv[k].y * zk; }
        s0 = r.x - f0; s1 = r.y - f1;
      }
      if (pass == 0) {
        double w[3] = {0, 0, 0};
        if (active) {
#pragma unroll
          for (int k = 0; k < 3; ++k) w[k] = Ev[k].x * s0 + Ev[k].y * s1;
        }
        block_sum_all<3>(w, red);
#pragma unroll
        for (int k = 0; k < 3; ++k) a[k] += w[k];
      } else {
        double mc[1] = {0.0};
        if (active) {
          const double m0 = (s0 - r.x) + (Ev[0].x * u[0] + Ev[1].x * u[1] + Ev[2].x * u[2]);
          const double m1 = (s1 - r.y) + (Ev[0].y * u[0] + Ev[1].y * u[1] + Ev[2].y * u[2]);
          mc[0] = m0 * (r.x + m0 / 2.0) + m1 * (r.y + m1 / 2.0);
        }
        block_sum_all<1>(mc, red);
        if (tid == 0) tile_mcc[G.tb + c] = mc[0];
      }
    }
    if (pass == 0) {
      const double* m = einv + (size_t)G.pnt * 6;
      const double y0 = m[0] * a[0] + m[1] * a[1] + m[2] * a[2];
      const double y1 = m[1] * a[0] + m[3] * a[1] + m[4] * a[2];
      const double y2 = m[2] * a[0] + m[4] * a[1] + m[5] * a[2];
      u[0] = -y0; u[1] = -y1; u[2] = -y2;
      if (tid < 3) step[(size_t)9 * L.n_cams + (size_t)G.pnt * 3 + tid] = u[tid];
    }
  }
}

__global__ void k_ba_export_jacobian(BaDev L, const double2* __restrict__ J2, double* __restrict__ F, double* __restrict__ E) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L.n_obs) return;
  const size_t O = (size_t)L.n_obs;
  for (int k = 0; k < 9; ++k) { const double2 f = J2[k * O + i]; F[(size_t)i * 18 + k] = f.x; F[(size_t)i * 18 + 9 + k] = f.y; }
  for (int k = 0; k < 3; ++k) { const double2 e = J2[(9 + k) * O + i]; E[(size_t)i * 6 + k] = e.x; E[(size_t)i * 6 + 3 + k] = e.y; }
}

// ------------------------------------------------------------------------------------------------
// Per-tile metadata records of the implicit-Schur product (ba_tile_rec.h), packed ON THE DEVICE from the layout arrays that are
// uploaded anyway: one CTA per tile.  Byte for byte what the host builder (ba_layout.cu: build_tile_records, kept for the
// host-side checks and as SKERES_TILE_REC=host) writes into a zero-filled buffer; packing 19,539 records of the Venice shape
// took 35 ms on the host and another 66 MB of upload -- a third of the solver's set-up time.
__global__ void __launch_bounds__(T) k_ba_build_tile_records(BaDev L, unsigned char* __restrict__ rec, int stride, int sp, int pp, int sc) {
  const int t = blockIdx.x, tid = threadIdx.x;
  unsigned char* base = rec + (size_t)t * stride;
  unsigned short* slot = reinterpret_cast<unsigned short*>(base); unsigned short* ptl = slot + T; unsigned short* sperm = ptl + T;
  unsigned short* srank = sperm + T;
  int* sptr = reinterpret_cast<int*>(base + 8 * T); int* pptr = sptr + sp; int* scam = pptr + pp; int* spos = scam + sp;
  unsigned short* pchunk = reinterpret_cast<unsigned short*>(spos + sp); unsigned short* pcptr = pchunk + T; unsigned short* schunk = pcptr + pp;
  unsigned short* scptr = schunk + sc;
  const int ob = L.tile_obs[t], no = L.tile_obs[t + 1] - ob, pb = L.tile_pt[t];
  const int npe = L.tile_np[t];                            // > 0: points of a regular tile; < 0: chunk tile of a long track
  const int sb = L.tile_seg[t], ns = L.tile_seg[t + 1] - sb;
  __shared__ int s_sptr[T + 1], s_pptr[T + 1], s_cnt[T + 1];
  if (tid < no) {
    const unsigned short sp_ = L.seg_perm[ob + tid];
    slot[tid] = L.obs_slot[ob + tid]; ptl[tid] = L.obs_ptl[ob + tid]; sperm[tid] = sp_;
    srank[sp_] = (unsigned short)tid;
  }
  for (int s = tid; s <= ns; s += T) { const int v = L.seg_ptr[sb + s] - ob; sptr[s] = v; s_sptr[s] = v; }
  for (int s = tid; s < ns; s += T) { scam[s] = L.seg_cam[sb + s]; spos[s] = L.seg_pos[sb + s]; }
  __syncthreads();
  // chunk tables of the two-level sums: first chunk of every segment = exclusive prefix sum of the segments' chunk counts
  if (tid == 0) {
    int nc = 0;
    for (int s = 0; s < ns; ++s) { s_cnt[s] = nc; nc += (s_sptr[s + 1] - s_sptr[s] + kSegChunk - 1) / kSegChunk; }
    s_cnt[ns] = nc;
  }
  __syncthreads();
  for (int s = tid; s <= ns; s += T) scptr[s] = (unsigned short)s_cnt[s];
  for (int s = tid; s < ns; s += T) {
    int nc = s_cnt[s];
    for (int b = s_sptr[s]; b < s_sptr[s + 1]; b += kSegChunk) schunk[nc++] = (unsigned short)(b | ((min(kSegChunk, s_sptr[s + 1] - b) - 1) << 8));
  }
  if (npe > 0) {                                           // chunk tiles of long tracks have no per-point phase in the tile kernels
    const int np = npe;
    __syncthreads();
    for (int q = tid; q <= np; q += T) { const int v = L.pt_ptr[pb + q] - ob; pptr[q] = v; s_pptr[q] = v; }
    __syncthreads();
    if (tid == 0) {
      int nc = 0;
      for (int q = 0; q < np; ++q) { s_cnt[q] = nc; nc += (s_pptr[q + 1] - s_pptr[q] + kPtChunk - 1) / kPtChunk; }
      s_cnt[np] = nc;
    }
    __syncthreads();
    for (int q = tid; q <= np; q += T) pcptr[q] = (unsigned short)s_cnt[q];
    for (int q = tid; q < np; q += T) {
      int nc = s_cnt[q];
      for (int b = s_pptr[q]; b < s_pptr[q + 1]; b += kPtChunk) pchunk[nc++] = (unsigned short)(b | ((min(kPtChunk, s_pptr[q + 1] - b) - 1) << 8));
    }
  }
}

template <class K>
void set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) SK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

}  // namespace

void launch_ba_evaluate(const BaDev& L, const double* x, const double* scale, LossSpec loss, bool with_jacobian,
                        bool write_jacobian, double2* J2, double2* r2, double* grad, double* cnorm2, double* seg_g,
                        double* seg_n, double* tile_cost, double* chunk_pt, int* fail_flag, const int* guard, cudaStream_t s,
                        int functor_id) {
  if (L.n_tiles == 0) return;
  const bool user = functor_id >= kUserFunctorBase;      // a functor of the same shape compiled from source (user_functor.cu)
  SK_REQUIRE(user || functor_id == SK_FUNCTOR_SNAVELY_REPROJECTION_ERROR, SK_ERR_INTERNAL, "tile evaluation: functor %d has no tile kernel", functor_id);
  if (with_jacobian) {
    const size_t smem = sizeof(double) * ((size_t)2 * (L.max_seg_tile * 9 + L.max_pt_tile * 3) + 18 * VLD + 8) +
                        sizeof(int) * ((size_t)L.max_seg_tile + L.max_pt_tile + 2) + sizeof(unsigned short) * T;
    if (user) {
      launch_user_ba_evaluate(functor_id, true, smem, L, x, scale, loss, write_jacobian ? 1 : 0, J2, r2, grad, cnorm2, seg_g, seg_n,
                              tile_cost, chunk_pt, fail_flag, guard, s);
    } else {
      set_smem(k_ba_evaluate<true>, smem);
      k_ba_evaluate<true><<<L.n_tiles, T, smem, s>>>(L, x, scale, loss, write_jacobian ? 1 : 0, J2, r2, grad, cnorm2, seg_g,
                                                     seg_n, tile_cost, chunk_pt, fail_flag, guard);
    }
    if (L.n_giant) k_ba_giant_point_combine<<<cdiv(L.n_giant, 64), 64, 0, s>>>(L, chunk_pt, grad, cnorm2, guard);
  } else {
    const size_t smem = sizeof(double) * ((size_t)(L.max_seg_tile * 9 + L.max_pt_tile * 3) + 8);
    if (user) {
      launch_user_ba_evaluate(functor_id, false, smem, L, x, scale, loss, 0, J2, r2, grad, cnorm2, seg_g, seg_n, tile_cost, chunk_pt,
                              fail_flag, guard, s);
    } else {
      set_smem(k_ba_evaluate<false>, smem);
      k_ba_evaluate<false><<<L.n_tiles, T, smem, s>>>(L, x, scale, loss, 0, J2, r2, grad, cnorm2, seg_g, seg_n, tile_cost,
                                                      chunk_pt, fail_flag, guard);
    }
  }
  check_launch("k_ba_evaluate");
}

void launch_cam_reduce(const BaDev& L, int K, const double* seg, double* out, const int* guard, cudaStream_t s) {
  SK_REQUIRE(K >= 1 && K <= kCamReduceThreads, SK_ERR_INTERNAL, "k_cam_reduce: bad K");
  if (L.n_cams == 0) return;
  k_cam_reduce<<<L.n_cams, kCamReduceThreads, 0, s>>>(L, K, seg, out, guard);
  check_launch("k_cam_reduce");
}

void launch_ba_schur_setup(const BaDev& L, const double2* J2, const double2* r2, const double* D, double* einv,
                           double* seg_rhs, double* seg_M, int* error_flag, bool ftf_only, cudaStream_t s) {
  if (L.n_tiles == 0) return;
  const size_t smem = sizeof(double) * ((size_t)kSetupPlanes * VLD + (size_t)L.max_pt_tile * 9) +
                      sizeof(int) * ((size_t)L.max_seg_tile + L.max_pt_tile + 2) + sizeof(unsigned short) * T;
  set_smem(k_ba_schur_setup, smem);
  k_ba_schur_setup<<<L.n_tiles, T, smem, s>>>(L, J2, r2, D, einv, seg_rhs, seg_M, error_flag, ftf_only ? 1 : 0);
  if (L.n_giant) {
    const size_t smem_g = sizeof(double) * ((size_t)9 * VLD + 9 * 8);
    set_smem(k_ba_schur_setup_giant, smem_g);
    k_ba_schur_setup_giant<<<L.n_giant, T, smem_g, s>>>(L, J2, r2, D, einv, seg_rhs, seg_M, error_flag, ftf_only ? 1 : 0);
  }
  check_launch("k_ba_schur_setup");
}

void launch_ba_precond_invert(const BaDev& L, const double* M45, const double* D, double* Minv, int* error_flag,
                              cudaStream_t s) {
  k_ba_precond_invert<<<cdiv(L.n_cams, 64), 64, 0, s>>>(L, M45, D, Minv, error_flag);
  check_launch("k_ba_precond_invert");
}

void launch_ba_matvec(const BaDev& L, const double2* J2, const double* p, const double* zdir, const PcgDev* pcg, const double* einv,
                      double* seg_y, const int* guard, cudaStream_t s, const CUtensorMap* tmapJ) {
  if (L.n_tiles == 0) return;
  if (L.n_giant) {             // long tracks first (any order works: the two kernels write disjoint segments)
    const size_t smem_g = sizeof(double) * ((size_t)L.max_seg_tile * 9 + 9 * VLD + 3 * 8);
    set_smem(k_ba_matvec_giant, smem_g);
    k_ba_matvec_giant<<<L.n_giant, T, smem_g, s>>>(L, J2, p, zdir, pcg, einv, seg_y, guard);
    check_launch("k_ba_matvec_giant");
  }
  const bool chunked = !L.matvec_serial_sums && L.tile_rec != nullptr;
  const size_t smem_tail = sizeof(double) * ((size_t)L.max_pt_tile * 6) + sizeof(int) * ((size_t)L.max_seg_tile + L.max_pt_tile + 2) +
                           sizeof(unsigned short) * T + 16 + sizeof(double) * (size_t)seg_chunk_scratch(L.max_seg_tile);
  const size_t smem = sizeof(double) * ((size_t)L.max_seg_tile * 9 + 9 * VLD + 6 * T) + smem_tail;
  if (L.matvec_classic != 1 && L.tile_rec != nullptr) {
    // Default: persistent TMA-prefetching kernel, one wave of as many CTAs per SM as its shared memory allows.
    const size_t smem_p = ProductPass<true, false, false>::smem_bytes(L);
    static int sms = 0, smem_max = 0;
    if (sms == 0) {
      int dev = 0; SK_CUDA(cudaGetDevice(&dev));
      SK_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
      SK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    if (smem_p <= (size_t)smem_max) {
      static size_t cfg_smem[4] = {0}; static int per_sm[4] = {0};   // function attributes / occupancy, redone when the size changes
      const int v = (tmapJ != nullptr ? 1 : 0) + (chunked ? 2 : 0);
      using Kernel = void (*)(const CUtensorMap, BaDev, const double2*, const double*, const double*, const PcgDev*, const double*, double*, const int*);
      static const Kernel kernels[4] = {k_ba_matvec_tma<false, false>, k_ba_matvec_tma<true, false>, k_ba_matvec_tma<false, true>, k_ba_matvec_tma<true, true>};
      const Kernel kernel = kernels[v];
      if (smem_p != cfg_smem[v]) {
        set_smem(kernel, smem_p);
        SK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        SK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[v], kernel, T, smem_p));
        cfg_smem[v] = smem_p;
      }
      if (per_sm[v] > 0) {
        const int grid = std::min(L.n_tiles, per_sm[v] * sms);
        static const CUtensorMap no_map{};
        kernel<<<grid, T, smem_p, s>>>((v & 1) ? *tmapJ : no_map, L, J2, p, zdir, pcg, einv, seg_y, guard);
        check_launch("k_ba_matvec_tma");
        return;
      }
    }
  }
  // Classic kernel (one CTA per tile, Jacobian straight into registers): SKERES_MATVEC=classic, or when the per-tile
  // maxima of a problem make the prefetching kernel's shared memory exceed what one CTA may have.
  const size_t smem_v2 = smem;
  // development knob: SKERES_MATVEC_PFDIST=<tiles> L2 prefetch distance (0 = off; measured +5 %, profiles/r01_v5_*)
  static const int pf_dist = [] { const char* e = getenv("SKERES_MATVEC_PFDIST"); return e ? atoi(e) : 0; }();
  if (chunked) {
    set_smem(k_ba_matvec<true>, smem_v2);
    k_ba_matvec<true><<<L.n_tiles, T, smem_v2, s>>>(L, J2, p, zdir, pcg, einv, seg_y, guard, pf_dist);
  } else {
    set_smem(k_ba_matvec<false>, smem_v2);
    k_ba_matvec<false><<<L.n_tiles, T, smem_v2, s>>>(L, J2, p, zdir, pcg, einv, seg_y, guard, pf_dist);
  }
  check_launch("k_ba_matvec");
}

// 2-D view of the plane-major Jacobian for TMA: inner dimension = the 2 * n_obs doubles of one plane, outer = 12 planes.
// A box is [12 planes][128 observations] = 24 KB.  The encoder lives in libcuda; it is fetched through the runtime so that
// the library keeps no link-time dependency on the driver library.
bool make_jacobian_tensor_map(const double2* J2, int n_obs, CUtensorMap* out) {
  if (n_obs < T / 2) return false;                              // the box would exceed the tensor
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = [] {
    void* fn = nullptr; cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) fn = nullptr;
    return reinterpret_cast<EncodeFn>(fn);
  }();
  if (encode == nullptr) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)2 * (cuuint64_t)n_obs, (cuuint64_t)kJPlanes};
  const cuuint64_t strides[1] = {(cuuint64_t)n_obs * 16};      // bytes between planes
  const cuuint32_t box[2] = {(cuuint32_t)T, (cuuint32_t)kJPlanes};   // 256 doubles = 128 observations
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double2*>(J2), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

void launch_ba_back_substitute(const BaDev& L, const double2* J2, const double2* r2, const double* z, const double* einv,
                               double* step, double* tile_mcc, cudaStream_t s) {
  if (L.n_tiles == 0) return;
  const size_t smem = sizeof(double) * ((size_t)L.max_seg_tile * 9 + 6 * T + 8);
  set_smem(k_ba_back_substitute, smem);
  k_ba_back_substitute<<<L.n_tiles, T, smem, s>>>(L, J2, r2, z, einv, step, tile_mcc);
  if (L.n_giant) {
    const size_t smem_g = sizeof(double) * ((size_t)L.max_seg_tile * 9 + 3 * 8);
    set_smem(k_ba_back_substitute_giant, smem_g);
    k_ba_back_substitute_giant<<<L.n_giant, T, smem_g, s>>>(L, J2, r2, z, einv, step, tile_mcc);
  }
  check_launch("k_ba_back_substitute");
}

void launch_ba_build_tile_records(const BaDev& L, unsigned char* rec, const TileRecDims& d, cudaStream_t s) {
  if (L.n_tiles == 0) return;
  SK_CUDA(cudaMemsetAsync(rec, 0, (size_t)L.n_tiles * d.stride, s));
  k_ba_build_tile_records<<<L.n_tiles, T, 0, s>>>(L, rec, d.stride, d.sp, d.pp, d.sc);
  check_launch("k_ba_build_tile_records");
}

void launch_ba_export_jacobian(const BaDev& L, const double2* J2, double* F, double* E, cudaStream_t s) {
  k_ba_export_jacobian<<<cdiv(L.n_obs, 256), 256, 0, s>>>(L, J2, F, E);
  check_launch("k_ba_export_jacobian");
}

}  // namespace sk
