// ba_evaluate.cuh — the tile evaluation kernel of the bundle-adjustment path as a device template over the cost functor
// (ProgramEvaluator / ResidualBlock::Evaluate / Corrector, SURVEY.md A.2, on top of AutodiffCostFunction.scala:74-134).
// Instantiated in ba_kernels.cu with the built-in SnavelyReprojectionError (SimpleBundleAdjuster.scala:79-119) and -- this
// header is handed to NVRTC verbatim -- in the translation unit generated for a functor of shape (2; 9, 3) given as source
// (user_functor.cu), so that a user's camera model runs in the same tile kernel, with the same reductions, as the built-in one.
// Device code only; depends on ba_dev.cuh and jet.cuh (LossSpec, Corrector).
#pragma once
#include "ba_tile.cuh"
#include "jet.cuh"

namespace sk {
namespace {

// Deterministic block sum (fixed shuffle tree, then warp 0 over the 8 warp totals).
__device__ __forceinline__ double block_sum(double x, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = x;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < (int)(blockDim.x >> 5)) ? red[l] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
  }
  return r;   // valid in thread 0
}

// Tile metadata needed by the later phases, fetched at the top of the kernel so that its latency
// overlaps the streaming Jacobian loads: tile-local segment permutation / segment starts / point starts.
struct TileMetaSmem { unsigned short* sperm; int* sptr; int* pptr; };

__device__ __forceinline__ void stage_tile_meta(const BaDev& L, const Tile& q, const TileMetaSmem& m) {
  const int tid = threadIdx.x;
  if (tid < q.no) m.sperm[tid] = L.seg_perm[q.ob + tid];
  for (int idx = tid; idx <= q.ns; idx += T) m.sptr[idx] = L.seg_ptr[q.sb + idx] - q.ob;
  for (int idx = tid; idx <= q.np; idx += T) m.pptr[idx] = L.pt_ptr[q.pb + idx] - q.ob;
}

// Segment sums of NP staged planes with the tile's segment structure in shared memory: item (s, k) adds plane k over
// segment s in the fixed (point) order, four staged values in flight at a time.  k < split goes to out0 (row stride
// s0), the rest to out1 (row stride s1, offset o1).
template <int NP>
__device__ __forceinline__ void seg_reduce_planes(const Tile& q, const TileMetaSmem& m, const double* v, int split,
                                                  double* out0, int s0, double* out1, int s1, int o1) {
  for (int idx = threadIdx.x; idx < q.ns * NP; idx += T) {
    const int s = idx / NP, k = idx - s * NP;
    const int b = m.sptr[s], e = m.sptr[s + 1];
    const double* vk = v + k * VLD;
    double sum = 0.0;
    int pos = b;
    for (; pos + 4 <= e; pos += 4) {
      const int i0 = m.sperm[pos], i1 = m.sperm[pos + 1], i2 = m.sperm[pos + 2], i3 = m.sperm[pos + 3];
      const double x0 = vk[i0], x1 = vk[i1], x2 = vk[i2], x3 = vk[i3];
      sum += x0; sum += x1; sum += x2; sum += x3;
    }
    for (; pos < e; ++pos) sum += vk[m.sperm[pos]];
    if (k < split) out0[(size_t)(q.sb + s) * s0 + k] = sum;
    else out1[(size_t)(q.sb + s) * s1 + o1 + (k - split)] = sum;
  }
}

// ------------------------------------------------------------------------------------------------
// The body of the tile evaluation kernel (one CTA of T threads per tile, thread i = observation i of the tile): residuals,
// robust-loss correction, cost partial; JAC: the Jacobian (stored column-scaled), the gradient and the squared column norms.
//   x            state [9C + 3P]
//   scale        Jacobi column scaling [9C + 3P] or nullptr (= 1)
//   J2, r2       outputs (only when write_j)
//   grad, cnorm2 POINT part [9C .. 9C+3P) written directly; the camera part goes to seg_g / seg_n partials [S][9]
//   tile_cost    [n_tiles] partial costs
//   chunk_pt     [n_chunks][6] point-part partials of the chunk tiles of long tracks
// FUN: the cost functor of shape (2; 9, 3) with two constants (the observed x, y):
//   static bool FUN::residual(cam, pt, ox, oy, res[2]);  static bool FUN::residual_jacobian(cam, pt, ox, oy, res[2], F[18], E[6])
// F = d res / d camera (2 x 9 row-major), E = d res / d point (2 x 3 row-major); false = the functor failed (CostFunctor.scala:15-26).
// Dynamic shared memory: ba_evaluate_smem_bytes.
template <bool JAC, class FUN>
__device__ __forceinline__ void ba_evaluate_tile(const BaDev& L, const double* __restrict__ x, const double* __restrict__ scale,
                                                 const LossSpec& loss, int write_j, double2* __restrict__ J2,
                                                 double2* __restrict__ r2, double* __restrict__ grad,
                                                 double* __restrict__ cnorm2, double* __restrict__ seg_g,
                                                 double* __restrict__ seg_n, double* __restrict__ tile_cost,
                                                 double* __restrict__ chunk_pt, int* fail_flag, const int* guard) {
  if (guard != nullptr && *guard == 0) return;
  extern __shared__ double sm[];
  const Tile q = load_tile(L, blockIdx.x);
  const int tid = threadIdx.x;
  double* cam_s = sm;                               // [max_seg][9]
  double* pt_s = cam_s + L.max_seg_tile * 9;        // [max_pt][3]
  double* csc_s = pt_s + L.max_pt_tile * 3;         // [max_seg][9]  column scales (JAC)
  double* psc_s = csc_s + L.max_seg_tile * 9;       // [max_pt][3]
  double* v = psc_s + L.max_pt_tile * 3;            // [18][VLD]     (JAC)
  double* red = JAC ? (v + 18 * VLD) : csc_s;       // [8]
  TileMetaSmem meta;                                // the tile's segment structure (JAC), fetched behind the functor's arithmetic
  meta.sptr = reinterpret_cast<int*>(red + 8);                       // [max_seg + 1]
  meta.pptr = meta.sptr + L.max_seg_tile + 1;                        // [max_pt + 1]
  meta.sperm = reinterpret_cast<unsigned short*>(meta.pptr + L.max_pt_tile + 1);   // [T]
  if (JAC) stage_tile_meta(L, q, meta);
  const size_t pbase = (size_t)9 * L.n_cams;
  for (int idx = tid; idx < q.ns * 9; idx += T) {
    const int s = idx / 9, k = idx - s * 9;
    const int c = L.seg_cam[q.sb + s];
    cam_s[idx] = x[(size_t)c * 9 + k];
    if (JAC) csc_s[idx] = scale ? scale[(size_t)c * 9 + k] : 1.0;
  }
  for (int idx = tid; idx < q.np * 3; idx += T) {
    pt_s[idx] = x[pbase + (size_t)q.pb * 3 + idx];
    if (JAC) psc_s[idx] = scale ? scale[pbase + (size_t)q.pb * 3 + idx] : 1.0;
  }
  __syncthreads();
  const bool active = tid < q.no;
  const int i = q.ob + tid;
  double cost = 0.0, res[2] = {0.0, 0.0};
  double F[18], E[6];
  int slot = 0, ptl = 0;
  if (active) {
    const double2 o = L.obs[i];
    slot = L.obs_slot[i]; ptl = L.obs_ptl[i];
    bool ok;
    if (JAC) ok = FUN::residual_jacobian(cam_s + slot * 9, pt_s + ptl * 3, o.x, o.y, res, F, E);
    else ok = FUN::residual(cam_s + slot * 9, pt_s + ptl * 3, o.x, o.y, res);
    if (!ok) atomicOr(fail_flag, 1);
    const double sq = res[0] * res[0] + res[1] * res[1];
    double rho[3];
    loss_evaluate(loss, sq, rho);
    cost = 0.5 * rho[0];
    if (!(cost == cost)) atomicOr(fail_flag, 1);    // NaN residual == failed evaluation
    if (JAC && loss.type != SK_LOSS_TRIVIAL) {
      const Corrector corr(sq, rho);
      corr.correct_jacobian(2, 9, 9, res, F);
      corr.correct_jacobian(2, 3, 3, res, E);
      corr.correct_residuals(2, res);
    }
  }
  const double csum = block_sum(cost, red);
  if (tid == 0) tile_cost[blockIdx.x] = csum;
  if (!JAC) return;
  if (active) {
    // point part: gradient (unscaled J, as the evaluator computes it) and squared column norms of
    // the column-scaled J (what the LM strategy sees after TrustRegionMinimizer scales in place).
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double sc = psc_s[ptl * 3 + k];
      const double e0 = E[k] * sc, e1 = E[3 + k] * sc;
      v[k * VLD + tid] = E[k] * res[0] + E[3 + k] * res[1];
      v[(3 + k) * VLD + tid] = e0 * e0 + e1 * e1;
      if (write_j) J2[(size_t)(9 + k) * L.n_obs + i] = make_double2(e0, e1);
    }
    if (write_j) r2[i] = make_double2(res[0], res[1]);
  }
  __syncthreads();
  if (q.chunk >= 0) {
    // chunk of a long track: this tile's share of the point sums; k_ba_giant_point_combine adds the chunks in order
    if (tid < 6) {
      double sum = 0.0;
      for (int j = 0; j < q.no; ++j) sum += v[tid * VLD + j];
      chunk_pt[(size_t)q.chunk * 6 + tid] = sum;
    }
  } else if (tid < q.np) {
    const int p = q.pb + tid;
    const int b = meta.pptr[tid], e = meta.pptr[tid + 1];
    double s6[6] = {0, 0, 0, 0, 0, 0};
    for (int j = b; j < e; ++j) {
#pragma unroll
      for (int k = 0; k < 6; ++k) s6[k] += v[k * VLD + j];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { grad[pbase + (size_t)p * 3 + k] = s6[k]; cnorm2[pbase + (size_t)p * 3 + k] = s6[3 + k]; }
  }
  __syncthreads();
  // camera part: planes 0..8 = gradient terms, planes 9..17 = squared column norms of the scaled Jacobian; one pass of
  // segment sums over the 18 planes
  if (active) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const double sc = csc_s[slot * 9 + k];
      const double f0 = F[k] * sc, f1 = F[9 + k] * sc;
      v[k * VLD + tid] = F[k] * res[0] + F[9 + k] * res[1];
      v[(9 + k) * VLD + tid] = f0 * f0 + f1 * f1;
      if (write_j) J2[(size_t)k * L.n_obs + i] = make_double2(f0, f1);
    }
  }
  __syncthreads();
  seg_reduce_planes<18>(q, meta, v, 9, seg_g, 9, seg_n, 9, 0);
}


// The built-in functor: SnavelyReprojectionError through the staged 6 + 6-wide dual evaluation of jet.cuh.
struct SnavelyBuiltin {
  static __device__ __forceinline__ bool residual(const double* cam, const double* pt, double ox, double oy, double* res) {
    snavely_residual(cam, pt, ox, oy, res); return true;
  }
  static __device__ __forceinline__ bool residual_jacobian(const double* cam, const double* pt, double ox, double oy, double* res,
                                                           double* F, double* E) {
    snavely_residual_jacobian(cam, pt, ox, oy, res, F, E); return true;
  }
};

}  // namespace
}  // namespace sk
