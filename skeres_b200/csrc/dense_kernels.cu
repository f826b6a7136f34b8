// dense_kernels.cu — see dense_kernels.cuh.  Restates: the generic
// evaluator (A.2), DenseQRSolver (A.8: unpivoted Householder QR of [J; D] with Eigen's reflector
// convention), the off-diagonal part of SchurEliminator::Eliminate (A.5) and Eigen LLT.
#include "dense_kernels.cuh"
#include "user_functor.cuh"

#include "lm_kernels.cuh"

namespace sk {

namespace {

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// blockDim.x <= 1024; returns the sum in every thread.
__device__ double block_sum_all(double x, double* red) {
  x = warp_sum(x);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = x;
  __syncthreads();
  double r = (l < nw) ? red[l] : 0.0;
  r = warp_sum(r);
  return r;
}

template <bool JAC>
__global__ void k_dense_evaluate(int nrb, const DenseRb* __restrict__ rbs, const double* __restrict__ x, double* __restrict__ J,
                                 int m, double* __restrict__ b, double* __restrict__ block_cost, int* fail_flag, const int* guard) {
  if (guard != nullptr && *guard == 0) return;
  __shared__ double red[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double cost = 0.0;
  if (i < nrb && rbs[i].functor < kUserFunctorBase) {     // run-time compiled functors: their own kernels (user_functor.cu)
    const DenseRb rb = rbs[i];
    FunctorInfo fi;
    functor_info(rb.functor, &fi);
    double xx[SK_MAX_TOTAL_PARAMS], res[SK_MAX_RESIDUALS], jac[SK_MAX_RESIDUALS * SK_MAX_TOTAL_PARAMS];
    int t = 0;
    for (int k = 0; k < fi.nblk; ++k)
      for (int c = 0; c < fi.sizes[k]; ++c) xx[t++] = x[rb.col[k] + c];
    const bool ok = evaluate_functor(rb.functor, rb.consts, xx, res, JAC ? jac : nullptr);
    if (!ok) atomicOr(fail_flag, 1);
    double sq = 0.0;
    for (int q = 0; q < fi.nres; ++q) sq += res[q] * res[q];
    double rho[3];
    LossSpec ls{rb.loss_type, rb.loss_a, rb.loss_b};
    loss_evaluate(ls, sq, rho);
    cost = 0.5 * rho[0];
    if (!(cost == cost)) atomicOr(fail_flag, 1);
    if (JAC) {
      const Corrector corr(sq, rho);
      corr.correct_jacobian(fi.nres, fi.ntot, fi.ntot, res, jac);
      corr.correct_residuals(fi.nres, res);
      t = 0;
      for (int k = 0; k < fi.nblk; ++k) {
        for (int c = 0; c < fi.sizes[k]; ++c) {
          for (int q = 0; q < fi.nres; ++q) J[(size_t)(rb.col[k] + c) * m + rb.row + q] = jac[q * fi.ntot + t + c];
        }
        t += fi.sizes[k];
      }
      for (int q = 0; q < fi.nres; ++q) b[rb.row + q] = res[q];
    }
  }
  const double tot = block_sum_all(cost, red);
  if (threadIdx.x == 0) block_cost[blockIdx.x] = tot;
}

__global__ void k_dense_gradient(int m, int n, const double* __restrict__ J, const double* __restrict__ b, double* __restrict__ g,
                                 const int* guard) {
  if (guard != nullptr && *guard == 0) return;
  __shared__ double red[32];
  const int j = blockIdx.x;
  const double* cj = J + (size_t)j * m;
  double a = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) a += cj[i] * b[i];
  a = block_sum_all(a, red);
  if (threadIdx.x == 0) g[j] = a;
}

__global__ void k_dense_scale_norms(int m, int n, double* __restrict__ J, const double* __restrict__ scale,
                                    double* __restrict__ cnorm2, const int* guard) {
  if (guard != nullptr && *guard == 0) return;
  __shared__ double red[32];
  const int j = blockIdx.x;
  double* cj = J + (size_t)j * m;
  const double sc = scale ? scale[j] : 1.0;
  double a = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    double v = cj[i];
    if (scale) { v *= sc; cj[i] = v; }
    a += v * v;
  }
  a = block_sum_all(a, red);
  if (threadIdx.x == 0) cnorm2[j] = a;
}

// Single-CTA Householder QR least squares. W is (m+n) x (n+1) column-major; the last column is the rhs.
__global__ void __launch_bounds__(1024) k_dense_qr(int m, int n, const double* __restrict__ J, const double* __restrict__ b,
                                                   const double* __restrict__ D, double* __restrict__ W, double* __restrict__ step) {
  __shared__ double red[32];
  __shared__ double sh_tau, sh_div;
  const int M = m + n;
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
  for (size_t idx = tid; idx < (size_t)M * (n + 1); idx += nthr) {
    const int col = (int)(idx / M), row = (int)(idx - (size_t)col * M);
    double v;
    if (col < n) v = (row < m) ? J[(size_t)col * m + row] : ((row - m) == col ? D[col] : 0.0);
    else v = (row < m) ? b[row] : 0.0;
    W[idx] = v;
  }
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    double* ck = W + (size_t)k * M;
    double a = 0.0;
    for (int i = k + 1 + tid; i < M; i += nthr) a += ck[i] * ck[i];
    const double tail = block_sum_all(a, red);
    if (tid == 0) {
      const double c0 = ck[k];
      if (tail <= 2.2250738585072014e-308) { sh_tau = 0.0; sh_div = 0.0; }
      else {
        double beta = sqrt(c0 * c0 + tail);
        if (c0 >= 0.0) beta = -beta;
        sh_div = c0 - beta;
        sh_tau = (beta - c0) / beta;
        ck[k] = beta;
      }
    }
    __syncthreads();
    const double tau = sh_tau, dv = sh_div;
    if (dv != 0.0) { for (int i = k + 1 + tid; i < M; i += nthr) ck[i] = ck[i] / dv; }
    else { for (int i = k + 1 + tid; i < M; i += nthr) ck[i] = 0.0; }
    __syncthreads();
    for (int j = k + 1 + warp; j <= n; j += nwarps) {      // remaining columns and the rhs, one warp each
      double* cj = W + (size_t)j * M;
      double s = 0.0;
      for (int i = k + 1 + lane; i < M; i += 32) s += ck[i] * cj[i];
      s = warp_sum(s);
      s += cj[k];
      const double ts = tau * s;
      for (int i = k + 1 + lane; i < M; i += 32) cj[i] -= ts * ck[i];
      __syncwarp();
      if (lane == 0) cj[k] -= ts;
    }
    __syncthreads();
  }
  if (tid == 0) {                                          // R x = (Q^T rhs)[0:n]; LM steps by -x
    const double* rhs = W + (size_t)n * M;
    for (int i = n - 1; i >= 0; --i) {
      double s = rhs[i];
      for (int j = i + 1; j < n; ++j) s -= W[(size_t)j * M + i] * (-step[j]);
      step[i] = -(s / W[(size_t)i * M + i]);
    }
  }
}

// part[block] = sum over rows of m_i (b_i + m_i / 2), m = J * step
__global__ void k_dense_model(int m, int n, const double* __restrict__ J, const double* __restrict__ b, const double* __restrict__ step,
                              double* __restrict__ part) {
  __shared__ double red[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double mc = 0.0;
  if (i < m) {
    double mi = 0.0;
    for (int j = 0; j < n; ++j) mi += J[(size_t)j * m + i] * step[j];
    mc = mi * (b[i] + mi / 2.0);
  }
  mc = block_sum_all(mc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = mc;
}

// ---- explicit Schur ------------------------------------------------------------------------------
__global__ void k_schur_diag(BaDev L, const double* __restrict__ M45, const double* __restrict__ D, double* __restrict__ S) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L.n_cams * 81) return;
  const int c = idx / 81, e = idx - c * 81, a = e / 9, b = e - a * 9;
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  const int packed = lo * 9 - lo * (lo - 1) / 2 + (hi - lo);   // upper-triangle row-major index
  double v = M45[(size_t)c * 45 + packed];
  if (a == b) { const double d = D[(size_t)c * 9 + a]; v += d * d; }
  const size_t nc = (size_t)L.n_cams * 9;
  S[((size_t)c * 9 + a) * nc + (size_t)c * 9 + b] = v;
}

__global__ void k_schur_offdiag(BaDev L, const int* __restrict__ pair_ptr, const int* __restrict__ pair_c1,
                                const int* __restrict__ pair_c2, const int* __restrict__ pair_o1, const int* __restrict__ pair_o2,
                                const int* __restrict__ pair_pt, const double2* __restrict__ J2, const double* __restrict__ einv,
                                double* __restrict__ S) {
  const int g = blockIdx.x, t = threadIdx.x;
  if (t >= 81) return;
  const int a = t / 9, b = t - a * 9;
  const size_t O = (size_t)L.n_obs;
  // entry (a, b) of G_o1^T (E^T E)^-1 G_o2 for one shared point; four points in flight at a time (the loads of a point hang off
  // its three indices: one after the other they cost two global round trips per point), added in list order as before
  auto term = [&](int q) {
    const int o1 = pair_o1[q], o2 = pair_o2[q];
    const double* m = einv + (size_t)pair_pt[q] * 6;
    const double2 fa = J2[a * O + o1], fb = J2[b * O + o2];
    double g1[3], g2[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const double2 e1 = J2[(9 + u) * O + o1], e2 = J2[(9 + u) * O + o2];
      g1[u] = e1.x * fa.x + e1.y * fa.y;
      g2[u] = e2.x * fb.x + e2.y * fb.y;
    }
    const double h0 = m[0] * g2[0] + m[1] * g2[1] + m[2] * g2[2];
    const double h1 = m[1] * g2[0] + m[3] * g2[1] + m[4] * g2[2];
    const double h2 = m[2] * g2[0] + m[4] * g2[1] + m[5] * g2[2];
    return g1[0] * h0 + g1[1] * h1 + g1[2] * h2;
  };
  double val = 0.0;
  int q = pair_ptr[g];
  const int qe = pair_ptr[g + 1];
  for (; q + 4 <= qe; q += 4) {
    double tq[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) tq[u] = term(q + u);
#pragma unroll
    for (int u = 0; u < 4; ++u) val += tq[u];
  }
  for (; q < qe; ++q) val += term(q);
  const size_t nc = (size_t)L.n_cams * 9;
  const size_t r = (size_t)pair_c1[g] * 9 + a, c = (size_t)pair_c2[g] * 9 + b;
  S[r * nc + c] = -val;
  S[c * nc + r] = -val;
}

// ---- blocked Cholesky (lower), both triangles kept consistent (upper = L^T) ------------------------
constexpr int NB = 32;

// In-place lower Cholesky of the kb x kb block T (shared memory, row stride NB + 1) by ONE warp, left-looking: column j is
// L_ij = (A_ij - sum_{k<j} L_ik L_jk) / L_jj with lane i = row i (L_jk is a broadcast read, L_ik conflict-free at the odd stride).
// No CTA barrier inside (round 1 factored the block with 1024 threads and three barriers per column).  A register-resident
// right-looking variant (row per lane, L_kj by shuffle, fully unrolled) measured SLOWER: 56.6 against 35.0 us per panel kernel
// (profiles/r02_v3_dense_path.md).  A pivot that is not > 0 raises bit 4 of *error_flag.
__device__ __forceinline__ void chol_block_warp(double (*T)[NB + 1], int kb, int* error_flag, bool report) {
  const int i = threadIdx.x & 31;
  for (int j = 0; j < kb; ++j) {
    double s = (i >= j && i < kb) ? T[i][j] : 0.0;
    for (int k = 0; k < j; ++k) s -= ((i >= j && i < kb) ? T[i][k] : 0.0) * T[j][k];
    double d = __shfl_sync(0xffffffffu, s, j);
    if (!(d > 0.0)) { if (report && i == 0) atomicOr(error_flag, 4); d = 1.0; }
    const double r = sqrt(d);
    if (i == j) T[i][j] = r;
    else if (i > j && i < kb) T[i][j] = s / r;
    __syncwarp();
  }
}

// Diagonal block of panel k0 (factored redundantly by every CTA: no launch and no trip through memory in between; CTA 0
// writes it back, lower = L and upper = L^T) and L_ik = A_ik L_kk^-T for the CTA's 32 rows below it, mirrored into the upper
// triangle.  blockDim = (32, 32): the 32 threads that share ty are one warp and own one row of the panel; the row's triangular
// solve runs j = 0 .. kb-1 with the dot product over t < j spread over the warp's lanes (fixed shuffle tree).
// gridDim.x = max(1, row blocks below the panel).
__global__ void __launch_bounds__(1024) k_chol_panel(int n, int k0, int kb, double* __restrict__ A, int* error_flag) {
  __shared__ double Lk[NB][NB + 1];
  __shared__ double X[NB][NB + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int r0 = k0 + kb + blockIdx.x * NB;
  Lk[ty][tx] = (tx < kb && ty < kb) ? A[(size_t)(k0 + ty) * n + k0 + tx] : 0.0;
  const int row = r0 + ty;
  X[ty][tx] = (row < n && tx < kb) ? A[(size_t)row * n + k0 + tx] : 0.0;
  __syncthreads();
  if (ty == 0) chol_block_warp(Lk, kb, error_flag, blockIdx.x == 0);
  __syncthreads();
  if (blockIdx.x == 0 && tx < kb && ty < kb) A[(size_t)(k0 + ty) * n + k0 + tx] = (ty >= tx) ? Lk[ty][tx] : Lk[tx][ty];
  if (row < n) {
    double xj_mine = X[ty][tx];                            // lane tx keeps X[ty][tx]; finished entries are final
    for (int j = 0; j < kb; ++j) {
      double part = (tx < j) ? xj_mine * Lk[j][tx] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (tx == j) xj_mine = (xj_mine - part) / Lk[j][j];
    }
    if (tx < kb) {
      A[(size_t)row * n + k0 + tx] = xj_mine;
      A[(size_t)(k0 + tx) * n + row] = xj_mine;
    }
  }
}

// ---- FP64 tensor-core tile product -------------------------------------------------------------------------------------------
// D(8x8) += A(8x4) * B(4x8): mma.sync.aligned.m8n8k4.row.col.f64 (DMMA).  Lane l supplies a = A[l / 4][l % 4] and
// b = B[l % 4][l / 4] and holds d0 = D[l / 4][2 (l % 4)], d1 = D[l / 4][2 (l % 4) + 1].
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// A_ij -= sum_t L_i,k0+t L_j,k0+t for the trailing lower tiles: the one real dense contraction of the hot path (north_star (3)),
// (n - k0)^2 / 2 * kb multiply-adds per panel, on the FP64 tensor pipe.  One CTA of 16 warps per 32 x 32 tile of the trailing
// matrix; warp (wi, wj) owns the 8 x 8 sub-tile and walks the panel's kb columns four at a time.
__global__ void __launch_bounds__(512) k_chol_update(int n, int k0, int kb, double* __restrict__ A) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  __shared__ double Li[NB][NB + 4];                        // +4: the (row, 4 consecutive k) fragment loads hit distinct banks
  __shared__ double Lj[NB][NB + 4];
  const int tid = threadIdx.x;
  const int base = k0 + kb;
  for (int e = tid; e < NB * NB; e += 512) {
    const int r = e / NB, c = e - r * NB;
    const int i = base + bi * NB + r, j = base + bj * NB + r;
    Li[r][c] = (c < kb && i < n) ? A[(size_t)i * n + k0 + c] : 0.0;
    Lj[r][c] = (c < kb && j < n) ? A[(size_t)j * n + k0 + c] : 0.0;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  const int wi = warp >> 2, wj = warp & 3;
  const int fr = lane >> 2, fk = lane & 3;
  double d0 = 0.0, d1 = 0.0;
#pragma unroll
  for (int ks = 0; ks < NB; ks += 4) dmma_8x8x4(d0, d1, Li[wi * 8 + fr][ks + fk], Lj[wj * 8 + fr][ks + fk]);
  const int i = base + bi * NB + wi * 8 + fr;
  const int j = base + bj * NB + wj * 8 + 2 * fk;
  if (i < n) {
    if (j < n && j <= i) A[(size_t)i * n + j] -= d0;
    if (j + 1 < n && j + 1 <= i) A[(size_t)i * n + j + 1] -= d1;
  }
}

// Forward (L y = b) and backward (L^T z = y) substitution in blocks of 32 unknowns: the diagonal block is solved by one warp out
// of shared memory (lane i keeps unknown i, one shuffle per step), then all threads subtract the block's contribution from the
// unknowns still open (coalesced rows of the mirrored triangle).  Two CTA barriers per 32 unknowns (round 1: two per unknown).
__global__ void __launch_bounds__(1024) k_chol_solve(int n, const double* __restrict__ A, const double* __restrict__ rhs,
                                                     double* __restrict__ z) {
  __shared__ double Dg[NB][NB + 1];
  __shared__ double zb[NB];
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < n; i += nthr) z[i] = rhs[i];
  __syncthreads();
  // forward: L y = b
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int kb = min(NB, n - k0);
    for (int e = tid; e < NB * NB; e += nthr) { const int r = e / NB, c = e - r * NB; Dg[r][c] = (r < kb && c < kb) ? A[(size_t)(k0 + r) * n + k0 + c] : 0.0; }
    __syncthreads();
    if (tid < 32) {
      double zi = (tid < kb) ? z[k0 + tid] : 0.0;
      for (int j = 0; j < kb; ++j) {
        const double zj = __shfl_sync(0xffffffffu, zi, j) / Dg[j][j];
        if (tid == j) zi = zj;
        else if (tid > j && tid < kb) zi -= Dg[tid][j] * zj;               // L[i][j], i > j
      }
      if (tid < kb) { z[k0 + tid] = zi; zb[tid] = zi; }
    }
    __syncthreads();
    for (int i = k0 + kb + tid; i < n; i += nthr) {
      double s = 0.0;
      for (int j = 0; j < kb; ++j) s += A[(size_t)(k0 + j) * n + i] * zb[j];  // U[k0 + j][i] = L[i][k0 + j]: coalesced over i
      z[i] -= s;
    }
    __syncthreads();
  }
  // backward: L^T z = y
  for (int k1 = n; k1 > 0; k1 -= NB) {
    const int k0 = max(0, k1 - NB), kb = k1 - k0;
    for (int e = tid; e < NB * NB; e += nthr) { const int r = e / NB, c = e - r * NB; Dg[r][c] = (r < kb && c < kb) ? A[(size_t)(k0 + r) * n + k0 + c] : 0.0; }
    __syncthreads();
    if (tid < 32) {
      double zi = (tid < kb) ? z[k0 + tid] : 0.0;
      for (int j = kb - 1; j >= 0; --j) {
        const double zj = __shfl_sync(0xffffffffu, zi, j) / Dg[j][j];
        if (tid == j) zi = zj;
        else if (tid < j) zi -= Dg[j][tid] * zj;                            // L^T[i][j] = L[j][i], i < j
      }
      if (tid < kb) { z[k0 + tid] = zi; zb[tid] = zi; }
    }
    __syncthreads();
    for (int i = tid; i < k0; i += nthr) {
      double s = 0.0;
      for (int j = 0; j < kb; ++j) s += A[(size_t)(k0 + j) * n + i] * zb[j];  // L[k0 + j][i], i < k0: coalesced over i
      z[i] -= s;
    }
    __syncthreads();
  }
}

}  // namespace

void launch_dense_evaluate(int nrb, const DenseRb* rbs, const double* x, bool with_jacobian, double* J, int m, double* b,
                           double* block_cost, int* fail_flag, const int* guard, cudaStream_t s) {
  const int blocks = cdiv(nrb, 128);
  if (with_jacobian) k_dense_evaluate<true><<<blocks, 128, 0, s>>>(nrb, rbs, x, J, m, b, block_cost, fail_flag, guard);
  else k_dense_evaluate<false><<<blocks, 128, 0, s>>>(nrb, rbs, x, J, m, b, block_cost, fail_flag, guard);
  check_launch("k_dense_evaluate");
}
void launch_dense_gradient(int m, int n, const double* J, const double* b, double* g, const int* guard, cudaStream_t s) {
  k_dense_gradient<<<n, 256, 0, s>>>(m, n, J, b, g, guard); check_launch("k_dense_gradient");
}
void launch_dense_scale_norms(int m, int n, double* J, const double* scale, double* cnorm2, const int* guard, cudaStream_t s) {
  k_dense_scale_norms<<<n, 256, 0, s>>>(m, n, J, scale, cnorm2, guard); check_launch("k_dense_scale_norms");
}
void launch_dense_qr_solve(int m, int n, const double* J, const double* b, const double* D, double* W, double* step,
                           double* mcc_part, int* nparts_out_host, cudaStream_t s) {
  k_dense_qr<<<1, 1024, 0, s>>>(m, n, J, b, D, W, step); check_launch("k_dense_qr");
  const int blocks = cdiv(m, 256);
  SK_REQUIRE(blocks <= kMaxPartials, SK_ERR_UNSUPPORTED, "dense problem with more than %d residuals", kMaxPartials * 256);
  k_dense_model<<<blocks, 256, 0, s>>>(m, n, J, b, step, mcc_part); check_launch("k_dense_model");
  *nparts_out_host = blocks;
}

void launch_schur_diag(const BaDev& L, const double* M45, const double* D, double* S, cudaStream_t s) {
  const size_t nc = (size_t)L.n_cams * 9;
  SK_CUDA(cudaMemsetAsync(S, 0, nc * nc * sizeof(double), s));
  k_schur_diag<<<cdiv((int64_t)L.n_cams * 81, 256), 256, 0, s>>>(L, M45, D, S); check_launch("k_schur_diag");
}
void launch_schur_offdiag(const BaDev& L, int n_groups, const int* pair_ptr, const int* pair_c1, const int* pair_c2, const int* pair_o1,
                          const int* pair_o2, const int* pair_pt, const double2* J2, const double* einv, double* S, cudaStream_t s) {
  if (n_groups == 0) return;
  k_schur_offdiag<<<n_groups, 96, 0, s>>>(L, pair_ptr, pair_c1, pair_c2, pair_o1, pair_o2, pair_pt, J2, einv, S);
  check_launch("k_schur_offdiag");
}
int launch_cholesky_solve(int n, double* S, const double* rhs, double* z, int* error_flag, cudaStream_t s) {
  int launches = 0;
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int kb = std::min(NB, n - k0);
    const int rem = n - k0 - kb;
    const int nt = cdiv(rem, NB);
    k_chol_panel<<<std::max(nt, 1), dim3(NB, NB), 0, s>>>(n, k0, kb, S, error_flag); ++launches;       // diagonal block + the rows below it
    if (rem > 0) { k_chol_update<<<dim3(nt, nt), 512, 0, s>>>(n, k0, kb, S); ++launches; }   // trailing update on the FP64 tensor pipe
  }
  k_chol_solve<<<1, 1024, 0, s>>>(n, S, rhs, z); ++launches;
  check_launch("cholesky");
  return launches;
}

}  // namespace sk
