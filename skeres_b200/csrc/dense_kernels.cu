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
  double val = 0.0;
  for (int q = pair_ptr[g]; q < pair_ptr[g + 1]; ++q) {
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
    val += g1[0] * h0 + g1[1] * h1 + g1[2] * h2;
  }
  const size_t nc = (size_t)L.n_cams * 9;
  const size_t r = (size_t)pair_c1[g] * 9 + a, c = (size_t)pair_c2[g] * 9 + b;
  S[r * nc + c] = -val;
  S[c * nc + r] = -val;
}

// ---- blocked Cholesky (lower), both triangles kept consistent (upper = L^T) ------------------------
constexpr int NB = 32;

__global__ void k_chol_diag(int n, int k0, int kb, double* __restrict__ A, int* error_flag) {
  __shared__ double T[NB][NB + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (tx < kb && ty < kb) T[ty][tx] = A[(size_t)(k0 + ty) * n + k0 + tx];
  __syncthreads();
  for (int j = 0; j < kb; ++j) {
    if (tx == 0 && ty == 0) {
      const double d = T[j][j];
      if (!(d > 0.0)) { atomicOr(error_flag, 4); T[j][j] = 1.0; } else T[j][j] = sqrt(d);
    }
    __syncthreads();
    if (ty == 0 && tx > j && tx < kb) T[tx][j] = T[tx][j] / T[j][j];
    __syncthreads();
    if (tx > j && ty > j && tx < kb && ty < kb && ty >= tx) T[ty][tx] -= T[ty][j] * T[tx][j];
    __syncthreads();
  }
  if (tx < kb && ty < kb) {
    const double v = (ty >= tx) ? T[ty][tx] : T[tx][ty];      // lower = L, upper = L^T
    A[(size_t)(k0 + ty) * n + k0 + tx] = v;
  }
}

// L_ik = A_ik L_kk^-T for the rows below the diagonal block; also mirrors into the upper triangle.
__global__ void k_chol_panel(int n, int k0, int kb, double* __restrict__ A) {
  __shared__ double Lk[NB][NB + 1];
  __shared__ double X[NB][NB + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int r0 = k0 + kb + blockIdx.x * NB;
  if (tx < kb && ty < kb) Lk[ty][tx] = A[(size_t)(k0 + ty) * n + k0 + tx];
  const int row = r0 + ty;
  if (row < n && tx < kb) X[ty][tx] = A[(size_t)row * n + k0 + tx];
  __syncthreads();
  if (tx == 0 && row < n) {
    for (int j = 0; j < kb; ++j) {
      double s = X[ty][j];
      for (int t = 0; t < j; ++t) s -= X[ty][t] * Lk[j][t];
      X[ty][j] = s / Lk[j][j];
    }
  }
  __syncthreads();
  if (row < n && tx < kb) {
    const double v = X[ty][tx];
    A[(size_t)row * n + k0 + tx] = v;
    A[(size_t)(k0 + tx) * n + row] = v;
  }
}

// A_ij -= sum_t L_i,k0+t L_j,k0+t for the trailing lower tiles.
__global__ void k_chol_update(int n, int k0, int kb, double* __restrict__ A) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  __shared__ double Li[NB][NB + 1];
  __shared__ double Lj[NB][NB + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int base = k0 + kb;
  const int i = base + bi * NB + ty, jrow = base + bj * NB + ty;
  if (tx < kb) {
    Li[ty][tx] = (i < n) ? A[(size_t)i * n + k0 + tx] : 0.0;
    Lj[ty][tx] = (jrow < n) ? A[(size_t)jrow * n + k0 + tx] : 0.0;
  }
  __syncthreads();
  const int j = base + bj * NB + tx;
  if (i < n && j < n && j <= i) {
    double s = 0.0;
    for (int t = 0; t < kb; ++t) s += Li[ty][t] * Lj[tx][t];
    A[(size_t)i * n + j] -= s;
  }
}

// Forward (L y = b, using the mirrored upper triangle row-wise) and backward (L^T z = y) substitution.
__global__ void __launch_bounds__(1024) k_chol_solve(int n, const double* __restrict__ A, const double* __restrict__ rhs,
                                                     double* __restrict__ z) {
  __shared__ double piv;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < n; i += nthr) z[i] = rhs[i];
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    if (tid == 0) { piv = z[j] / A[(size_t)j * n + j]; z[j] = piv; }
    __syncthreads();
    const double yj = piv;
    const double* urow = A + (size_t)j * n;              // U[j][i] = L[i][j]
    for (int i = j + 1 + tid; i < n; i += nthr) z[i] -= urow[i] * yj;
    __syncthreads();
  }
  for (int j = n - 1; j >= 0; --j) {
    if (tid == 0) { piv = z[j] / A[(size_t)j * n + j]; z[j] = piv; }
    __syncthreads();
    const double zj = piv;
    const double* lrow = A + (size_t)j * n;              // L[j][i], i < j  == U[i][j]
    for (int i = tid; i < j; i += nthr) z[i] -= lrow[i] * zj;
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
  const dim3 tb(NB, NB);
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int kb = std::min(NB, n - k0);
    k_chol_diag<<<1, tb, 0, s>>>(n, k0, kb, S, error_flag); ++launches;
    const int rem = n - k0 - kb;
    if (rem > 0) {
      const int nt = cdiv(rem, NB);
      k_chol_panel<<<nt, tb, 0, s>>>(n, k0, kb, S); ++launches;
      k_chol_update<<<dim3(nt, nt), tb, 0, s>>>(n, k0, kb, S); ++launches;
    }
  }
  k_chol_solve<<<1, 1024, 0, s>>>(n, S, rhs, z); ++launches;
  check_launch("cholesky");
  return launches;
}

}  // namespace sk
