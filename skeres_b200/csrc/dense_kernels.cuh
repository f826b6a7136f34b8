// dense_kernels.cuh — small dense problems (DENSE_QR) and the dense reduced camera system
// (DENSE_SCHUR / SPARSE_SCHUR): generic-functor evaluation into a dense Jacobian, Householder QR,
// explicit Schur complement assembly, blocked Cholesky.
#pragma once
#include "ba_kernels.cuh"
#include "common.cuh"
#include "eval_abi.cuh"
#include "jet.cuh"

namespace sk {

// Residuals + (unscaled) dense Jacobian J (m x n column-major, ld = m); b = corrected residuals.
// Entries of J not covered by a residual block must already be zero (structure is fixed).
void launch_dense_evaluate(int nrb, const DenseRb* rbs, const double* x, bool with_jacobian, double* J, int m,
                           double* b, double* block_cost, int* fail_flag, const int* guard, cudaStream_t s);
// g[j] = sum_i J[i,j] b[i];  one CTA per column.
void launch_dense_gradient(int m, int n, const double* J, const double* b, double* g, const int* guard, cudaStream_t s);
// optional in-place column scaling J[:,j] *= scale[j], then cnorm2[j] = ||J[:,j]||^2
void launch_dense_scale_norms(int m, int n, double* J, const double* scale, double* cnorm2, const int* guard, cudaStream_t s);
// x = argmin || [J; diag(D)] x - [b; 0] ||  by unpivoted Householder QR (DenseQRSolver, A.8); then
// step = -x and part[i-block] = partial sums of m_i (b_i + m_i/2), m = J step. W: (m+n) x (n+1) workspace.
void launch_dense_qr_solve(int m, int n, const double* J, const double* b, const double* D, double* W, double* step,
                           double* mcc_part /*[kMaxPartials]*/, int* nparts_out_host, cudaStream_t s);

// ---- explicit Schur complement -------------------------------------------------------------------
// S (nc x nc, row-major, both triangles) = blockdiag(M45 + D_c^2) ; off-diagonal blocks added by pairs.
void launch_schur_diag(const BaDev& L, const double* M45, const double* D, double* S, cudaStream_t s);
// For every camera pair group g: S[c1,c2] -= sum_pairs G_o1^T (E^T E)^-1 G_o2  (and the transpose block).
void launch_schur_offdiag(const BaDev& L, int n_groups, const int* pair_ptr, const int* pair_c1, const int* pair_c2,
                          const int* pair_o1, const int* pair_o2, const int* pair_pt, const double2* J2,
                          const double* einv, double* S, cudaStream_t s);
// In-place lower Cholesky of S (n x n row-major, symmetric, both triangles given) and solve S z = rhs.
// error_flag bit 4 is set when a pivot is not positive. Returns the number of kernels launched.
int launch_cholesky_solve(int n, double* S, const double* rhs, double* z, int* error_flag, cudaStream_t s);

}  // namespace sk
