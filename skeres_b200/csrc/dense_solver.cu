// dense_solver.cu — DENSE_QR back end for small generic problems (CurveFitting.scala:119-122:
// ExponentialResidual blocks, two scalar parameter blocks, DENSE_QR), plus the explicit-Schur part
// of the bundle-adjustment back end (pair lists, assembly, Cholesky).
#include "dense_solver.cuh"

#include <algorithm>
#include <array>
#include <map>

#include "ba_solver.cuh"
#include "user_functor.cuh"

namespace sk {

namespace {
__global__ void k_gather_ptr_blocks(int n, const double* const* __restrict__ src, double* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = *src[i];
}
__global__ void k_scatter_ptr_blocks(int n, double* const* __restrict__ dst, const double* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) *dst[i] = x[i];
}
}  // namespace

DenseSolver::DenseSolver(const sk_solver_options& opt, cudaStream_t stream, const std::vector<DenseRb>& rbs, int num_rows,
                         const std::vector<double*>& scalar_ptrs, int num_param_blocks)
    : LmSolver(opt, stream), nrb_((int)rbs.size()), m_(num_rows), nparam_blocks_(num_param_blocks) {
  SK_REQUIRE(comm_ == nullptr || comm_->world == 1, SK_ERR_UNSUPPORTED, "DENSE_QR runs on one GPU");
  const int64_t n = (int64_t)scalar_ptrs.size();
  SK_REQUIRE(n >= 1 && n <= 4096, SK_ERR_UNSUPPORTED, "DENSE_QR on the device supports 1..4096 parameters (got %lld)", (long long)n);
  SK_REQUIRE((double)(m_ + n) * (double)(n + 1) < 2.5e8, SK_ERR_UNSUPPORTED, "dense problem too large for DENSE_QR (%d x %lld)", m_, (long long)n);
  allocate(n, n);
  for (const DenseRb& rb : rbs) if (rb.functor >= kUserFunctorBase) user_functors_.push_back(rb.functor);
  std::sort(user_functors_.begin(), user_functors_.end());
  user_functors_.erase(std::unique(user_functors_.begin(), user_functors_.end()), user_functors_.end());
  d_rbs_.upload(rbs, stream_);
  d_ptrs_.upload(scalar_ptrs, stream_);
  J_.alloc((size_t)m_ * n); J_.zero(stream_);
  b_.alloc(m_); W_.alloc((size_t)(m_ + n) * (n + 1));
  block_cost_.alloc(cdiv(nrb_, 128)); mcc_part_.alloc(kMaxPartials);
  SK_CUDA(cudaStreamSynchronize(stream_));
}

void DenseSolver::load_state() {
  KScope k(prof_, SK_KF_LM);
  k_gather_ptr_blocks<<<cdiv(n_, 256), 256, 0, stream_>>>((int)n_, d_ptrs_.p, x_.p); check_launch("k_gather_ptr_blocks");
}
void DenseSolver::store_state() {
  KScope k(prof_, SK_KF_LM);
  k_scatter_ptr_blocks<<<cdiv(n_, 256), 256, 0, stream_>>>((int)n_, d_ptrs_.p, x_.p); check_launch("k_scatter_ptr_blocks");
}
void DenseSolver::fill_summary(sk_solver_summary_data* d) {
  d->num_parameter_blocks = nparam_blocks_; d->num_parameters = n_; d->num_residual_blocks = nrb_; d->num_residuals = m_;
}
ReduceJob DenseSolver::cost_job() { return {block_cost_.p, cdiv(nrb_, 128), SB_COST, 0}; }

// The built-in functors in one launch (it writes every cost slot), then one launch per run-time compiled functor, in id order,
// each adding its residual blocks' cost to the slots: a fixed order, hence reproducible.
void DenseSolver::evaluate(const double* xv, bool with_jacobian, const int* guard) {
  launch_dense_evaluate(nrb_, d_rbs_.p, xv, with_jacobian, with_jacobian ? J_.p : nullptr, m_, with_jacobian ? b_.p : nullptr, block_cost_.p,
                        &st_.p->eval_failed, guard, stream_);
  for (int id : user_functors_)
    launch_user_dense_evaluate(id, with_jacobian, nrb_, d_rbs_.p, xv, with_jacobian ? J_.p : nullptr, m_, with_jacobian ? b_.p : nullptr,
                               block_cost_.p, &st_.p->eval_failed, guard, stream_);
}

void DenseSolver::eval_jacobian(bool scale_valid, bool /*store*/, const int* guard) {
  KScope k(prof_, SK_KF_EVALUATE_JACOBIAN, 3 + (int)user_functors_.size());
  evaluate(x_.p, true, guard);
  launch_dense_gradient(m_, (int)n_, J_.p, b_.p, grad_.p, guard, stream_);
  launch_dense_scale_norms(m_, (int)n_, J_.p, scale_valid ? scale_.p : nullptr, cnorm2_.p, guard, stream_);
}
void DenseSolver::eval_cost(const double* xv, const int* guard) {
  KScope k(prof_, SK_KF_EVALUATE_COST, 1 + (int)user_functors_.size());
  evaluate(xv, false, guard);
}
ReduceJob DenseSolver::linear_solve(const PcgDev** pcg_out) {
  *pcg_out = nullptr;
  KScope k(prof_, SK_KF_DENSE, 3);
  int nparts = 0;
  launch_dense_qr_solve(m_, (int)n_, J_.p, b_.p, D_.p, W_.p, step_.p, mcc_part_.p, &nparts, stream_);
  // DenseQRSolver always reports success with one iteration
  LmDev* st = st_.p;
  const int vals[2] = {1, LIN_SUCCESS};
  SK_CUDA(cudaMemcpyAsync(&st->lin_iterations, vals, sizeof(vals), cudaMemcpyHostToDevice, stream_));
  return {mcc_part_.p, nparts, SB_MCC, 0};
}

// ---- explicit Schur complement for the BA back end --------------------------------------------------
void BaSolver::build_pair_lists() {
  const BaLayoutHost& H = H_;
  std::map<std::pair<int, int>, std::vector<std::array<int, 3>>> groups;
  for (int p = 0; p < H.n_pts; ++p)
    for (int i = H.pt_ptr[p]; i < H.pt_ptr[p + 1]; ++i)
      for (int j = i + 1; j < H.pt_ptr[p + 1]; ++j)
        groups[{H.obs_cam[i], H.obs_cam[j]}].push_back({i, j, p});    // cameras ascending inside a point
  std::vector<int> ptr{0}, c1, c2, o1, o2, pt;
  for (auto& g : groups) {
    c1.push_back(g.first.first); c2.push_back(g.first.second);
    for (auto& t : g.second) { o1.push_back(t[0]); o2.push_back(t[1]); pt.push_back(t[2]); }
    ptr.push_back((int)o1.size());
  }
  n_pair_groups_ = (int)c1.size();
  if (o1.empty()) { o1.push_back(0); o2.push_back(0); pt.push_back(0); }
  if (c1.empty()) { c1.push_back(0); c2.push_back(0); }
  pair_ptr_.upload(ptr, stream_); pair_c1_.upload(c1, stream_); pair_c2_.upload(c2, stream_);
  pair_o1_.upload(o1, stream_); pair_o2_.upload(o2, stream_); pair_pt_.upload(pt, stream_);
  SK_CUDA(cudaStreamSynchronize(stream_));
}

void BaSolver::explicit_schur_solve() {
  KScope k(prof_, SK_KF_DENSE, 0);
  launch_schur_diag(L_, M45_.p, D_.p, S_.p, stream_);
  launch_schur_offdiag(L_, n_pair_groups_, pair_ptr_.p, pair_c1_.p, pair_c2_.p, pair_o1_.p, pair_o2_.p, pair_pt_.p,
                       reinterpret_cast<const double2*>(J2_.p), einv_.p, S_.p, stream_);
  const int launches = launch_cholesky_solve((int)nc_, S_.p, rhs_.p, px_.p, &st_.p->lin_error, stream_);
  prof_.launches[SK_KF_DENSE] += 2 + launches;
  const int vals[2] = {1, LIN_SUCCESS};   // SchurComplementSolver: num_iterations = 1
  SK_CUDA(cudaMemcpyAsync(&st_.p->lin_iterations, vals, sizeof(vals), cudaMemcpyHostToDevice, stream_));
}

}  // namespace sk
