// batch.cu — BASELINE.json configs[3]: many independent CurveFitting-shaped problems
// (CurveFitting.scala:92-122: residual y - exp(m x + c), parameter blocks m and c, DENSE_QR,
// trust-region LM), one thread per problem, ONE kernel launch per LM iteration.
//
// Each thread runs exactly the iteration the single-problem path runs (TrustRegionMinimizer A.3,
// LevenbergMarquardtStrategy A.4, Householder QR of the 2-column [J; D] A.8), with the problem's LM
// state kept in SoA device arrays between launches.
//
// Two passes over a problem's observations per LM iteration.  The sums a Jacobian evaluation at the current point feeds into
// the iteration (cost, gradient, column norms, the first Householder reflector's dot products: `Sums`) are part of the saved
// state: they are produced by the pass that evaluates the CANDIDATE point (which an accepted step turns into the current point;
// TrustRegionMinimizer evaluates cost and Jacobian there anyway) and stay valid after a rejected step.  An iteration therefore
// reads the observations once for the second Householder reflector and once for the candidate -- 134 exponentials -- where the
// first version of this file read them four times (201 exponentials, 69 KB of shared memory per CTA for the stashed
// exponentials: 12 resident warps per SM).  Same arithmetic in the same order for every quantity that is used.
#include "batch.cuh"

#include <cfloat>
#include <chrono>

#include "jet.cuh"
#include "lm_kernels.cuh"

namespace sk {

namespace {

constexpr int BT = 128;            // threads (= problems) per CTA

struct BatchState {                // SoA, index = problem
  double *radius, *decrease, *x_cost, *x_norm, *scale1, *scale2, *diag1, *diag2, *initial_cost, *final_cost, *grad_max;
  double* sums;                    // [kSums][np]: the Sums of the Jacobian evaluation at the current point
  int *iteration, *reuse, *nci, *done, *term, *nsucc, *nunsucc;
};

struct Sums { double cost, g1, g2, n1, n2, a0, b0, r0, S11, S12, S1r, S22, S2r; };
constexpr int kSums = sizeof(Sums) / sizeof(double);
static_assert(kSums == 13, "Sums is saved field by field");
__device__ __forceinline__ void save_sums(double* base, int64_t np, int64_t p, const Sums& s) {
  const double v[kSums] = {s.cost, s.g1, s.g2, s.n1, s.n2, s.a0, s.b0, s.r0, s.S11, s.S12, s.S1r, s.S22, s.S2r};
#pragma unroll
  for (int k = 0; k < kSums; ++k) base[(int64_t)k * np + p] = v[k];
}
__device__ __forceinline__ Sums load_sums(const double* base, int64_t np, int64_t p) {
  double v[kSums];
#pragma unroll
  for (int k = 0; k < kSums; ++k) v[k] = base[(int64_t)k * np + p];
  return Sums{v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], v[12]};
}

// Pass over the observations at (m, c): stashes e_i = exp(m x_i + c) and accumulates everything the
// iteration needs from the (optionally column-scaled) Jacobian  J_i = (-x_i e_i, -e_i)  — the
// infinitesimal parts spire's Jet yields for ExponentialResidual (CurveFitting.scala:96).
// exp(m x + c) with the multiply-add spelled out: the passes must agree on every e_i bit for bit.
__device__ __forceinline__ double model(double m, double x, double c) { return exp(__fma_rn(m, x, c)); }

// Measured on a B200 (1M problems, profiles/r02_v12_batch_curve_fits.md): four rows per load batch, the next batch loaded ahead,
// four CTAs per SM (128 registers) 7.9 ms for the whole batch; without the load-ahead and five CTAs per SM (96 registers, 120
// bytes of spills) 8.6 ms; no cap (140 registers, three CTAs) 9.5 ms; one / two rows per batch 11.1 / 10.3 ms; the four-pass
// version with the shared-memory stash 24.1 ms.
#ifndef SK_BATCH_RB                // development only (tools/gpu/build_variants.sh): A/B of the load batching and of the register cap
#define SK_BATCH_RB 4
#endif
#ifndef SK_BATCH_MINB
#define SK_BATCH_MINB 4
#endif
constexpr int RB = SK_BATCH_RB;    // observation rows whose loads are issued together (the adds keep the row order)

#ifndef SK_BATCH_PIPE              // 1 = the loads of the next batch of rows are issued before the current batch is used (0: development A/B)
#define SK_BATCH_PIPE 1
#endif
// row(i, x_i, y_i) for i = first .. nobs - 1, in that order; the loads are issued RB rows at a time.  STREAM: last use of the
// lines (evict-first), else plain read-only loads (the same lines are read again by a later pass of this launch).
template <bool STREAM, class F>
__device__ __forceinline__ void for_rows(const double* __restrict__ x, const double* __restrict__ y, int64_t np, int64_t p, int first,
                                         int nobs, F&& row) {
  auto ld = [&](const double* a, int i) { return STREAM ? __ldcs(a + (int64_t)i * np + p) : __ldg(a + (int64_t)i * np + p); };
  int i = first;
#if SK_BATCH_PIPE
  double xn[RB], yn[RB];
  if (i + RB <= nobs) {
#pragma unroll
    for (int u = 0; u < RB; ++u) { xn[u] = ld(x, i + u); yn[u] = ld(y, i + u); }
  }
  for (; i + RB <= nobs; i += RB) {
    double xi[RB], yi[RB];
#pragma unroll
    for (int u = 0; u < RB; ++u) { xi[u] = xn[u]; yi[u] = yn[u]; }
    if (i + 2 * RB <= nobs) {
#pragma unroll
      for (int u = 0; u < RB; ++u) { xn[u] = ld(x, i + RB + u); yn[u] = ld(y, i + RB + u); }
    }
#pragma unroll
    for (int u = 0; u < RB; ++u) row(i + u, xi[u], yi[u]);
  }
#else
  for (; i + RB <= nobs; i += RB) {
    double xi[RB], yi[RB];
#pragma unroll
    for (int u = 0; u < RB; ++u) { xi[u] = ld(x, i + u); yi[u] = ld(y, i + u); }
#pragma unroll
    for (int u = 0; u < RB; ++u) row(i + u, xi[u], yi[u]);
  }
#endif
  for (; i < nobs; ++i) row(i, ld(x, i), ld(y, i));
}

// Pass over the observations at (m, c): accumulates everything an iteration needs from the (optionally column-scaled)
// Jacobian  J_i = (-x_i e_i, -e_i), e_i = exp(m x_i + c)  -- the infinitesimal parts spire's Jet yields for
// ExponentialResidual (CurveFitting.scala:96).
__device__ __forceinline__ void pass_jacobian(const double* __restrict__ x, const double* __restrict__ y, int64_t np, int64_t p,
                                              int nobs, double m, double c, double s1, double s2, Sums* o) {
  Sums s{};
  auto row = [&](int i, double xi, double yi) {
    const double e = model(m, xi, c);
    const double r = yi - e;
    const double j1 = -(e * xi), j2 = -e;
    s.cost += r * r;
    s.g1 += j1 * r; s.g2 += j2 * r;                 // gradient uses the unscaled Jacobian
    const double a = j1 * s1, b = j2 * s2;
    s.n1 += a * a; s.n2 += b * b;
    if (i == 0) { s.a0 = a; s.b0 = b; s.r0 = r; }
    else { s.S11 += a * a; s.S12 += a * b; s.S1r += a * r; }
    s.S22 += a * b; s.S2r += b * r;                 // full sums (including row 0) for the model cost
  };
  for_rows<true>(x, y, np, p, 0, nobs, row);
  s.cost *= 0.5;
  *o = s;
}

__device__ __forceinline__ bool finalize(BatchState st, int64_t p, bool successful, double grad_max, double row_cost,
                                         LmParams prm) {
  // FinalizeIterationAndCheckIfMinimizerCanContinue
  if (successful) st.nsucc[p] += 1; else st.nunsucc[p] += 1;
  st.final_cost[p] = fmin(st.final_cost[p], row_cost);
  int term = -1;
  if (st.iteration[p] >= prm.max_num_iterations) term = SK_NO_CONVERGENCE;
  else if (successful && grad_max <= prm.gradient_tolerance) term = SK_CONVERGENCE;
  else if (st.radius[p] <= prm.min_radius) term = SK_CONVERGENCE;
  if (term >= 0) { st.done[p] = 1; st.term[p] = term; return false; }
  return true;
}

// One CTA-wide count, one global atomic per CTA.
__device__ __forceinline__ void publish_active(bool active, int* active_count) {
  const int cnt = __syncthreads_count(active ? 1 : 0);
  if (threadIdx.x == 0 && cnt) atomicAdd(active_count, cnt);
}

__global__ void __launch_bounds__(BT) k_batch_init(int64_t np, int nobs, const double* __restrict__ x, const double* __restrict__ y,
                                                   const double* __restrict__ mc, BatchState st, LmParams prm, double radius0,
                                                   int jacobi_scaling, int* active_count) {
  const int64_t p = blockIdx.x * (int64_t)BT + threadIdx.x;
  // No thread may leave before publish_active: it contains a CTA-wide barrier, and a warp whose lanes reach
  // a barrier at two different program points (n_problems not a multiple of 32) is undefined behaviour.
  bool active = false;
  if (p < np) {
    const double m = mc[p], c = mc[np + p];
    Sums s;
    pass_jacobian(x, y, np, p, nobs, m, c, 1.0, 1.0, &s);
    const double s1 = jacobi_scaling ? 1.0 / (1.0 + sqrt(s.n1)) : 1.0, s2 = jacobi_scaling ? 1.0 / (1.0 + sqrt(s.n2)) : 1.0;
    st.scale1[p] = s1; st.scale2[p] = s2;
    if (jacobi_scaling) pass_jacobian(x, y, np, p, nobs, m, c, s1, s2, &s);   // the sums of the column-scaled Jacobian (cost and gradient: the same bits)
    save_sums(st.sums, np, p, s);
    st.radius[p] = radius0; st.decrease[p] = 2.0; st.reuse[p] = 0; st.nci[p] = 0;
    st.x_cost[p] = s.cost; st.x_norm[p] = sqrt(m * m + c * c);
    st.initial_cost[p] = s.cost; st.final_cost[p] = s.cost;
    st.iteration[p] = 0; st.done[p] = 0; st.term[p] = SK_NO_CONVERGENCE; st.nsucc[p] = 0; st.nunsucc[p] = 0;
    const double gm = fmax(fabs(m - (m + (-s.g1))), fabs(c - (c + (-s.g2))));
    st.grad_max[p] = gm;
    active = finalize(st, p, true, gm, s.cost, prm);
  }
  publish_active(active, active_count);
}

// One LM iteration of problem p; returns whether the problem is still active afterwards.
__device__ bool iterate_one(int64_t np, int nobs, const double* __restrict__ x, const double* __restrict__ y,
                            double* __restrict__ mc, BatchState st, LmParams prm, int64_t p) {
  const double m = mc[p], c = mc[np + p];
  const double s1 = st.scale1[p], s2 = st.scale2[p];
  double radius = st.radius[p];
  st.iteration[p] += 1;
  const Sums s = load_sums(st.sums, np, p);          // Jacobian evaluation at (m, c): done by the iteration that moved here
  const double x_cost = st.x_cost[p];
  // LevenbergMarquardtStrategy::ComputeStep
  double d1, d2;
  if (!st.reuse[p]) {
    d1 = fmin(fmax(s.n1, prm.min_lm_diagonal), prm.max_lm_diagonal);
    d2 = fmin(fmax(s.n2, prm.min_lm_diagonal), prm.max_lm_diagonal);
    st.diag1[p] = d1; st.diag2[p] = d2;
  } else { d1 = st.diag1[p]; d2 = st.diag2[p]; }
  const double D1 = sqrt(d1 / radius), D2 = sqrt(d2 / radius);
  st.reuse[p] = 1;
  // Householder QR of [a b; D1 0; 0 D2] with rhs [r; 0; 0]  (DenseQRSolver, Eigen reflector convention)
  double tau1, beta1, dv1;
  {
    const double tail = s.S11 + D1 * D1;
    if (tail <= DBL_MIN) { tau1 = 0.0; beta1 = s.a0; dv1 = 0.0; }
    else { beta1 = sqrt(s.a0 * s.a0 + tail); if (s.a0 >= 0.0) beta1 = -beta1; dv1 = s.a0 - beta1; tau1 = (beta1 - s.a0) / beta1; }
  }
  const double inv1 = (dv1 != 0.0) ? 1.0 / dv1 : 0.0;
  const double sc2 = s.S12 * inv1 + s.b0;            // v1 . column 2
  const double scr = s.S1r * inv1 + s.r0;            // v1 . rhs
  const double b0p = s.b0 - tau1 * sc2, r0p = s.r0 - tau1 * scr;
  const double essD1 = D1 * inv1;
  const double bD1 = -tau1 * sc2 * essD1, rD1 = -tau1 * scr * essD1, bD2 = D2;
  // second pass: column 2 and rhs after the first reflector, rows 1..m-1
  double b1p = 0.0, r1p = 0.0, T22 = 0.0, T2r = 0.0;
  {
    auto row = [&](int i, double xi, double yi) {
      const double e = model(m, xi, c);
      const double r = yi - e;
      const double a = -(e * xi) * s1, b = -e * s2;
      const double ess = a * inv1;
      const double bp = b - tau1 * sc2 * ess, rp = r - tau1 * scr * ess;
      if (i == 1) { b1p = bp; r1p = rp; }
      else { T22 += bp * bp; T2r += bp * rp; }
    };
    for_rows<false>(x, y, np, p, 1, nobs, row);      // plain loads: the candidate pass below reads the same lines again
  }
  double tau2, beta2, dv2;
  {
    const double tail = T22 + bD1 * bD1 + bD2 * bD2;
    if (tail <= DBL_MIN) { tau2 = 0.0; beta2 = b1p; dv2 = 0.0; }
    else { beta2 = sqrt(b1p * b1p + tail); if (b1p >= 0.0) beta2 = -beta2; dv2 = b1p - beta2; tau2 = (beta2 - b1p) / beta2; }
  }
  const double inv2 = (dv2 != 0.0) ? 1.0 / dv2 : 0.0;
  const double sbr = (T2r + bD1 * rD1) * inv2 + r1p;  // v2 . rhs  (the D2 row of the rhs is zero)
  const double r1pp = r1p - tau2 * sbr;
  const double y2 = r1pp / beta2;
  const double y1 = (r0p - b0p * y2) / beta1;
  const double st1 = -y1, st2 = -y2;                 // LM solves J y = r and steps by -y
  // model cost change = -(J step).(r + J step / 2) from the accumulated sums
  const double Jr = st1 * (s.S1r + s.a0 * s.r0) + st2 * s.S2r;
  const double JJ = st1 * st1 * s.n1 + 2.0 * st1 * st2 * s.S22 + st2 * st2 * s.n2;
  const double mcc = -(Jr + 0.5 * JJ);
  const bool finite_step = (st1 == st1) && (st2 == st2) && !isinf(st1) && !isinf(st2);
  const bool valid = finite_step && (mcc > 0.0);
  if (!valid) {                                       // HandleInvalidStep
    st.nci[p] += 1;
    if (st.nci[p] >= prm.max_num_consecutive_invalid_steps) { st.done[p] = 1; st.term[p] = SK_FAILURE; return false; }
    st.radius[p] = radius / st.decrease[p]; st.decrease[p] *= 2.0;
    return finalize(st, p, false, st.grad_max[p], x_cost, prm);
  }
  st.nci[p] = 0;
  const double dm = st1 * s1, dc = st2 * s2;
  const double mcand = m + dm, ccand = c + dc;
  Sums t;                                            // cost AND Jacobian sums at the candidate: an accepted step needs both
  pass_jacobian(x, y, np, p, nobs, mcand, ccand, s1, s2, &t);
  const double cand_cost = t.cost;
  const double cc = (cand_cost == cand_cost) ? cand_cost : DBL_MAX;
  const double dxm = m - mcand, dxc = c - ccand;
  const double step_norm = sqrt(dxm * dxm + dxc * dxc);
  if (step_norm <= prm.parameter_tolerance * (st.x_norm[p] + prm.parameter_tolerance)) { st.done[p] = 1; st.term[p] = SK_CONVERGENCE; return false; }
  const double cost_change = x_cost - cc;
  if (fabs(cost_change) <= prm.function_tolerance * x_cost) { st.done[p] = 1; st.term[p] = SK_CONVERGENCE; return false; }
  const double rho = cost_change / mcc;
  if (rho > prm.min_relative_decrease) {              // HandleSuccessfulStep
    mc[p] = mcand; mc[np + p] = ccand;
    save_sums(st.sums, np, p, t);
    st.x_cost[p] = t.cost; st.x_norm[p] = sqrt(mcand * mcand + ccand * ccand);
    const double gm = fmax(fabs(mcand - (mcand + (-t.g1))), fabs(ccand - (ccand + (-t.g2))));
    st.grad_max[p] = gm;
    const double q = 2.0 * rho - 1.0;
    radius = radius / fmax(1.0 / 3.0, 1.0 - q * q * q);
    st.radius[p] = fmin(prm.max_radius, radius);
    st.decrease[p] = 2.0; st.reuse[p] = 0;
    return finalize(st, p, true, gm, t.cost, prm);
  }
  st.radius[p] = radius / st.decrease[p]; st.decrease[p] *= 2.0;   // HandleUnsuccessfulStep
  return finalize(st, p, false, st.grad_max[p], cc, prm);
}

__global__ void __launch_bounds__(BT, SK_BATCH_MINB) k_batch_iterate(int64_t np, int nobs, const double* __restrict__ x, const double* __restrict__ y,
                                                      double* __restrict__ mc, BatchState st, LmParams prm, int* active_count) {
  const int64_t p = blockIdx.x * (int64_t)BT + threadIdx.x;
  bool active = false;
  if (p < np && !st.done[p]) active = iterate_one(np, nobs, x, y, mc, st, prm, p);
  publish_active(active, active_count);
}

}  // namespace

void curve_fit_batch_solve(const sk_solver_options& opt, int64_t np, int nobs, const double* x, const double* y, double* mc,
                           double* out_initial_cost, double* out_final_cost, int32_t* out_num_iterations,
                           int32_t* out_termination_type, sk_solver_summary* S) {
  SK_REQUIRE(nobs >= 2, SK_ERR_INVALID_ARGUMENT, "batched curve fits need at least 2 observations per problem");
  const auto t0 = std::chrono::steady_clock::now();
  cudaStream_t stream;
  SK_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sg{stream};
  DBuf<double> dbl((size_t)(11 + kSums) * np);
  DBuf<int> ints((size_t)7 * np);
  DBuf<int> active(1);
  HBuf<int> active_h(1);
  BatchState st;
  double* dp = dbl.p;
  st.radius = dp; st.decrease = dp + np; st.x_cost = dp + 2 * np; st.x_norm = dp + 3 * np; st.scale1 = dp + 4 * np;
  st.scale2 = dp + 5 * np; st.diag1 = dp + 6 * np; st.diag2 = dp + 7 * np; st.initial_cost = dp + 8 * np;
  st.final_cost = dp + 9 * np; st.grad_max = dp + 10 * np; st.sums = dp + 11 * np;
  int* ip = ints.p;
  st.iteration = ip; st.reuse = ip + np; st.nci = ip + 2 * np; st.done = ip + 3 * np; st.term = ip + 4 * np;
  st.nsucc = ip + 5 * np; st.nunsucc = ip + 6 * np;
  LmParams prm{};
  prm.max_num_iterations = opt.max_num_iterations; prm.max_num_consecutive_invalid_steps = opt.max_num_consecutive_invalid_steps;
  prm.max_radius = opt.max_trust_region_radius; prm.min_radius = opt.min_trust_region_radius;
  prm.min_relative_decrease = opt.min_relative_decrease; prm.min_lm_diagonal = opt.min_lm_diagonal; prm.max_lm_diagonal = opt.max_lm_diagonal;
  prm.function_tolerance = opt.function_tolerance; prm.gradient_tolerance = opt.gradient_tolerance;
  prm.parameter_tolerance = opt.parameter_tolerance; prm.eta = opt.eta;
  const int blocks = cdiv(np, BT);
  Profiler prof; prof.enabled = opt.profile_kernels != 0; prof.stream = stream;
  int launches = 0;
  active.zero(stream);
  { KScope k(prof, SK_KF_EVALUATE_JACOBIAN);
    k_batch_init<<<blocks, BT, 0, stream>>>(np, nobs, x, y, mc, st, prm, opt.initial_trust_region_radius, opt.jacobi_scaling, active.p); }
  check_launch("k_batch_init"); ++launches;
  SK_CUDA(cudaMemcpyAsync(active_h.p, active.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  SK_CUDA(cudaStreamSynchronize(stream));
  int lm_iters = 0;
  while (*active_h.p > 0) {
    // every active problem retires after at most max_num_iterations iterations; never spin past that
    SK_REQUIRE(lm_iters <= opt.max_num_iterations + 1, SK_ERR_INTERNAL,
               "batched solve: %d problems still active after %d launches (max_num_iterations %d)", *active_h.p, lm_iters, opt.max_num_iterations);
    active.zero(stream);
    { KScope k(prof, SK_KF_DENSE);
      k_batch_iterate<<<blocks, BT, 0, stream>>>(np, nobs, x, y, mc, st, prm, active.p); }
    check_launch("k_batch_iterate"); ++launches; ++lm_iters;
    SK_CUDA(cudaMemcpyAsync(active_h.p, active.p, sizeof(int), cudaMemcpyDeviceToHost, stream));   // the one scalar readback
    SK_CUDA(cudaStreamSynchronize(stream));
  }
  prof.collect();
  std::vector<int> h_int((size_t)7 * np);
  std::vector<double> h_init((size_t)np), h_final((size_t)np);
  SK_CUDA(cudaMemcpyAsync(h_int.data(), ints.p, sizeof(int) * h_int.size(), cudaMemcpyDeviceToHost, stream));
  SK_CUDA(cudaMemcpyAsync(h_init.data(), st.initial_cost, sizeof(double) * np, cudaMemcpyDeviceToHost, stream));
  SK_CUDA(cudaMemcpyAsync(h_final.data(), st.final_cost, sizeof(double) * np, cudaMemcpyDeviceToHost, stream));
  SK_CUDA(cudaStreamSynchronize(stream));
  sk_solver_summary_data& d = S->data;
  d = sk_solver_summary_data{};
  int64_t conv = 0;
  for (int64_t p = 0; p < np; ++p) {
    const int iters = h_int[5 * np + p] + h_int[6 * np + p];     // successful + unsuccessful rows
    if (out_initial_cost) out_initial_cost[p] = h_init[p];
    if (out_final_cost) out_final_cost[p] = h_final[p];
    if (out_num_iterations) out_num_iterations[p] = iters;
    if (out_termination_type) out_termination_type[p] = h_int[4 * np + p];
    d.initial_cost += h_init[p]; d.final_cost += h_final[p];
    d.num_successful_steps += h_int[5 * np + p]; d.num_unsuccessful_steps += h_int[6 * np + p];
    conv += h_int[4 * np + p] == SK_CONVERGENCE;
  }
  d.termination_type = conv == np ? SK_CONVERGENCE : SK_NO_CONVERGENCE;
  d.num_iterations = lm_iters + 1;
  d.linear_solver_type_used = SK_DENSE_QR; d.num_gpus = 1;
  d.num_parameter_blocks = 2 * np; d.num_parameters = 2 * np; d.num_residual_blocks = np * nobs; d.num_residuals = np * nobs;
  d.num_kernel_launches = launches;
  d.kernel_launches[SK_KF_EVALUATE_JACOBIAN] = 1; d.kernel_launches[SK_KF_DENSE] = lm_iters;
  for (int f = 0; f < SK_KF_COUNT; ++f) d.kernel_ms[f] = prof.ms[f];
  d.total_time_in_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  d.minimizer_time_in_seconds = d.total_time_in_seconds;
  S->rows.clear();
  S->message = fmt("batched solve: %lld of %lld problems converged in at most %d launches", (long long)conv, (long long)np, launches);
  format_reports(S);
}

}  // namespace sk
