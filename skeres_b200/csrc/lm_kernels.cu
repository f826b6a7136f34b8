// lm_kernels.cu — see lm_kernels.cuh.  Restates (Ceres sources are not in
// the reference tree): trust_region_minimizer.cc (A.3), levenberg_marquardt_strategy.cc (A.4),
// conjugate_gradients_solver.cc (A.7).
#include "lm_kernels.cuh"

#include <algorithm>
#include <cfloat>

namespace sk {

namespace {

constexpr int VT = 256;

__device__ __forceinline__ double block_sum256(double x, double* red) {
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
  return r;
}
__device__ __forceinline__ double block_max256(double x, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_down_sync(0xffffffffu, x, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = x;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < (int)(blockDim.x >> 5)) ? red[l] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) r = fmax(r, __shfl_down_sync(0xffffffffu, r, o));
  }
  return r;
}

// Fixed-order sum of a partial array by one block.
// Thread t adds part[t], part[t + 256], ... in that order; eight loads are in flight at a time (a tile-sized partial array
// is ~80 elements per thread, and one L2 round trip per element made this kernel cost 38 us).
__device__ double sum_partials(const double* part, int n, double* red) {
  double a = 0.0;
  const int step = blockDim.x;
  int i = threadIdx.x;
  for (; i + 7 * step < n; i += 8 * step) {
    double x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = __ldcg(part + i + u * step);
#pragma unroll
    for (int u = 0; u < 8; ++u) a += x[u];
  }
  for (; i < n; i += step) a += __ldcg(part + i);
  return block_sum256(a, red);
}

__global__ void k_jacobi_scale(int64_t n, const double* __restrict__ cnorm2, double* __restrict__ scale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    scale[i] = 1.0 / (1.0 + sqrt(cnorm2[i]));       // TrustRegionMinimizer: 1 / (1 + ||col||)
}

__global__ void k_lm_diagonal(int64_t n, const double* __restrict__ cnorm2, double* __restrict__ diagonal,
                              double* __restrict__ D, const LmDev* st, LmParams prm) {
  const bool reuse = st->reuse_diagonal != 0;
  const double radius = st->radius;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double d;
    if (!reuse) { d = fmin(fmax(cnorm2[i], prm.min_lm_diagonal), prm.max_lm_diagonal); diagonal[i] = d; }
    else d = diagonal[i];
    D[i] = sqrt(d / radius);
  }
}

// gradient norms through Plus(x, -g): (x - (x + (-g)))  — TrustRegionMinimizer::EvaluateGradientAndJacobian
__global__ void k_grad_norms(int64_t n, int64_t nc, const double* __restrict__ x, const double* __restrict__ g,
                             double* __restrict__ part, const int* guard) {
  if (guard != nullptr && *guard == 0) return;
  __shared__ double red[8];
  double sc = 0.0, sp = 0.0, mx = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double xi = x[i];
    const double d = xi - (xi + (-g[i]));
    if (i < nc) sc += d * d; else sp += d * d;
    mx = fmax(mx, fabs(d));
  }
  const double a = block_sum256(sc, red), b = block_sum256(sp, red), c = block_max256(mx, red);
  if (threadIdx.x == 0) { part[blockIdx.x] = a; part[kMaxPartials + blockIdx.x] = b; part[2 * kMaxPartials + blockIdx.x] = c; }
}

__global__ void k_candidate(int64_t n, int64_t nc, const double* __restrict__ x, const double* __restrict__ step,
                            const double* __restrict__ scale, double* __restrict__ cand, double* __restrict__ part) {
  __shared__ double red[8];
  double sc = 0.0, sp = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double delta = step[i] * scale[i];          // undo the Jacobian column scaling
    const double xi = x[i];
    const double ci = xi + delta;                     // Plus
    cand[i] = ci;
    const double d = xi - ci;
    if (i < nc) sc += d * d; else sp += d * d;
  }
  const double a = block_sum256(sc, red), b = block_sum256(sp, red);
  if (threadIdx.x == 0) { part[blockIdx.x] = a; part[kMaxPartials + blockIdx.x] = b; }
}

__global__ void k_accept(int64_t n, int64_t nc, double* __restrict__ x, const double* __restrict__ cand,
                         double* __restrict__ part, const int* guard) {
  if (guard != nullptr && *guard == 0) return;
  __shared__ double red[8];
  double sc = 0.0, sp = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ci = cand[i];
    x[i] = ci;
    if (i < nc) sc += ci * ci; else sp += ci * ci;
  }
  const double a = block_sum256(sc, red), b = block_sum256(sp, red);
  if (threadIdx.x == 0) { part[blockIdx.x] = a; part[kMaxPartials + blockIdx.x] = b; }
}

__global__ void k_negate(int64_t n, const double* __restrict__ in, double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = -in[i];
}

__global__ void k_fill(int64_t n, double value, double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = value;
}

struct ReduceJobs { ReduceJob j[8]; };
__global__ void k_reduce_jobs(ReduceJobs jobs, double* sbuf, const int* guard, FlagSources fl) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && fl.st != nullptr) {     // before the guard: the flags are consumed either way
    sbuf[SB_FLAG_EVAL] = fl.st->eval_failed ? 1.0 : 0.0;
    const bool fatal = (fl.pcg != nullptr && fl.pcg->termination == LIN_FATAL) || (fl.peer_error != nullptr && *fl.peer_error != 0);
    sbuf[SB_FLAG_LIN] = (fl.st->lin_error ? 1.0 : 0.0) + (fatal ? kFatalFlag : 0.0);
  }
  if (guard != nullptr && *guard == 0) return;
  __shared__ double red[8];
  const ReduceJob jb = jobs.j[blockIdx.x];
  double a = 0.0;
  if (jb.is_max) {
    const int step = blockDim.x;
    int i = threadIdx.x;
    for (; i + 7 * step < jb.n; i += 8 * step) {
      double x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) x[u] = __ldcg(jb.part + i + u * step);
#pragma unroll
      for (int u = 0; u < 8; ++u) a = fmax(a, x[u]);
    }
    for (; i < jb.n; i += step) a = fmax(a, __ldcg(jb.part + i));
    a = block_max256(a, red);
  }
  else a = sum_partials(jb.part, jb.n, red);
  if (threadIdx.x == 0) sbuf[jb.slot] = a;
}

// ---------------- scalar logic -------------------------------------------------------------------
__device__ void step_rejected(LmDev* st) {           // LevenbergMarquardtStrategy::StepRejected
  st->radius = st->radius / st->decrease_factor;
  st->decrease_factor *= 2.0;
  st->reuse_diagonal = 1;
}

__global__ void k_lm_init(LmDev* st, double initial_radius) {
  memset(st, 0, sizeof(LmDev));
  st->radius = initial_radius; st->decrease_factor = 2.0; st->reuse_diagonal = 0;
  st->x_cost = DBL_MAX; st->minimum_cost = DBL_MAX;
  st->termination_type = SK_NO_CONVERGENCE;
}

__global__ void k_lm_iter0(LmDev* st, const double* sbuf, LmParams prm) {   // IterationZero
  sk_iteration_summary& row = st->row;
  memset(&row, 0, sizeof(row));
  row.iteration = 0; row.eta = prm.eta;
  st->g_finalize = 1;
  if (st->eval_failed || sbuf[SB_FLAG_EVAL] != 0.0) {
    st->terminate = 1; st->termination_type = SK_FAILURE; st->term_reason = TR_EVALUATION_FAILED; st->g_finalize = 0; return;
  }
  st->x_cost = sbuf[SB_COST];
  st->x_norm = sqrt(sbuf[SB_XNORM_SQ_CAM] + sbuf[SB_XNORM_SQ_PT]);
  row.cost = st->x_cost + prm.fixed_cost;
  row.gradient_max_norm = sbuf[SB_GRAD_MAX];
  row.gradient_norm = sqrt(sbuf[SB_GRAD_SQ_CAM] + sbuf[SB_GRAD_SQ_PT]);
  row.step_is_valid = 1; row.step_is_successful = 1;
}

// After the linear solve + candidate: ComputeTrustRegionStep tail and HandleInvalidStep.
__global__ void k_lm_decide_a(LmDev* st, const PcgDev* pcg, const double* sbuf, LmParams prm) {
  sk_iteration_summary prev = st->row;
  sk_iteration_summary& row = st->row;
  memset(&row, 0, sizeof(row));
  st->iteration += 1;
  row.iteration = st->iteration;
  st->g_eval_cand = 0; st->g_accept = 0; st->g_finalize = 1;
  if (pcg != nullptr) { st->lin_iterations = pcg->iter; st->lin_termination = pcg->termination; }
  const double lin_flags = sbuf[SB_FLAG_LIN];            // summed over the ranks: every rank takes the same branch
  if (st->lin_error || lin_flags != 0.0) { st->lin_termination = LIN_FAILURE; st->lin_error = 0; }
  if (lin_flags >= kFatalFlag) st->lin_termination = LIN_FATAL;
  row.linear_solver_iterations = st->lin_iterations;
  if (st->lin_termination == LIN_FATAL) {
    st->terminate = 1; st->termination_type = SK_FAILURE; st->term_reason = TR_LINEAR_SOLVER_FATAL; st->g_finalize = 0; return;
  }
  bool valid = false;
  if (st->lin_termination != LIN_FAILURE) {
    st->model_cost_change = -sbuf[SB_MCC];
    valid = st->model_cost_change > 0.0;              // NaN (non-finite step) compares false
  }
  row.step_is_valid = valid ? 1 : 0;
  if (!valid) {                                       // HandleInvalidStep
    st->num_consecutive_invalid += 1;
    if (st->num_consecutive_invalid >= prm.max_num_consecutive_invalid_steps) {
      st->terminate = 1; st->termination_type = SK_FAILURE; st->term_reason = TR_INVALID_STEPS; st->g_finalize = 0; return;
    }
    step_rejected(st);                                // StepIsInvalid == StepRejected(0)
    row.cost = st->x_cost + prm.fixed_cost; row.cost_change = 0.0;
    row.gradient_max_norm = prev.gradient_max_norm; row.gradient_norm = prev.gradient_norm;
    row.step_norm = 0.0; row.relative_decrease = 0.0; row.eta = prm.eta;
    return;
  }
  st->num_consecutive_invalid = 0;
  st->g_eval_cand = 1;
}

// After the candidate cost: ParameterToleranceReached, FunctionToleranceReached, IsStepSuccessful,
// HandleUnsuccessfulStep.
__global__ void k_lm_decide_b(LmDev* st, const double* sbuf, LmParams prm) {
  if (st->g_eval_cand == 0) return;
  sk_iteration_summary& row = st->row;
  double cand = sbuf[SB_COST];
  if (st->eval_failed || sbuf[SB_FLAG_EVAL] != 0.0 || !(cand == cand)) { cand = DBL_MAX; st->eval_failed = 0; }   // "step failed to evaluate"
  st->cand_cost = cand;
  row.step_norm = sqrt(sbuf[SB_STEP_SQ_CAM] + sbuf[SB_STEP_SQ_PT]);
  const double step_size_tolerance = prm.parameter_tolerance * (st->x_norm + prm.parameter_tolerance);
  if (row.step_norm <= step_size_tolerance) {
    st->terminate = 1; st->termination_type = SK_CONVERGENCE; st->term_reason = TR_PARAMETER_TOLERANCE;
    st->term_v1 = row.step_norm / (st->x_norm + prm.parameter_tolerance); st->term_v2 = prm.parameter_tolerance;
    st->g_finalize = 0; st->g_accept = 0; return;
  }
  row.cost_change = st->x_cost - cand;
  const double absolute_function_tolerance = prm.function_tolerance * st->x_cost;
  if (fabs(row.cost_change) <= absolute_function_tolerance) {
    st->terminate = 1; st->termination_type = SK_CONVERGENCE; st->term_reason = TR_FUNCTION_TOLERANCE;
    st->term_v1 = fabs(row.cost_change) / st->x_cost; st->term_v2 = prm.function_tolerance;
    st->g_finalize = 0; st->g_accept = 0; return;
  }
  row.relative_decrease = (st->x_cost - cand) / st->model_cost_change;   // StepQuality, monotonic steps
  row.eta = prm.eta;
  if (row.relative_decrease > prm.min_relative_decrease) { st->g_accept = 1; return; }
  st->g_accept = 0;                                  // HandleUnsuccessfulStep
  row.step_is_successful = 0;
  step_rejected(st);
  row.cost = cand + prm.fixed_cost;
}

// HandleSuccessfulStep, after x = candidate and the Jacobian evaluation at the new point.
__global__ void k_lm_post_accept(LmDev* st, const double* sbuf, LmParams prm) {
  if (st->g_accept == 0) return;
  sk_iteration_summary& row = st->row;
  if (st->eval_failed || sbuf[SB_FLAG_EVAL] != 0.0) {
    st->terminate = 1; st->termination_type = SK_FAILURE; st->term_reason = TR_EVALUATION_FAILED; st->g_finalize = 0; return;
  }
  st->x_norm = sqrt(sbuf[SB_XNORM_SQ_CAM] + sbuf[SB_XNORM_SQ_PT]);
  st->x_cost = sbuf[SB_COST];
  row.cost = st->x_cost + prm.fixed_cost;
  row.gradient_max_norm = sbuf[SB_GRAD_MAX];
  row.gradient_norm = sqrt(sbuf[SB_GRAD_SQ_CAM] + sbuf[SB_GRAD_SQ_PT]);
  row.step_is_successful = 1;
  // LevenbergMarquardtStrategy::StepAccepted
  const double q = row.relative_decrease;
  const double t = 2.0 * q - 1.0;
  st->radius = st->radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
  st->radius = fmin(prm.max_radius, st->radius);
  st->decrease_factor = 2.0;
  st->reuse_diagonal = 0;
}

// FinalizeIterationAndCheckIfMinimizerCanContinue
__global__ void k_lm_finalize(LmDev* st, sk_iteration_summary* rows, int cap, LmParams prm) {
  if (st->g_finalize == 0) return;
  sk_iteration_summary& row = st->row;
  if (row.step_is_successful) {
    st->num_successful += 1;
    if (st->x_cost < st->minimum_cost) { st->minimum_cost = st->x_cost; row.step_is_nonmonotonic = 0; }
    else row.step_is_nonmonotonic = 1;
  } else st->num_unsuccessful += 1;
  row.trust_region_radius = st->radius;
  if (st->num_rows < cap) rows[st->num_rows] = row;
  st->num_rows += 1;
  if (row.iteration >= prm.max_num_iterations) {
    st->terminate = 1; st->termination_type = SK_NO_CONVERGENCE; st->term_reason = TR_MAX_ITERATIONS; st->term_v1 = row.iteration; return;
  }
  if (row.step_is_successful && row.gradient_max_norm <= prm.gradient_tolerance) {
    st->terminate = 1; st->termination_type = SK_CONVERGENCE; st->term_reason = TR_GRADIENT_TOLERANCE;
    st->term_v1 = row.gradient_max_norm; st->term_v2 = prm.gradient_tolerance; return;
  }
  if (row.trust_region_radius <= prm.min_radius) {
    st->terminate = 1; st->termination_type = SK_CONVERGENCE; st->term_reason = TR_MIN_RADIUS;
    st->term_v1 = row.trust_region_radius; st->term_v2 = prm.min_radius; return;
  }
}

}  // namespace

int vec_blocks(int64_t n) { return (int)std::min<int64_t>((n + VT - 1) / VT, 1184); }

void launch_jacobi_scale(int64_t n, const double* cnorm2, double* scale, cudaStream_t s) {
  k_jacobi_scale<<<vec_blocks(n), VT, 0, s>>>(n, cnorm2, scale); check_launch("k_jacobi_scale");
}
void launch_lm_diagonal(int64_t n, const double* cnorm2, double* diagonal, double* D, const LmDev* st, LmParams prm, cudaStream_t s) {
  k_lm_diagonal<<<vec_blocks(n), VT, 0, s>>>(n, cnorm2, diagonal, D, st, prm); check_launch("k_lm_diagonal");
}
void launch_grad_norms(int64_t n, int64_t nc, const double* x, const double* g, double* part, const int* guard, cudaStream_t s) {
  k_grad_norms<<<vec_blocks(n), VT, 0, s>>>(n, nc, x, g, part, guard); check_launch("k_grad_norms");
}
void launch_candidate(int64_t n, int64_t nc, const double* x, const double* step, const double* scale, double* cand, double* part, cudaStream_t s) {
  k_candidate<<<vec_blocks(n), VT, 0, s>>>(n, nc, x, step, scale, cand, part); check_launch("k_candidate");
}
void launch_accept(int64_t n, int64_t nc, double* x, const double* cand, double* part, const int* guard, cudaStream_t s) {
  k_accept<<<vec_blocks(n), VT, 0, s>>>(n, nc, x, cand, part, guard); check_launch("k_accept");
}
void launch_negate(int64_t n, const double* in, double* out, cudaStream_t s) {
  k_negate<<<vec_blocks(n), VT, 0, s>>>(n, in, out); check_launch("k_negate");
}
void launch_fill(int64_t n, double value, double* out, cudaStream_t s) {
  k_fill<<<vec_blocks(n), VT, 0, s>>>(n, value, out); check_launch("k_fill");
}
void launch_reduce_jobs(const ReduceJob* jobs, int njobs, double* sbuf, const int* guard, FlagSources flags, cudaStream_t s) {
  ReduceJobs j;
  SK_REQUIRE(njobs >= 1 && njobs <= 8, SK_ERR_INTERNAL, "bad reduce job count");
  for (int i = 0; i < njobs; ++i) j.j[i] = jobs[i];
  k_reduce_jobs<<<njobs, VT, 0, s>>>(j, sbuf, guard, flags); check_launch("k_reduce_jobs");
}
void launch_lm_init(LmDev* st, double r0, cudaStream_t s) { k_lm_init<<<1, 1, 0, s>>>(st, r0); check_launch("k_lm_init"); }
void launch_lm_iter0(LmDev* st, const double* sbuf, LmParams prm, cudaStream_t s) { k_lm_iter0<<<1, 1, 0, s>>>(st, sbuf, prm); check_launch("k_lm_iter0"); }
void launch_lm_decide_a(LmDev* st, const PcgDev* pcg, const double* sbuf, LmParams prm, cudaStream_t s) { k_lm_decide_a<<<1, 1, 0, s>>>(st, pcg, sbuf, prm); check_launch("k_lm_decide_a"); }
void launch_lm_decide_b(LmDev* st, const double* sbuf, LmParams prm, cudaStream_t s) { k_lm_decide_b<<<1, 1, 0, s>>>(st, sbuf, prm); check_launch("k_lm_decide_b"); }
void launch_lm_post_accept(LmDev* st, const double* sbuf, LmParams prm, cudaStream_t s) { k_lm_post_accept<<<1, 1, 0, s>>>(st, sbuf, prm); check_launch("k_lm_post_accept"); }
void launch_lm_finalize(LmDev* st, sk_iteration_summary* rows, int cap, LmParams prm, cudaStream_t s) { k_lm_finalize<<<1, 1, 0, s>>>(st, rows, cap, prm); check_launch("k_lm_finalize"); }

}  // namespace sk
