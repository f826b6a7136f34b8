// lm_solver.cu — host side of the device-resident LM loop (see lm_solver.cuh).
#include "lm_solver.cuh"

#include <chrono>
#include <cmath>
#include <cstdlib>

namespace sk {

static double wall() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

LmSolver::LmSolver(const sk_solver_options& opt, cudaStream_t stream) : opt_(opt), stream_(stream), comm_(opt.comm) {
  prm_.max_num_iterations = opt.max_num_iterations;
  prm_.max_num_consecutive_invalid_steps = opt.max_num_consecutive_invalid_steps;
  prm_.max_radius = opt.max_trust_region_radius; prm_.min_radius = opt.min_trust_region_radius;
  prm_.min_relative_decrease = opt.min_relative_decrease;
  prm_.min_lm_diagonal = opt.min_lm_diagonal; prm_.max_lm_diagonal = opt.max_lm_diagonal;
  prm_.function_tolerance = opt.function_tolerance; prm_.gradient_tolerance = opt.gradient_tolerance;
  prm_.parameter_tolerance = opt.parameter_tolerance; prm_.eta = opt.eta; prm_.fixed_cost = 0.0;
  prof_.enabled = opt.profile_kernels != 0;
  prof_.family_mask = (opt.profile_kernels == 2) ? ((1u << SK_KF_SCHUR_MATVEC) | (1u << SK_KF_PCG_SOLVE)) : ~0u;
  prof_.stream = stream;
}

LmSolver::~LmSolver() { if (ev_start_) { cudaEventDestroy(ev_start_); cudaEventDestroy(ev_stop_); } }

void LmSolver::allocate(int64_t n, int64_t nc) {
  n_ = n; nc_ = nc;
  x_.alloc(n); cand_.alloc(n); step_.alloc(n); scale_.alloc(n); grad_.alloc(n); cnorm2_.alloc(n);
  diagonal_.alloc(n); D_.alloc(n);
  sbuf_.alloc(SB_COUNT); sbuf_.zero(stream_);
  part_a_.alloc(3 * kMaxPartials); part_b_.alloc(3 * kMaxPartials);
  st_.alloc(1); st_h_.alloc(1);
  rows_cap_ = std::max(opt_.max_num_iterations, 0) + 2;
  rows_.alloc(rows_cap_);
}

void LmSolver::reduce(std::initializer_list<ReduceJob> jobs, const int* guard) {
  std::vector<ReduceJob> v(jobs);
  KScope k(prof_, SK_KF_LM);
  launch_reduce_jobs(v.data(), (int)v.size(), sbuf_.p, guard, FlagSources{st_.p, flag_pcg_, flag_peer_error_}, stream_);
}

// The sum / max allreduces that follow a Jacobian evaluation: cost, evaluation-failure flag, |g|^2 and |x|^2 of the point part;
// max |g| and the elapsed host time.
void LmSolver::allreduce_after_jacobian(double t_start) {
  if (comm_ == nullptr || comm_->world <= 1) return;
  time_h_.p[0] = wall() - t_start;
  SK_CUDA(cudaMemcpyAsync(sbuf_.p + SB_TIME, time_h_.p, sizeof(double), cudaMemcpyHostToDevice, stream_));
  comm_allreduce_sum(comm_, sbuf_.p, 4, stream_);
  comm_allreduce_max(comm_, sbuf_.p + SB_GRAD_MAX, 2, stream_);
}

void LmSolver::minimize(sk_solver_summary* S, int max_num_iterations_override) {
  const double t_start = wall();
  sk_solver_summary_data& d = S->data;
  if (max_num_iterations_override >= 0) {
    SK_REQUIRE(max_num_iterations_override <= opt_.max_num_iterations, SK_ERR_INVALID_ARGUMENT,
               "max_num_iterations override %d exceeds the %d the solver was created with", max_num_iterations_override, opt_.max_num_iterations);
    prm_.max_num_iterations = max_num_iterations_override;
  } else prm_.max_num_iterations = opt_.max_num_iterations;
  n_res_evals_ = n_jac_evals_ = n_lin_solves_ = n_lin_iters_ = 0;
  for (int f = 0; f < SK_KF_COUNT; ++f) { prof_.launches[f] = 0; prof_.ms[f] = 0.0; }
  if (!ev_start_) { SK_CUDA(cudaEventCreate(&ev_start_)); SK_CUDA(cudaEventCreate(&ev_stop_)); }
  SK_CUDA(cudaEventRecord(ev_start_, stream_));
  const int nb = vec_blocks(n_);
  const bool multi = comm_ != nullptr && comm_->world > 1;
  if (time_h_.n == 0) time_h_.alloc(2);
  time_h_.p[0] = time_h_.p[1] = 0.0;
  std::vector<double> t_iter, t_cum;
  auto readback = [&]() -> const LmDev& {
    SK_CUDA(cudaMemcpyAsync(st_h_.p, st_.p, sizeof(LmDev), cudaMemcpyDeviceToHost, stream_));
    if (multi) SK_CUDA(cudaMemcpyAsync(time_h_.p + 1, sbuf_.p + SB_TIME, sizeof(double), cudaMemcpyDeviceToHost, stream_));
    SK_CUDA(cudaStreamSynchronize(stream_));
    prof_.collect();
    return *st_h_.p;
  };
  auto print_row = [&](const sk_iteration_summary& r, double it_s, double cum_s) {
    if (!opt_.minimizer_progress_to_stdout) return;
    if (r.iteration == 0)
      std::printf("iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius  ls_iter  iter_time  total_time\n");
    std::printf("% 4d % 8e   % 3.2e   % 3.2e  % 3.2e  % 3.2e % 3.2e     % 4d   % 3.2e   % 3.2e\n", r.iteration, r.cost, r.cost_change,
                r.gradient_max_norm, r.step_norm, r.relative_decrease, r.trust_region_radius, r.linear_solver_iterations, it_s, cum_s);
    std::fflush(stdout);
  };
  int* g_eval_cand = &st_.p->g_eval_cand;
  int* g_accept = &st_.p->g_accept;

  { KScope k(prof_, SK_KF_LM); launch_lm_init(st_.p, opt_.initial_trust_region_radius, stream_); }
  load_state();
  // ---- IterationZero -----------------------------------------------------------------------------
  double t_it = wall();
  { KScope k(prof_, SK_KF_LM); launch_accept(n_, nc_, x_.p, x_.p, part_b_.p, nullptr, stream_); }   // ||x||^2 partials
  if (opt_.jacobi_scaling) {
    eval_jacobian(false, false, nullptr);
    KScope k(prof_, SK_KF_LM);
    launch_jacobi_scale(n_, cnorm2_.p, scale_.p, stream_);
  } else {
    KScope k(prof_, SK_KF_LM);
    launch_fill(n_, 1.0, scale_.p, stream_);
  }
  eval_jacobian(true, true, nullptr);
  ++n_jac_evals_;
  { KScope k(prof_, SK_KF_LM); launch_grad_norms(n_, nc_, x_.p, grad_.p, part_a_.p, nullptr, stream_); }
  reduce({cost_job(), {part_a_.p, nb, SB_GRAD_SQ_CAM, 0}, {part_a_.p + kMaxPartials, nb, SB_GRAD_SQ_PT, 0},
          {part_a_.p + 2 * kMaxPartials, nb, SB_GRAD_MAX, 1}, {part_b_.p, nb, SB_XNORM_SQ_CAM, 0},
          {part_b_.p + kMaxPartials, nb, SB_XNORM_SQ_PT, 0}}, nullptr);
  allreduce_after_jacobian(t_start);
  { KScope k(prof_, SK_KF_LM, 2); launch_lm_iter0(st_.p, sbuf_.p, prm_, stream_); launch_lm_finalize(st_.p, rows_.p, rows_cap_, prm_, stream_); }
  {
    const LmDev& h = readback();
    d.initial_cost = h.x_cost + prm_.fixed_cost;
    const double now = wall();
    if (h.g_finalize) { t_iter.push_back(now - t_it); t_cum.push_back(now - t_start + d.preprocessor_time_in_seconds); print_row(h.row, t_iter.back(), t_cum.back()); }
  }
  // ---- main loop ---------------------------------------------------------------------------------
  while (!st_h_.p->terminate) {
    t_it = wall();
    // Multi-GPU: every rank must take this decision alike (a rank that left alone would leave the others waiting in the next
    // collective), so it is taken on the maximum over the ranks of the elapsed time each rank recorded before the last scalar
    // allreduce (SB_TIME), which came back with the state block.
    if ((multi ? time_h_.p[1] : wall() - t_start) > opt_.max_solver_time_in_seconds) {
      st_h_.p->terminate = 1; st_h_.p->termination_type = SK_NO_CONVERGENCE; st_h_.p->term_reason = TR_MAX_TIME; break;
    }
    { KScope k(prof_, SK_KF_LM); launch_lm_diagonal(n_, cnorm2_.p, diagonal_.p, D_.p, st_.p, prm_, stream_); }
    const PcgDev* pcg = nullptr;
    const ReduceJob mcc_job = linear_solve(&pcg);
    ++n_lin_solves_;
    { KScope k(prof_, SK_KF_LM); launch_candidate(n_, nc_, x_.p, step_.p, scale_.p, cand_.p, part_a_.p, stream_); }
    reduce({mcc_job, {part_a_.p, nb, SB_STEP_SQ_CAM, 0}, {part_a_.p + kMaxPartials, nb, SB_STEP_SQ_PT, 0}}, nullptr);
    comm_allreduce_sum(comm_, sbuf_.p + SB_MCC, 3, stream_);       // model cost change, |step|^2 (points), linear-solver failure flags
    { KScope k(prof_, SK_KF_LM); launch_lm_decide_a(st_.p, pcg, sbuf_.p, prm_, stream_); }
    eval_cost(cand_.p, g_eval_cand);
    ++n_res_evals_;
    reduce({cost_job()}, g_eval_cand);
    comm_allreduce_sum(comm_, sbuf_.p, 2, stream_);                 // candidate cost, evaluation-failure flag
    { KScope k(prof_, SK_KF_LM); launch_lm_decide_b(st_.p, sbuf_.p, prm_, stream_); }
    // HandleSuccessfulStep (all guarded by g_accept)
    { KScope k(prof_, SK_KF_LM); launch_accept(n_, nc_, x_.p, cand_.p, part_b_.p, g_accept, stream_); }
    eval_jacobian(true, true, g_accept);
    { KScope k(prof_, SK_KF_LM); launch_grad_norms(n_, nc_, x_.p, grad_.p, part_a_.p, g_accept, stream_); }
    reduce({cost_job(), {part_a_.p, nb, SB_GRAD_SQ_CAM, 0}, {part_a_.p + kMaxPartials, nb, SB_GRAD_SQ_PT, 0},
            {part_a_.p + 2 * kMaxPartials, nb, SB_GRAD_MAX, 1}, {part_b_.p, nb, SB_XNORM_SQ_CAM, 0},
            {part_b_.p + kMaxPartials, nb, SB_XNORM_SQ_PT, 0}}, g_accept);
    allreduce_after_jacobian(t_start);
    { KScope k(prof_, SK_KF_LM, 2); launch_lm_post_accept(st_.p, sbuf_.p, prm_, stream_); launch_lm_finalize(st_.p, rows_.p, rows_cap_, prm_, stream_); }
    const LmDev& h = readback();
    if (h.g_accept) ++n_jac_evals_;
    n_lin_iters_ += h.lin_iterations;
    note_linear_iterations(h.lin_iterations);
    const double now = wall();
    if (h.g_finalize) { t_iter.push_back(now - t_it); t_cum.push_back(now - t_start + d.preprocessor_time_in_seconds); print_row(h.row, t_iter.back(), t_cum.back()); }
  }
  // ---- summary -------------------------------------------------------------------------------------
  const LmDev h = *st_h_.p;
  const int nrows = std::min(h.num_rows, rows_cap_);
  S->rows.resize(nrows);
  if (nrows) SK_CUDA(cudaMemcpyAsync(S->rows.data(), rows_.p, sizeof(sk_iteration_summary) * nrows, cudaMemcpyDeviceToHost, stream_));
  d.termination_type = h.termination_type;
  const bool usable = h.termination_type == SK_CONVERGENCE || h.termination_type == SK_NO_CONVERGENCE || h.termination_type == SK_USER_SUCCESS;
  if (usable) store_state();
  SK_CUDA(cudaEventRecord(ev_stop_, stream_));
  SK_CUDA(cudaStreamSynchronize(stream_));
  prof_.collect();
  { float ms = 0; cudaEventElapsedTime(&ms, ev_start_, ev_stop_); d.minimizer_device_time_in_seconds = ms * 1e-3; }
  for (int i = 0; i < nrows && i < (int)t_iter.size(); ++i) {
    S->rows[i].iteration_time_in_seconds = t_iter[i];
    S->rows[i].cumulative_time_in_seconds = t_cum[i];
  }
  d.num_successful_steps = h.num_successful; d.num_unsuccessful_steps = h.num_unsuccessful;
  d.num_iterations = nrows;
  d.fixed_cost = prm_.fixed_cost;
  d.final_cost = d.initial_cost;                       // SetSummaryFinalCost (solver.cc)
  for (auto& r : S->rows) d.final_cost = std::min(d.final_cost, r.cost);
  d.num_residual_evaluations = n_res_evals_; d.num_jacobian_evaluations = n_jac_evals_;
  d.num_linear_solves = n_lin_solves_; d.total_linear_solver_iterations = n_lin_iters_;
  d.num_kernel_launches = 0;
  for (int f = 0; f < SK_KF_COUNT; ++f) { d.kernel_launches[f] = prof_.launches[f]; d.kernel_ms[f] = prof_.ms[f]; d.num_kernel_launches += prof_.launches[f]; }
  d.linear_solver_type_used = opt_.linear_solver_type; d.preconditioner_type_used = opt_.preconditioner_type;
  d.num_gpus = comm_ ? comm_->world : 1;
  fill_summary(&d);
  S->message = termination_message(h, opt_);
  d.minimizer_time_in_seconds = wall() - t_start;
  if (getenv("SKERES_TRACE_HOST")) {
    std::fprintf(stderr, "[skeres trace] minimize: host %.1f ms, device %.1f ms, launches %lld, nccl enqueue host %.1f ms over %ld calls (cumulative)\n",
                 1e3 * d.minimizer_time_in_seconds, 1e3 * d.minimizer_device_time_in_seconds, (long long)d.num_kernel_launches,
                 1e3 * g_comm_host_seconds, g_comm_calls);
  }
}

std::string termination_message(const LmDev& st, const sk_solver_options& opt) {
  switch (st.term_reason) {
    case TR_MAX_ITERATIONS: return fmt("Maximum number of iterations reached. Number of iterations: %d.", (int)st.term_v1);
    case TR_GRADIENT_TOLERANCE: return fmt("Gradient tolerance reached. Gradient max norm: %e <= %e", st.term_v1, st.term_v2);
    case TR_MIN_RADIUS: return fmt("Minimum trust region radius reached. Trust region radius: %e <= %e", st.term_v1, st.term_v2);
    case TR_PARAMETER_TOLERANCE: return fmt("Parameter tolerance reached. Relative step_norm: %e <= %e.", st.term_v1, st.term_v2);
    case TR_FUNCTION_TOLERANCE: return fmt("Function tolerance reached. |cost_change|/cost: %e <= %e", st.term_v1, st.term_v2);
    case TR_INVALID_STEPS: return fmt("Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps: %d", opt.max_num_consecutive_invalid_steps);
    case TR_LINEAR_SOLVER_FATAL: return "Linear solver failed due to unrecoverable non-numeric causes. Please see the error log for clues. ";
    case TR_EVALUATION_FAILED: return "Residual and Jacobian evaluation failed.";
    case TR_MAX_TIME: return "Maximum solver time reached.";
  }
  return "";
}

static const char* linear_solver_name(int t) {
  switch (t) {
    case SK_DENSE_NORMAL_CHOLESKY: return "DENSE_NORMAL_CHOLESKY"; case SK_DENSE_QR: return "DENSE_QR";
    case SK_SPARSE_NORMAL_CHOLESKY: return "SPARSE_NORMAL_CHOLESKY"; case SK_DENSE_SCHUR: return "DENSE_SCHUR";
    case SK_SPARSE_SCHUR: return "SPARSE_SCHUR"; case SK_ITERATIVE_SCHUR: return "ITERATIVE_SCHUR"; case SK_CGNR: return "CGNR";
  }
  return "UNKNOWN";
}
static const char* termination_name(int t) {
  switch (t) {
    case SK_CONVERGENCE: return "CONVERGENCE"; case SK_NO_CONVERGENCE: return "NO_CONVERGENCE"; case SK_FAILURE: return "FAILURE";
    case SK_USER_SUCCESS: return "USER_SUCCESS"; case SK_USER_FAILURE: return "USER_FAILURE";
  }
  return "UNKNOWN";
}
static const char* preconditioner_name(int t) {
  switch (t) { case SK_IDENTITY: return "IDENTITY"; case SK_JACOBI: return "JACOBI"; case SK_SCHUR_JACOBI: return "SCHUR_JACOBI";
               case SK_CLUSTER_JACOBI: return "CLUSTER_JACOBI"; case SK_CLUSTER_TRIDIAGONAL: return "CLUSTER_TRIDIAGONAL"; }
  return "UNKNOWN";
}

// Solver::Summary::BriefReport / FullReport (ceres/solver.cc), restated for the fields this library fills.
void format_reports(sk_solver_summary* s) {
  const sk_solver_summary_data& d = s->data;
  s->brief = fmt("Ceres Solver Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s",
                 d.num_successful_steps + d.num_unsuccessful_steps, d.initial_cost, d.final_cost, termination_name(d.termination_type));
  std::string r = "\nSolver Summary (v skeres-b200 / sm_100a device solver)\n\n";
  r += fmt("%-25s %12s\n", "", "Original");
  r += fmt("%-25s %12lld\n", "Parameter blocks", (long long)d.num_parameter_blocks);
  r += fmt("%-25s %12lld\n", "Parameters", (long long)d.num_parameters);
  r += fmt("%-25s %12lld\n", "Residual blocks", (long long)d.num_residual_blocks);
  r += fmt("%-25s %12lld\n", "Residuals", (long long)d.num_residuals);
  r += fmt("\nMinimizer %19s\n", "TRUST_REGION");
  r += fmt("Trust region strategy %19s\n", "LEVENBERG_MARQUARDT");
  r += fmt("\n%-25s %25s\n", "Linear solver", linear_solver_name(d.linear_solver_type_used));
  if (d.linear_solver_type_used == SK_ITERATIVE_SCHUR) r += fmt("%-25s %25s\n", "Preconditioner", preconditioner_name(d.preconditioner_type_used));
  r += fmt("%-25s %25d\n", "GPUs", d.num_gpus);
  r += fmt("\nCost:\n%-25s %25e\n%-25s %25e\n%-25s %25e\n", "Initial", d.initial_cost, "Final", d.final_cost, "Change", d.initial_cost - d.final_cost);
  r += fmt("\nMinimizer iterations %16d\nSuccessful steps %20d\nUnsuccessful steps %18d\n",
           d.num_successful_steps + d.num_unsuccessful_steps, d.num_successful_steps, d.num_unsuccessful_steps);
  r += fmt("\nTime (in seconds):\nPreprocessor %24.6f\n", d.preprocessor_time_in_seconds);
  r += fmt("\n  Residual only evaluation (%lld)\n  Jacobian & residual evaluation (%lld)\n  Linear solver (%lld), iterations %lld\n",
           (long long)d.num_residual_evaluations, (long long)d.num_jacobian_evaluations, (long long)d.num_linear_solves,
           (long long)d.total_linear_solver_iterations);
  r += fmt("Minimizer %27.6f\n\nTotal %31.6f\n", d.minimizer_time_in_seconds, d.total_time_in_seconds);
  r += fmt("\nKernel launches %21lld\n", (long long)d.num_kernel_launches);
  r += fmt("\nTermination: %25s (%s)\n", termination_name(d.termination_type), s->message.c_str());
  s->full = r;
}

}  // namespace sk
