// lm_solver.cuh — host driver of the device-resident trust-region Levenberg–Marquardt loop
// (TrustRegionMinimizer::Minimize).  Back ends (bundle adjustment with Schur
// solvers, small dense problems with DENSE_QR) plug in evaluation and the linear solve.
#pragma once
#include <string>
#include <vector>

#include "comm.cuh"
#include "common.cuh"
#include "lm_kernels.cuh"

struct sk_solver_summary {
  sk_solver_summary_data data{};
  std::vector<sk_iteration_summary> rows;
  std::string message, brief, full;
};

namespace sk {

class LmSolver {
 public:
  LmSolver(const sk_solver_options& opt, cudaStream_t stream);
  virtual ~LmSolver();
  // Runs the minimisation; state vector x must have been loaded by the back end.
  void minimize(sk_solver_summary* summary, int max_num_iterations_override = -1);
  // sk_solver_time_schur_product: mean device ms of one application of the back end's iterative linear operator
  virtual double time_linear_operator(int /*reps*/) { throw Error(SK_ERR_UNSUPPORTED, "this solver has no iterative linear operator to time"); }

 protected:
  // --- back-end hooks -----------------------------------------------------------------------------
  // Residuals + Jacobian at x: fills the back end's Jacobian storage, grad[0..n), cnorm2[0..n) (squared
  // column norms of the column-scaled Jacobian) and a cost partial array. `scale_valid` is false for the
  // first pass of iteration 0 (scaling not yet known: norms of the unscaled Jacobian are wanted).
  virtual void eval_jacobian(bool scale_valid, bool store, const int* guard) = 0;
  virtual void eval_cost(const double* xv, const int* guard) = 0;      // residual-only at xv
  virtual ReduceJob cost_job() = 0;                                    // where the cost partials are
  // Solves the LM system for the current D; writes step[0..n) (already negated) and returns the job
  // that sums to  sum_i m_i.(r_i + m_i/2)  with m = J*step.  Must leave lin_iterations/termination in st.
  virtual ReduceJob linear_solve(const PcgDev** pcg_out) = 0;
  virtual void load_state() = 0;       // user parameter arrays -> x
  virtual void store_state() = 0;      // x -> user parameter arrays
  virtual void fill_summary(sk_solver_summary_data* d) = 0;
  virtual void note_linear_iterations(int /*iterations*/) {}   // after the readback of an LM iteration: what its linear solve took

  void allocate(int64_t n, int64_t nc);
  void reduce(std::initializer_list<ReduceJob> jobs, const int* guard);
  void allreduce_after_jacobian(double t_start);   // multi-GPU: sum / max of the point-partitioned sbuf slots, flags, elapsed time
  // where the reduce kernel finds the linear solver's fatal outcome (set by the back end; nullptr = none)
  const PcgDev* flag_pcg_ = nullptr;
  const int* flag_peer_error_ = nullptr;
  HBuf<double> time_h_;                // [0] elapsed seconds written by this rank, [1] the maximum over the ranks read back

  sk_solver_options opt_;
  LmParams prm_{};
  cudaStream_t stream_;
  sk_comm* comm_ = nullptr;
  Profiler prof_;
  int64_t n_ = 0, nc_ = 0;
  DBuf<double> x_, cand_, step_, scale_, grad_, cnorm2_, diagonal_, D_, sbuf_;
  DBuf<double> part_a_, part_b_;       // [3][kMaxPartials] partial arrays of the vector kernels
  DBuf<LmDev> st_;
  HBuf<LmDev> st_h_;
  DBuf<sk_iteration_summary> rows_;
  int rows_cap_ = 0;
  cudaEvent_t ev_start_ = nullptr, ev_stop_ = nullptr;
  int64_t n_res_evals_ = 0, n_jac_evals_ = 0, n_lin_solves_ = 0, n_lin_iters_ = 0;
};

// Message text of Solver::Summary::message for a termination reason (mirrors Ceres' wording).
std::string termination_message(const LmDev& st, const sk_solver_options& opt);
void format_reports(sk_solver_summary* s);

}  // namespace sk
