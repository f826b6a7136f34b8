// lm_kernels.cuh — device-resident trust-region / LM / PCG state and the small kernels around it.
//
// All step-control decisions (TrustRegionMinimizer + LevenbergMarquardtStrategy, SURVEY.md A.3/A.4;
// ConjugateGradientsSolver, A.7) are taken ON THE DEVICE by single-thread "scalar" kernels that read
// reduced sums from `sbuf`; the host only enqueues a fixed kernel sequence whose members are
// guarded by flags in LmDev, and reads the state block back once per LM iteration.
#pragma once
#include "common.cuh"

namespace sk {

enum Slot {                 // sbuf layout (doubles).  "pt" slots are summed across ranks (points are
  SB_COST = 0,              // partitioned); "cam" slots are identical on every rank.
  SB_FLAG_EVAL = 1,         // != 0: an evaluation failed on some rank (every rank must take the same branch: a rank that
                            // terminated alone would leave the others waiting in the next collective)
  SB_GRAD_SQ_PT = 2,        // slots [0,4) are fresh after a Jacobian evaluation,
  SB_XNORM_SQ_PT = 3,
  SB_MCC = 4,               // slots [4,7) after the linear solve + candidate,
  SB_STEP_SQ_PT = 5,        // slots [0,2) after the candidate cost.
  SB_FLAG_LIN = 6,          // 1 per rank whose linear solve failed numerically + kFatalFlag per rank whose solve failed fatally
  SB_SUM_COUNT = 7,         // number of slots that go through the sum-allreduces
  SB_STEP_SQ_CAM = 7,
  SB_XNORM_SQ_CAM = 8,
  SB_GRAD_SQ_CAM = 9,
  SB_GRAD_MAX = 10,         // slots [10,12) are max-allreduced
  SB_TIME = 11,             // host wall-clock seconds since minimize() started, max over ranks (the time limit is then the
                            // same decision on every rank)
  SB_COUNT = 16
};
constexpr double kFatalFlag = 1024.0;

enum TermReason {
  TR_NONE = 0, TR_MAX_ITERATIONS, TR_GRADIENT_TOLERANCE, TR_MIN_RADIUS, TR_PARAMETER_TOLERANCE,
  TR_FUNCTION_TOLERANCE, TR_INVALID_STEPS, TR_LINEAR_SOLVER_FATAL, TR_EVALUATION_FAILED, TR_MAX_TIME
};

enum LinTerm { LIN_SUCCESS = 0, LIN_NO_CONVERGENCE = 1, LIN_FAILURE = 2, LIN_FATAL = 3 };

struct LmParams {           // by-value kernel argument (from sk_solver_options)
  int max_num_iterations, max_num_consecutive_invalid_steps;
  double max_radius, min_radius, min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
  double function_tolerance, gradient_tolerance, parameter_tolerance, eta, fixed_cost;
};

struct LmDev {
  // LevenbergMarquardtStrategy
  double radius, decrease_factor;
  int reuse_diagonal;
  // TrustRegionMinimizer
  int iteration, num_consecutive_invalid, num_successful, num_unsuccessful, num_rows;
  double x_cost, x_norm, minimum_cost, cand_cost, model_cost_change;
  // linear solver outcome of the current iteration
  int lin_iterations, lin_termination;
  // flow control of the current iteration (kernel guards: run when != 0)
  int g_eval_cand, g_accept, g_finalize;
  int terminate, termination_type, term_reason;
  double term_v1, term_v2;
  int eval_failed;            // sticky bits from the evaluator
  int lin_error;              // sticky bits from the Schur set-up (not positive definite)
  sk_iteration_summary row;   // the row being built
};

struct PcgDev {
  double rho, last_rho, pq, alpha, beta, Q0, Q1, norm_b;
  int iter;                   // ConjugateGradientsSolver summary.num_iterations
  int active;                 // guard: 1 while iterating
  int termination;            // LinTerm
  int pad_;                   // last FINISHED iteration (k_pcg_head)
  unsigned int done_count;    // CTAs of the current k_pcg_update / k_pcg_resid2 that have published their partials
  int pad2_;
};

struct PcgParams { int min_iterations, max_iterations; double q_tolerance; };

constexpr int kMaxPartials = 2048;   // upper bound on blocks of any partial-producing vector kernel

// ---- vector kernels over the state vector (n = 9C + 3P, camera part first: [0, nc)) --------------
int vec_blocks(int64_t n);
void launch_jacobi_scale(int64_t n, const double* cnorm2, double* scale, cudaStream_t s);
void launch_lm_diagonal(int64_t n, const double* cnorm2, double* diagonal, double* D, const LmDev* st, LmParams prm, cudaStream_t s);
void launch_grad_norms(int64_t n, int64_t nc, const double* x, const double* g, double* part /*[3][kMaxPartials]*/, const int* guard, cudaStream_t s);
void launch_candidate(int64_t n, int64_t nc, const double* x, const double* step, const double* scale, double* cand, double* part /*[2][kMaxPartials]*/, cudaStream_t s);
void launch_accept(int64_t n, int64_t nc, double* x, const double* cand, double* part /*[2][kMaxPartials]*/, const int* guard, cudaStream_t s);
void launch_negate(int64_t n, const double* in, double* out, cudaStream_t s);
void launch_fill(int64_t n, double value, double* out, cudaStream_t s);

// sbuf[slot] = fixed-order sum (or max) of part[0..n); one block per job, up to 8 jobs.
struct ReduceJob { const double* part; int n; int slot; int is_max; };
// flags: after the jobs (and whatever the guard says) block 0 publishes the sticky failure flags of `st` (+ the linear solver's
// fatal outcome: pcg termination, peer-window error word) into sbuf[SB_FLAG_EVAL] / sbuf[SB_FLAG_LIN], where the consumers
// read them after the scalar allreduce.
struct FlagSources { const LmDev* st; const PcgDev* pcg; const int* peer_error; };
void launch_reduce_jobs(const ReduceJob* jobs, int njobs, double* sbuf, const int* guard, FlagSources flags, cudaStream_t s);

// ---- scalar (single-thread) logic ---------------------------------------------------------------
void launch_lm_init(LmDev* st, double initial_radius, cudaStream_t s);
void launch_lm_iter0(LmDev* st, const double* sbuf, LmParams prm, cudaStream_t s);
void launch_lm_decide_a(LmDev* st, const PcgDev* pcg_or_null, const double* sbuf, LmParams prm, cudaStream_t s);
void launch_lm_decide_b(LmDev* st, const double* sbuf, LmParams prm, cudaStream_t s);
void launch_lm_post_accept(LmDev* st, const double* sbuf, LmParams prm, cudaStream_t s);
void launch_lm_finalize(LmDev* st, sk_iteration_summary* rows, int rows_capacity, LmParams prm, cudaStream_t s);

// The PCG kernels on camera vectors live in pcg_kernels.cu (fused: four launches per iteration).

}  // namespace sk
