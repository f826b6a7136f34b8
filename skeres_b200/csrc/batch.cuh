// batch.cuh — batched independent CurveFitting-shaped problems (see batch.cu).
#pragma once
#include "common.cuh"
#include "lm_solver.cuh"

namespace sk {

// x, y: device [n_obs][n_problems]; mc: device [2][n_problems] (in: start, out: solution).
void curve_fit_batch_solve(const sk_solver_options& opt, int64_t n_problems, int n_obs, const double* x, const double* y,
                           double* mc, double* out_initial_cost, double* out_final_cost, int32_t* out_num_iterations,
                           int32_t* out_termination_type, sk_solver_summary* summary);

}  // namespace sk
