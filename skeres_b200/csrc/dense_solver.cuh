// dense_solver.cuh — DENSE_QR back end of the LM driver for small generic problems.
#pragma once
#include <vector>
#include "dense_kernels.cuh"
#include "lm_solver.cuh"

namespace sk {

class DenseSolver : public LmSolver {
 public:
  // scalar_ptrs[j]: device address of the user scalar behind state-vector entry j (parameter blocks
  // may live in different DoubleArrays, as m and c do in CurveFitting.scala:103-106).
  DenseSolver(const sk_solver_options& opt, cudaStream_t stream, const std::vector<DenseRb>& rbs, int num_rows,
              const std::vector<double*>& scalar_ptrs, int num_param_blocks);

 protected:
  void eval_jacobian(bool scale_valid, bool store, const int* guard) override;
  void eval_cost(const double* xv, const int* guard) override;
  ReduceJob cost_job() override;
  ReduceJob linear_solve(const PcgDev** pcg_out) override;
  void load_state() override;
  void store_state() override;
  void fill_summary(sk_solver_summary_data* d) override;

 private:
  void evaluate(const double* xv, bool with_jacobian, const int* guard);
  int nrb_, m_, nparam_blocks_;
  std::vector<int> user_functors_;        // run-time compiled functors among the residual blocks (user_functor.cu), ascending
  DBuf<DenseRb> d_rbs_;
  DBuf<double*> d_ptrs_;
  DBuf<double> J_, b_, W_, block_cost_, mcc_part_;
};

}  // namespace sk
