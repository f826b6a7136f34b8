// eval_abi.cuh — argument blocks of the generic (per-residual-block) evaluation kernels.  Plain structs: this header is also
// handed to NVRTC, verbatim, when a user-defined functor is compiled at run time (user_functor.cu), so that the kernels
// generated for it take exactly the arguments the built-in ones take.
#pragma once
#include "../../include/skeres.h"

namespace sk {

// AutoDiffCostFunction.evaluate for ONE residual block (k_evaluate_single / sk_user_evaluate_single).
struct EvalArgs {
  int functor; int has_jac;
  double consts[SK_MAX_CONSTS];
  const double* params[SK_MAX_PARAMETER_BLOCKS];
  double* jac[SK_MAX_PARAMETER_BLOCKS];
  double* residuals;
};

// One residual block of a generic (dense-path) problem, device side.
struct DenseRb {
  int functor, row, loss_type, pad_;
  double loss_a, loss_b;
  double consts[SK_MAX_CONSTS];
  int col[SK_MAX_PARAMETER_BLOCKS];     // first column of each parameter block in the state vector
};

}  // namespace sk
