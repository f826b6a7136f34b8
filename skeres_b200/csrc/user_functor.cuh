// user_functor.cuh — cost functors supplied as CUDA source and compiled at run time (NVRTC) over the device Jet<N>:
// the device counterpart of a Scala `class F extends CostFunctor(kNumResiduals, N0, N1, ...) { def apply[T](x: Array[T]*) }`
// (core/.../CostFunctor.scala:31-51), which the reference evaluates on the JVM through ~45 JNI crossings per residual block.
#pragma once
#include "ba_dev.cuh"
#include "eval_abi.cuh"
#include "jet.cuh"

namespace sk {

constexpr int kUserFunctorBase = 1000;      // ids >= this are run-time compiled functors

// Compiles `source` (which must define `template <class T> __device__ bool NAME(const double* consts, T const* const* x, T* residuals)`)
// into the two generic evaluation kernels for the given sizes and registers it.  No device is needed to compile; the module is
// loaded on the current device at first use.  Throws sk::Error (SK_ERR_INVALID_ARGUMENT with the compiler log on a compile error).
int register_user_functor(const char* name, const char* source, int nres, int nblk, const int* sizes, int nconsts);
bool user_functor_info(int id, FunctorInfo* out);
// AutoDiffCostFunction.evaluate for one residual block of functor `id` (same contract as the built-in k_evaluate_single).
void launch_user_evaluate_single(int id, const EvalArgs& a, int* ok_out, cudaStream_t s);
// Residual blocks of functor `id` among rbs[0, nrb): residuals / Jacobian columns / corrected residuals as the built-in dense
// evaluation writes them; the CTA's cost is ADDED to block_cost[blockIdx.x] (the built-in kernel has written the slot before).
void launch_user_dense_evaluate(int id, bool with_jacobian, int nrb, const DenseRb* rbs, const double* x, double* J, int m, double* b,
                                double* block_cost, int* fail_flag, const int* guard, cudaStream_t s);
// A functor of the bundle-adjustment shape (2 residuals; parameter blocks of 9 and 3; 2 constants = the observation) is also
// compiled into the TILE evaluation kernel of the Schur solvers (ba_evaluate.cuh) and can stand in for the built-in
// SnavelyReprojectionError there: same arguments as k_ba_evaluate<with_jacobian> (ba_kernels.cu), smem = its dynamic shared memory.
bool user_functor_runs_on_tiles(int id);
void launch_user_ba_evaluate(int id, bool with_jacobian, size_t smem, const BaDev& L, const double* x, const double* scale, LossSpec loss,
                             int write_j, double2* J2, double2* r2, double* grad, double* cnorm2, double* seg_g, double* seg_n,
                             double* tile_cost, double* chunk_pt, int* fail_flag, const int* guard, cudaStream_t s);

}  // namespace sk
