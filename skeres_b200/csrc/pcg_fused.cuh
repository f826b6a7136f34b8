// pcg_fused.cuh — the whole PCG solve on the implicit Schur complement as one persistent cooperative kernel (pcg_fused.cu).
#pragma once
#include "ba_kernels.cuh"
#include "comm.cuh"
#include "lm_kernels.cuh"

namespace sk {

struct PcgSolveArgs {
  BaDev L;
  const double2* J2; const double* einv; double* seg_y;   // stored Jacobian, (E^T E + D^2)^-1 blocks, segment partials [S][9]
  const double* D; const double* Minv; const double* b;   // LM diagonal (camera part), preconditioner blocks (nullptr: identity), rhs
  double* x; double* p; double* r; double* z;              // PCG vectors [9 C]; x, r, z initialised by launch_pcg_begin
  double* part_pq; double* part_Q; double* part_rho;       // partial-sum slots, one per virtual block of 8 cameras
  PcgDev* st;                                              // scalar state, initialised by launch_pcg_begin
  PcgParams prm;
  int reset_period;                                        // r = b - S x recomputed every reset_period-th iteration (Ceres: 10)
  unsigned int* grid_bar;                                  // grid-barrier counter, zero at launch (launch_pcg_begin)
  unsigned long long* phase_ns;                            // [2] += ns spent in (products + exchange | vector phases), CTA 0's clock; or nullptr
  PeerWindow win;                                          // world > 1: the per-product exchange runs inside the kernel
  unsigned long long seq_base;                             // sequence number of the last exchange before this solve
};

// True when the fused solve can run this problem on the current device (no long tracks, records built, shared memory fits,
// cooperative launch available).
bool pcg_solve_supported(const BaDev& L, bool have_tmap);
// One launch = one linear solve.  Exchanges performed: iterations + iterations / reset_period (the caller advances its sequence
// number by that once the iteration count is known).
void launch_pcg_solve(const PcgSolveArgs& args, const CUtensorMap* tmapJ, cudaStream_t s);

}  // namespace sk
