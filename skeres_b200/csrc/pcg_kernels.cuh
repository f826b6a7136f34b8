// pcg_kernels.cuh — fused PCG kernels on camera vectors (see pcg_kernels.cu).
#pragma once
#include "ba_kernels.cuh"
#include "comm.cuh"
#include "lm_kernels.cuh"

namespace sk {

int pcg_blocks(int n_cams);   // CTAs (= partial sums) of the warp-per-camera kernels
// win != nullptr (multi-GPU peer window, comm.cuh): the reduced vector is written into this rank's window slot of exchange `seq`
// and published to every rank; launch_pcg_reduce / launch_pcg_resid2 with the same (win, seq) wait for all ranks and add their
// contributions in rank order.  win == nullptr: plain output into y (NCCL path / single GPU).
void launch_cam_reduce9_warp(const BaDev& L, const double* seg_y, double* y, const int* guard, cudaStream_t s,
                             const PeerWindow* win = nullptr, unsigned long long seq = 0);
void launch_pcg_begin(int n_cams, const double* rhs, const double* Minv, double* x, double* r, double* z, double* part_bb,
                      double* part_rho, PcgDev* st, int* lin_error, const double* global_lin_flag, unsigned int* grid_bar, cudaStream_t s);
// global_lin_flag: multi-GPU, != 0 when the Schur set-up failed on some rank (then it failed for all: *lin_error is raised).
// grid_bar: the fused solve's grid-barrier counter, zeroed here (nullptr: kernel sequence only).
void launch_pcg_head(PcgDev* st, const double* part_rho, const double* part_pq, const double* part_Q, int nparts, PcgParams prm,
                     int finish_only, cudaStream_t s);
void launch_pcg_reduce(const BaDev& L, const double* seg_y, const double* y_in, const double* D, double* z, double* p, double* part_pq,
                       const PcgDev* st, cudaStream_t s, const PeerWindow* win = nullptr, unsigned long long seq = 0);
void launch_pcg_update(int n_cams, const double* Minv, const double* b, double* x, const double* p, double* r, double* z,
                       const double* part_pq, int recompute, double* part_Q, double* part_rho, PcgDev* st, PcgParams prm, cudaStream_t s);
// update (without recompute) and resid2 also do what launch_pcg_head would do next: finish the iteration, open the next one.
void launch_pcg_resid2(const BaDev& L, const double* seg_y, const double* y_in, const double* D, const double* Minv, const double* b,
                       const double* x, double* r, double* z, double* part_Q, double* part_rho, PcgDev* st, const double* part_pq,
                       PcgParams prm, cudaStream_t s, const PeerWindow* win = nullptr, unsigned long long seq = 0);

}  // namespace sk
