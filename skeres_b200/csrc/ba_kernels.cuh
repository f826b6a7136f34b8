// ba_kernels.cuh — launch wrappers of the bundle-adjustment tile kernels (ba_kernels.cu).
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)

#include "ba_dev.cuh"
#include "ba_layout.h"
#include "ba_tile_rec.h"
#include "common.cuh"
#include "jet.cuh"

namespace sk {

struct Flags {           // device-resident control block shared by the LM / PCG kernels
  int skip;              // != 0: guarded kernels return immediately
  int error;             // sticky numerical-failure bits (E^T E or S_cc not positive definite)
};

// Residual (+ Jacobian) evaluation over all tiles.
//   x            state [9C + 3P]
//   scale        Jacobi column scaling [9C + 3P] or nullptr (= 1)
//   J2, r2       outputs (only when write_jacobian)
//   grad, cnorm2 outputs for the POINT part [9C .. 9C+3P) written directly; camera part goes to
//                seg_g / seg_n partials [S][9] (only when with_jacobian)
//   tile_cost    [n_tiles] partial costs
//   chunk_pt     [n_chunks][6] scratch: point-part partials of the chunk tiles of long tracks (with_jacobian)
//   guard        kernel returns immediately when *guard == 0 (nullptr = always run)
//   functor_id   the built-in SnavelyReprojectionError, or a functor of the same shape registered from source (user_functor.cuh)
void launch_ba_evaluate(const BaDev& L, const double* x, const double* scale, LossSpec loss, bool with_jacobian,
                        bool write_jacobian, double2* J2, double2* r2, double* grad, double* cnorm2,
                        double* seg_g, double* seg_n, double* tile_cost, double* chunk_pt, int* fail_flag, const int* guard,
                        cudaStream_t s, int functor_id = SK_FUNCTOR_SNAVELY_REPROJECTION_ERROR);

// out[c*K + k] = sum over the camera's segments of seg[s*K + k]   (deterministic, tile order)
void launch_cam_reduce(const BaDev& L, int K, const double* seg, double* out, const int* guard, cudaStream_t s);

// Schur set-up for one LM linear solve (SchurEliminator::Eliminate restricted to what the implicit
// solver needs): per point (E^T E + D_p^2)^-1 -> einv [P][6]; per (tile, camera) partials of the
// reduced right-hand side [S][9] and of the diagonal blocks of S (upper triangle, [S][45]).
void launch_ba_schur_setup(const BaDev& L, const double2* J2, const double2* r2, const double* D, double* einv,
                           double* seg_rhs, double* seg_M, int* error_flag, bool ftf_only, cudaStream_t s);

// Minv[c] = (sum of seg_M + D_c^2)^-1 as a full 9x9 row-major block (SchurJacobiPreconditioner).
// diag_only_identity: IDENTITY preconditioner (Minv = I).
void launch_ba_precond_invert(const BaDev& L, const double* M45 /*[C][45]*/, const double* D, double* Minv,
                              int* error_flag, cudaStream_t s);

// Implicit Schur product partials: seg_y[S][9] of  F^T (F p - E (E^T E)^-1 E^T F p), stored CAMERA-major (row seg_pos[s]): the
// partials of camera c are rows [cam_seg_ptr[c], cam_seg_ptr[c + 1]) -- contiguous for the second-level sums (pcg_kernels.cu)
struct PcgDev;
void launch_ba_matvec(const BaDev& L, const double2* J2, const double* p, const double* zdir, const PcgDev* pcg, const double* einv,
                      double* seg_y, const int* guard, cudaStream_t s, const CUtensorMap* tmapJ = nullptr);
// tmapJ: 2-D tensor map over the stored Jacobian (make_jacobian_tensor_map), or nullptr = one bulk copy per plane.
bool make_jacobian_tensor_map(const double2* J2, int n_obs, CUtensorMap* out);   // false when the problem is too small for the box

// Packs the per-tile metadata records (ba_tile_rec.h) on the device: rec = n_tiles * d.stride bytes, d = tile_rec_dims(...).
// Needs L.seg_pos and the layout arrays; byte-identical to the host builder.
void launch_ba_build_tile_records(const BaDev& L, unsigned char* rec, const TileRecDims& d, cudaStream_t s);

// Back-substitution + model cost change. z = reduced solution [9C] (not negated).
//   step[9C + 3p + k] = -y_p ; tile_mcc[t] = sum_i m_i . (r_i + m_i / 2), m = J * step
void launch_ba_back_substitute(const BaDev& L, const double2* J2, const double2* r2, const double* z,
                               const double* einv, double* step, double* tile_mcc, cudaStream_t s);

// Debug / test export: Jacobian of observation i in the Ceres layout (F 2x9 row-major, E 2x3).
void launch_ba_export_jacobian(const BaDev& L, const double2* J2, double* F /*[O][18]*/, double* E /*[O][6]*/, cudaStream_t s);

}  // namespace sk
