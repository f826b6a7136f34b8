// ba_dev.cuh — device view of the bundle-adjustment layout (ba_layout.h).
// Plain structs only: this header is also handed to NVRTC, verbatim, when a functor given as source is
// compiled into the tile evaluation kernel at run time (user_functor.cu).
#pragma once
#include "ba_tile_obs.h"

namespace sk {

// Device view of BaLayoutHost (plain pointers, passed to kernels by value).
struct BaDev {
  int n_obs, n_pts, n_cams, n_tiles, n_segs, max_seg_tile, max_pt_tile;
  int n_giant, n_chunks;               // tracks longer than one tile and the chunk tiles they are cut into (ba_layout.h)
  const int* tile_obs; const int* tile_pt; const int* tile_seg; const int* pt_ptr;
  const int* tile_np;                  // [T] > 0: points of a regular tile;  < 0: chunk tile, ordinal = -tile_np - 1
  const int* gp_tile_begin; const int* gp_tile_count; const int* gp_point;   // [n_giant]
  // per-tile metadata records for the prefetching matvec (ba_kernels.cu: RecView); nullptr when not built
  const unsigned char* tile_rec; int rec_stride, rec_sp, rec_pp, rec_sc;
  int matvec_classic;                  // 0: k_ba_matvec_tma / the fused PCG solve; 1: k_ba_matvec (SKERES_MATVEC=classic); read per solver
  int matvec_serial_sums;              // 1 (default): per-point / per-segment sums as one serial chain each; 0: the chunked
                                       // two-level sums (SKERES_MATVEC_SUMS=chunked)
  int l2_keep_tiles;                   // implicit-Schur product: tiles [0, l2_keep_tiles) are copied with the L2 evict_last policy, the
                                       // rest evict_first (the stored Jacobian does not change during a linear solve: this part of it
                                       // stays in the 126 MB L2 from one product to the next); < 0: no cache hints
  const unsigned short* obs_slot; const unsigned short* obs_ptl; const unsigned short* seg_perm;
  const int* seg_ptr; const int* seg_cam; const int* cam_seg_ptr; const int* cam_seg;
  const int* seg_pos;                  // [S] inverse of cam_seg: the implicit-Schur product stores a segment's partial at its camera-major position
  const double2* obs;   // [n_obs] observed (x, y)
};

// Stored Jacobian: 12 planes of double2, plane k at J2 + k * n_obs.
//   planes 0..8 : (F[0][k], F[1][k])  d res / d camera parameter k
//   planes 9..11: (E[0][k], E[1][k])  d res / d point coordinate k
// i.e. 192 bytes per observation, every access a coalesced 16-byte vector.
constexpr int kJPlanes = 12;

}  // namespace sk
