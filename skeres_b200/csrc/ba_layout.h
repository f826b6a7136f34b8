// ba_layout.h — HBM data layout of a bundle-adjustment problem (the SchurEliminator<2,3,9> shape).
//
// Observations (SimpleBundleAdjuster.scala:139-145: one residual block per observation, blocks
// (camera 9, point 3)) are sorted by (point, camera) and cut into TILES of at most kTileObs
// observations that contain whole points.  One CTA processes one tile:
//   * per-point reductions (E^T E, E^T r, back-substitution) never leave the tile;
//   * per-camera reductions are done in two levels without atomics: inside the tile the
//     observations of one camera form a SEGMENT (tile-local, camera-sorted permutation) that is
//     summed in a fixed order into one partial per (tile, camera); a second kernel sums the partials
//     of each camera in tile order (cam_seg CSR).  Deterministic by construction.
// State vector layout (device): x[9*C + 3*P], cameras first — the BalProblem layout
// (SimpleBundleAdjuster.scala:28-33), cameras/points ordered by their offset in the user array.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "ba_tile_obs.h"   // kTileObs

namespace sk {

struct BaLayoutHost {
  int32_t n_obs = 0, n_pts = 0, n_cams = 0, n_tiles = 0, n_segs = 0;
  int32_t max_seg_tile = 0, max_pt_tile = 0;
  bool input_was_sorted = true;
  std::vector<int64_t> cam_offset;     // [C] offset of each camera block in the user array
  std::vector<int64_t> pt_offset;      // [P]   this rank's points
  std::vector<int64_t> all_pt_offset;  // every point of the problem, sorted (filled only when world_size > 1)
  std::vector<int32_t> perm;           // [O] sorted position -> original residual-block index
  std::vector<double> obs;             // [2*O] sorted (x, y)
  const double* obs_src = nullptr;     // sorted (x, y): obs.data(), or -- input already sorted, one rank -- the CALLER's array
                                       // (no copy; valid only while the caller's array lives: BaSolver uploads it at once)
  std::vector<int32_t> obs_cam;        // [O] camera id (sorted order)
  std::vector<int32_t> obs_pt;         // [O] point id (sorted order)
  std::vector<int32_t> pt_ptr;         // [P+1]
  std::vector<int32_t> tile_obs, tile_pt, tile_seg;  // [T+1]  (tile_pt[t] = first point of tile t)
  std::vector<int32_t> tile_np;        // [T] points of tile t (regular: whole points; chunk of a long track: 1)
  // Tracks longer than kTileObs are cut into consecutive CHUNK tiles (tile_chunk >= 0). Chunk tiles behave like any
  // tile on the camera side (segments, slots); their per-point sums are formed by one CTA per long track that walks
  // the chunk tiles [gp_tile_begin[g], gp_tile_begin[g] + gp_tile_count[g]) of point gp_point[g].
  std::vector<int32_t> tile_chunk;     // [T] -1 for a regular tile, else the ordinal of the chunk among all chunk tiles
  std::vector<int32_t> gp_tile_begin;  // [G]
  std::vector<int32_t> gp_tile_count;  // [G]
  std::vector<int32_t> gp_point;       // [G]
  int32_t n_giant = 0, n_chunks = 0;
  std::vector<uint16_t> obs_slot;      // [O] tile-local segment id
  std::vector<uint16_t> obs_ptl;       // [O] tile-local point id
  std::vector<uint16_t> seg_perm;      // [O] tile-local obs ids in (segment, obs) order
  std::vector<int32_t> seg_ptr;        // [S+1] positions into seg_perm
  std::vector<int32_t> seg_cam;        // [S]
  std::vector<int32_t> cam_seg_ptr;    // [C+1]
  std::vector<int32_t> cam_seg;        // [S]
};

// cam_off / pt_off: per-observation block offsets inside the user's parameter array.
// Restricts the layout to the point range [pt_begin_rank, pt_end_rank) of the *sorted* point list
// when world_size > 1 (point partition balanced by observation count); cameras are replicated.
// Throws sk::Error on unsupported structure (duplicate (camera, point) pairs, tracks longer than
// overlapping blocks).
void build_ba_layout(int64_t n, const int64_t* cam_off, const int64_t* pt_off, const double* obs_xy,
                     int rank, int world_size, BaLayoutHost* out, int64_t offset_stride = 1,
                     const std::vector<int64_t>* extra_cam_off = nullptr);
// offset_stride: cam_off[i * stride] / pt_off[i * stride] -- 2 lets the builder read a problem's interleaved
// (camera, point) offset pairs in place.
// extra_cam_off: camera blocks that exist although no observation of this call uses them (declared with
// Problem::AddParameterBlock): they get ids like any camera, so ranks holding different observations agree on the table.

// Contiguous point ranges balanced by observation count (out_begin has world_size + 1 entries).
void partition_points(int64_t n_points, const int64_t* point_ptr, int world_size, int64_t* out_begin);

}  // namespace sk
