// ba_layout_device.cuh — device-resident bundle-adjustment layout, built by kernels (ba_layout_device.cu).
#pragma once
#include <vector>

#include "ba_layout.h"
#include "common.cuh"

namespace sk {

// The arrays of BaLayoutHost that the solver keeps on the device, already there.  BaSolver adopts them instead of uploading.
struct BaLayoutDevice {
  bool valid = false;
  DBuf<int> tile_obs, tile_pt, tile_seg, tile_np, pt_ptr, seg_ptr, seg_cam, cam_seg_ptr, cam_seg, seg_pos;
  DBuf<unsigned short> obs_slot, obs_ptl, seg_perm;
  DBuf<double> obs;                       // [2 n] observed (x, y), input order (= sorted order)
  DBuf<long long> cam_off, pt_off;        // block offsets in the user's parameter array, ordered by offset
};

// offsets2: n interleaved (camera offset, point offset) pairs; obs_xy: n (x, y) pairs; host memory.
// Returns false -- nothing usable in *D, *H untouched in what matters -- when the problem is not of the shape this path covers
// (observations not sorted by (point, camera), a track longer than one tile, duplicate observations, overlapping or sparse
// block offsets) or SKERES_LAYOUT=host is set: the caller then runs the host builder, which handles or reports all of these.
// On success *H carries the counts, the per-tile maxima, the camera table (cam_offset) and the camera CSR (cam_seg_ptr); its
// per-observation arrays stay empty.
bool build_ba_layout_device(int64_t n, const int64_t* offsets2, const double* obs_xy, const std::vector<int64_t>* extra_cam_off,
                            cudaStream_t s, BaLayoutHost* H, BaLayoutDevice* D);

}  // namespace sk
