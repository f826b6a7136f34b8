// ba_tile.cuh — the tile header every tile kernel starts from.  Per translation unit (anonymous namespace), device code only;
// handed to NVRTC with ba_evaluate.cuh (user_functor.cu).
#pragma once
#include "ba_dev.cuh"

namespace sk {
namespace {

constexpr int T = kTileObs;
constexpr int VLD = T + 1;   // padded leading dimension of the per-observation staging planes

struct Tile { int ob, no, pb, np, sb, ns, chunk; };   // chunk >= 0: chunk tile of a long track (np == 1), else -1

__device__ __forceinline__ Tile load_tile(const BaDev& L, int t) {
  Tile q;
  q.ob = L.tile_obs[t]; q.no = L.tile_obs[t + 1] - q.ob;
  q.pb = L.tile_pt[t];
  const int np = L.tile_np[t];
  q.np = np < 0 ? 1 : np; q.chunk = np < 0 ? -np - 1 : -1;
  q.sb = L.tile_seg[t]; q.ns = L.tile_seg[t + 1] - q.sb;
  return q;
}

}  // namespace
}  // namespace sk
