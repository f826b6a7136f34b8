// ba_tile_obs.h — the one constant the host layout builder, the tile kernels and the run-time compiled tile kernels
// (user_functor.cu hands this header to NVRTC) must agree on.
#pragma once
namespace sk {
constexpr int kTileObs = 256;   // observations per tile == threads per CTA
}  // namespace sk
