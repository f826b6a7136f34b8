// ba_tile_rec.h — per-tile metadata records of the implicit-Schur product and the two-level sums that read them.
// Host + device: the table builder (ba_layout.cu) and the per-item arithmetic of the sums are plain functions, so the
// CPU-only build host can check them against a direct evaluation (tests/hostcheck, tests/test_host_logic.py); the
// kernels (ba_kernels.cu) call the same functions with the item index taken from the thread index.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include "ba_layout.h"

#ifdef __CUDACC__
#define SK_HD __host__ __device__ __forceinline__
#else
#define SK_HD inline
#endif

namespace sk {

// Two-level sums: a point's observations are summed in chunks of kPtChunk, a (tile, camera) segment's in chunks of kSegChunk.
constexpr int kPtChunk = 4, kSegChunk = 8;
constexpr int kSegRow = 9;                 // row stride of the segment-ordered staging of v (odd: conflict-free scatter)

// Record of one tile (all starts relative to the tile's first observation):
//   u16 slot[T] | u16 ptl[T] | u16 sperm[T] | u16 srank[T] | i32 sptr[sp] | i32 pptr[pp] | i32 scam[sp] | i32 spos[sp]
//   | u16 pchunk[T] | u16 pcptr[pp] | u16 schunk[sc] | u16 scptr[sp]
// slot / ptl = tile-local segment and point of an observation; sperm = tile-local observation ids in (segment, point) order,
// srank its inverse; sptr / pptr = first position of every segment (in sperm order) / first observation of every point;
// scam = camera of a segment, spos = where the segment's partial sum goes: its position in the CAMERA-major list of all
// segments (BaLayoutHost::cam_seg), so that the partials of one camera are contiguous for the second-level sums;
// pchunk / schunk = chunks of the two-level sums (first position | (length - 1) << 8);
// pcptr / scptr = first chunk of every point / segment.
struct TileRecDims { int stride, sp, pp, sc; };
struct RecView {
  const unsigned short* slot; const unsigned short* ptl; const unsigned short* sperm; const unsigned short* srank;
  const int* sptr; const int* pptr; const int* scam; const int* spos;
  const unsigned short* pchunk; const unsigned short* pcptr; const unsigned short* schunk; const unsigned short* scptr;
};
SK_HD RecView rec_view(const unsigned char* base, int sp, int pp, int sc) {
  constexpr int T = kTileObs;
  RecView r;
  r.slot = reinterpret_cast<const unsigned short*>(base);
  r.ptl = r.slot + T; r.sperm = r.ptl + T; r.srank = r.sperm + T;
  r.sptr = reinterpret_cast<const int*>(base + 8 * T);
  r.pptr = r.sptr + sp; r.scam = r.pptr + pp; r.spos = r.scam + sp;
  r.pchunk = reinterpret_cast<const unsigned short*>(r.spos + sp);
  r.pcptr = r.pchunk + T; r.schunk = r.pcptr + pp; r.scptr = r.schunk + sc;
  return r;
}
inline TileRecDims tile_rec_dims(int max_seg_tile, int max_pt_tile) {
  const int T = kTileObs;
  TileRecDims d;
  d.sp = (max_seg_tile + 1 + 3) & ~3; d.pp = (max_pt_tile + 1 + 3) & ~3;
  d.sc = (T / kSegChunk + max_seg_tile + 1 + 7) & ~7;
  d.stride = (int)(((size_t)8 * T + 4 * ((size_t)3 * d.sp + d.pp) + 2 * ((size_t)T + d.pp + d.sc + d.sp) + 15) & ~(size_t)15);
  return d;
}
// Packs the records of all tiles (host, threaded).  out: n_tiles * dims.stride bytes.
void build_tile_records(const BaLayoutHost& H, TileRecDims* dims, std::vector<unsigned char>* out);

// ---- per-item arithmetic of the two-level sums ---------------------------------------------------------------------------
// Item idx = 3 c + k of the point sums: the sum of w[.][k] over chunk c (w staged [observation][3]).
SK_HD double point_chunk_sum(const RecView& R, const double* w, int idx) {
  const int c = idx / 3, k = idx - 3 * c;
  const unsigned d = R.pchunk[c];
  const int len = (int)(d >> 8) + 1;
  const double* b = w + (d & 255u) * 3 + k;
  const double x0 = b[0], x1 = len > 1 ? b[3] : 0.0, x2 = len > 2 ? b[6] : 0.0, x3 = len > 3 ? b[9] : 0.0;
  return (x0 + x1) + (x2 + x3);
}
// Point p: a = sum of its chunk sums pw[3 c + .] in chunk order.
SK_HD void point_combine(const RecView& R, const double* pw, int p, double& a0, double& a1, double& a2) {
  const int cb = R.pcptr[p], ce = R.pcptr[p + 1];
  a0 = 0.0; a1 = 0.0; a2 = 0.0;
  for (int c = cb; c < ce; ++c) { a0 += pw[3 * c]; a1 += pw[3 * c + 1]; a2 += pw[3 * c + 2]; }
}
// Item idx = 9 c + k of the segment sums: the sum of the segment-ordered v[.][k] over chunk c (v staged [position][kSegRow]).
SK_HD double seg_chunk_sum(const RecView& R, const double* vs, int idx) {
  const int c = idx / 9, k = idx - 9 * c;
  const unsigned d = R.schunk[c];
  const int len = (int)(d >> 8) + 1;
  const double* b = vs + (d & 255u) * kSegRow + k;
  const double x0 = b[0], x1 = len > 1 ? b[kSegRow] : 0.0, x2 = len > 2 ? b[2 * kSegRow] : 0.0, x3 = len > 3 ? b[3 * kSegRow] : 0.0;
  const double x4 = len > 4 ? b[4 * kSegRow] : 0.0, x5 = len > 5 ? b[5 * kSegRow] : 0.0, x6 = len > 6 ? b[6 * kSegRow] : 0.0,
               x7 = len > 7 ? b[7 * kSegRow] : 0.0;
  return ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
// Item idx = 9 s + k: the sum of segment s's chunk sums ps[9 c + k] in chunk order.
SK_HD double seg_combine(const RecView& R, const double* ps, int idx) {
  const int s = idx / 9, k = idx - 9 * s;
  const int cb = R.scptr[s], ce = R.scptr[s + 1];
  double sum = 0.0;
  for (int c = cb; c < ce; ++c) sum += ps[9 * c + k];
  return sum;
}
// doubles of the chunk-sum scratch ps of one tile (an even count)
SK_HD int seg_chunk_scratch(int max_seg_tile) { return ((kTileObs / kSegChunk + max_seg_tile) * 9 + 1) & ~1; }

}  // namespace sk
