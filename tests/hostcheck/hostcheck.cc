// tests/hostcheck/hostcheck.cc — TEST HARNESS (not part of the product).
// Compiles the __host__ __device__ arithmetic of skeres_b200/csrc/jet.cuh and the host-side layout
// builder (ba_layout.cu has no kernels) with g++ so that they can be checked against the oracle on
// the CPU-only build host.  The product never runs this code path: libskeres.so calls the same
// functions from CUDA kernels only.
#include <cstdint>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../skeres_b200/csrc/ba_layout.h"
#include "../../skeres_b200/csrc/ba_tile_rec.h"
#include "../../skeres_b200/csrc/common.cuh"
#include "../../skeres_b200/csrc/jet.cuh"

extern "C" {

int hc_evaluate(int functor, const double* consts, const double* x, double* res, double* jac) {
  return sk::evaluate_functor(functor, consts, x, res, jac) ? 1 : 0;
}
void hc_snavely_residual(const double* cam, const double* pt, double ox, double oy, double* res) {
  sk::snavely_residual(cam, pt, ox, oy, res);
}
void hc_loss(int type, double a, double b, double s, double* rho) { sk::LossSpec l{type, a, b}; sk::loss_evaluate(l, s, rho); }
void hc_correct(int type, double a, double b, int nrow, int ncol, double* res, double* J) {
  double sq = 0; for (int i = 0; i < nrow; ++i) sq += res[i] * res[i];
  double rho[3]; sk::LossSpec l{type, a, b}; sk::loss_evaluate(l, sq, rho);
  sk::Corrector c(sq, rho);
  c.correct_jacobian(nrow, ncol, ncol, res, J);
  c.correct_residuals(nrow, res);
}
int hc_invert_spd3(const double* m6, double* inv6) { return sk::invert_spd3(m6, inv6) ? 1 : 0; }
int hc_invert_spd9(double* A, int n) { return sk::invert_spd<9>(A, n) ? 1 : 0; }

// Layout builder: returns 0 on success, else the sk_status; message in err (256 bytes).
struct HcLayout { sk::BaLayoutHost L; };
HcLayout* hc_layout_build(int64_t n, const int64_t* cam_off, const int64_t* pt_off, const double* obs, int rank, int world,
                          int* status, char* err) {
  HcLayout* h = new HcLayout;
  try { sk::build_ba_layout(n, cam_off, pt_off, obs, rank, world, &h->L); *status = 0; }
  catch (const sk::Error& e) { *status = e.status; std::strncpy(err, e.what(), 255); err[255] = 0; delete h; return nullptr; }
  return h;
}
// The rank-local ingestion of the multi-GPU path: one-GPU layout of a rank's own observations, with camera blocks that
// exist without a local observation (extra).
HcLayout* hc_layout_build_extra(int64_t n, const int64_t* cam_off, const int64_t* pt_off, const double* obs, int64_t n_extra,
                                const int64_t* extra, int* status, char* err) {
  HcLayout* h = new HcLayout;
  std::vector<int64_t> ex(extra, extra + n_extra);
  try { sk::build_ba_layout(n, cam_off, pt_off, obs, 0, 1, &h->L, 1, &ex); *status = 0; }
  catch (const sk::Error& e) { *status = e.status; std::strncpy(err, e.what(), 255); err[255] = 0; delete h; return nullptr; }
  return h;
}
void hc_layout_free(HcLayout* h) { delete h; }
void hc_layout_dims(const HcLayout* h, int32_t* out /*8*/) {
  const auto& L = h->L;
  out[0] = L.n_obs; out[1] = L.n_pts; out[2] = L.n_cams; out[3] = L.n_tiles; out[4] = L.n_segs; out[5] = L.max_seg_tile; out[6] = L.max_pt_tile;
  out[7] = L.input_was_sorted ? 1 : 0;
}
int32_t hc_layout_n_giant(const HcLayout* h) { return h->L.n_giant; }
int32_t hc_layout_n_chunks(const HcLayout* h) { return h->L.n_chunks; }
#define HC_COPY(name, field, T) void hc_layout_##name(const HcLayout* h, T* out) { std::memcpy(out, h->L.field.data(), sizeof(T) * h->L.field.size()); }
HC_COPY(perm, perm, int32_t) HC_COPY(obs_cam, obs_cam, int32_t) HC_COPY(obs_pt, obs_pt, int32_t) HC_COPY(pt_ptr, pt_ptr, int32_t)
HC_COPY(tile_obs, tile_obs, int32_t) HC_COPY(tile_pt, tile_pt, int32_t) HC_COPY(tile_seg, tile_seg, int32_t)
HC_COPY(obs_slot, obs_slot, uint16_t) HC_COPY(obs_ptl, obs_ptl, uint16_t) HC_COPY(seg_perm, seg_perm, uint16_t)
HC_COPY(seg_ptr, seg_ptr, int32_t) HC_COPY(seg_cam, seg_cam, int32_t) HC_COPY(cam_seg_ptr, cam_seg_ptr, int32_t) HC_COPY(cam_seg, cam_seg, int32_t)
HC_COPY(tile_np, tile_np, int32_t) HC_COPY(tile_chunk, tile_chunk, int32_t) HC_COPY(gp_tile_begin, gp_tile_begin, int32_t)
HC_COPY(gp_tile_count, gp_tile_count, int32_t) HC_COPY(gp_point, gp_point, int32_t)
HC_COPY(cam_offset, cam_offset, int64_t) HC_COPY(pt_offset, pt_offset, int64_t)
void hc_layout_obs(const HcLayout* h, double* out) { std::memcpy(out, h->L.obs_src, sizeof(double) * 2 * (size_t)h->L.n_obs); }   // caller's array still alive

// Two-level sums of the implicit-Schur product (ba_tile_rec.h): runs the SAME per-item functions the kernels call, item by
// item in place of thread by thread, over the records of every regular tile with pseudo-random staged values, and compares
// with direct sums over the layout's own point / segment lists.  out[0] = max relative error of a point sum, out[1] = of a
// segment sum, out[2] = number of sums checked, out[3] = 1 if every position lies in exactly one chunk of the right owner.
void hc_check_two_level_sums(const HcLayout* h, uint64_t seed, double* out) {
  const auto& H = h->L;
  const int T = sk::kTileObs;
  sk::TileRecDims D; std::vector<unsigned char> rec;
  sk::build_tile_records(H, &D, &rec);
  auto rnd = [&seed]() { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return (double)((seed >> 11) & 0xfffff) / 1048576.0 - 0.5; };
  double e_pt = 0, e_seg = 0, n = 0; bool cover = true;
  std::vector<int> seg_pos((size_t)H.n_segs);
  for (int t = 0; t < H.n_segs; ++t) seg_pos[H.cam_seg[t]] = t;
  std::vector<double> w(3 * T), pw(3 * T), vs((size_t)T * sk::kSegRow), ps((size_t)sk::seg_chunk_scratch(std::max(H.max_seg_tile, 1))), v(9 * T);
  for (int t = 0; t < H.n_tiles; ++t) {
    if (H.tile_chunk[t] >= 0) continue;
    const sk::RecView R = sk::rec_view(rec.data() + (size_t)t * D.stride, D.sp, D.pp, D.sc);
    const int ob = H.tile_obs[t], no = H.tile_obs[t + 1] - ob, np = H.tile_np[t], sb = H.tile_seg[t], ns = H.tile_seg[t + 1] - sb;
    for (int i = 0; i < no; ++i) {
      for (int k = 0; k < 3; ++k) w[i * 3 + k] = rnd();
      for (int k = 0; k < 9; ++k) { v[i * 9 + k] = rnd(); vs[(size_t)R.srank[i] * sk::kSegRow + k] = v[i * 9 + k]; }
    }
    // chunk tables cover every position once
    std::vector<int> seen(no, 0);
    for (int p = 0; p < np; ++p)
      for (int c = R.pcptr[p]; c < R.pcptr[p + 1]; ++c) {
        const int st = R.pchunk[c] & 255, len = (R.pchunk[c] >> 8) + 1;
        if (len > sk::kPtChunk || st < R.pptr[p] || st + len > R.pptr[p + 1]) cover = false;
        for (int j = st; j < st + len; ++j) seen[j]++;
      }
    for (int i = 0; i < no; ++i) if (seen[i] != 1) cover = false;
    std::fill(seen.begin(), seen.end(), 0);
    for (int s = 0; s < ns; ++s)
      for (int c = R.scptr[s]; c < R.scptr[s + 1]; ++c) {
        const int st = R.schunk[c] & 255, len = (R.schunk[c] >> 8) + 1;
        if (len > sk::kSegChunk || st < R.sptr[s] || st + len > R.sptr[s + 1]) cover = false;
        for (int j = st; j < st + len; ++j) seen[j]++;
      }
    for (int i = 0; i < no; ++i) if (seen[i] != 1 || R.srank[R.sperm[i]] != i) cover = false;
    for (int s = 0; s < ns; ++s) if (R.spos[s] != seg_pos[sb + s] || R.scam[s] != H.seg_cam[sb + s]) cover = false;   // camera-major output row
    // point sums
    const int n3 = 3 * R.pcptr[np];
    for (int idx = 0; idx < n3; ++idx) pw[idx] = sk::point_chunk_sum(R, w.data(), idx);
    for (int p = 0; p < np; ++p) {
      double a[3]; sk::point_combine(R, pw.data(), p, a[0], a[1], a[2]);
      for (int k = 0; k < 3; ++k) {
        double ref = 0, mag = 1e-300;
        for (int j = H.pt_ptr[H.tile_pt[t] + p] - ob; j < H.pt_ptr[H.tile_pt[t] + p + 1] - ob; ++j) { ref += w[j * 3 + k]; mag += std::fabs(w[j * 3 + k]); }
        e_pt = std::max(e_pt, std::fabs(a[k] - ref) / mag); n += 1;
      }
    }
    // segment sums
    const int n9 = 9 * R.scptr[ns];
    for (int idx = 0; idx < n9; ++idx) ps[idx] = sk::seg_chunk_sum(R, vs.data(), idx);
    for (int idx = 0; idx < 9 * ns; ++idx) {
      const int s = idx / 9, k = idx % 9;
      double ref = 0, mag = 1e-300;
      for (int q = H.seg_ptr[sb + s]; q < H.seg_ptr[sb + s + 1]; ++q) { const double x = v[(size_t)H.seg_perm[q] * 9 + k]; ref += x; mag += std::fabs(x); }
      e_seg = std::max(e_seg, std::fabs(sk::seg_combine(R, ps.data(), idx) - ref) / mag); n += 1;
    }
  }
  out[0] = e_pt; out[1] = e_seg; out[2] = n; out[3] = cover ? 1.0 : 0.0;
}
}
