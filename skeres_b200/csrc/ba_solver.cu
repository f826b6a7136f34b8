// ba_solver.cu — bundle-adjustment back end: SchurEliminator<2,3,9>-shaped problems solved with
// ITERATIVE_SCHUR (implicit Schur complement + SCHUR_JACOBI / JACOBI-free IDENTITY PCG) or with the
// explicit reduced camera system (DENSE_SCHUR / SPARSE_SCHUR, dense Cholesky).
#include "ba_solver.cuh"

#include "host_parallel.h"

#include <chrono>
#include <thread>
#include <cstdlib>

#include "dense_kernels.cuh"
#include "pcg_fused.cuh"
#include "pcg_kernels.cuh"

namespace sk {

constexpr int kResidualResetPeriod = 10;   // ConjugateGradientsSolver: r = b - S x recomputed every 10th iteration
static bool lst_is_explicit(int lst) { return lst == SK_DENSE_SCHUR || lst == SK_SPARSE_SCHUR; }

namespace {
__global__ void k_store_flag(const int* flag, double* out) { *out = *flag ? 1.0 : 0.0; }
__global__ void k_gather_blocks(int nblocks, int bsize, const long long* __restrict__ offsets, const double* __restrict__ user,
                                double* __restrict__ x) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nblocks * bsize) return;
  const int b = idx / bsize, k = idx - b * bsize;
  x[idx] = user[offsets[b] + k];
}
__global__ void k_scatter_blocks(int nblocks, int bsize, const long long* __restrict__ offsets, const double* __restrict__ x,
                                 double* __restrict__ user) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nblocks * bsize) return;
  const int b = idx / bsize, k = idx - b * bsize;
  user[offsets[b] + k] = x[idx];
}
}  // namespace

// L2 cache policies on the copies of the implicit-Schur product (ba_product.cuh: ProductPass::issue).  Measured
// (profiles/r02_v13_l2_policy.md): with the copies marked evict_first the operands of the vector phases -- the (tile, camera)
// partials, the preconditioner blocks, the camera-sized vectors -- survive the Jacobian stream in the 126 MB L2 as long as they
// fit: 2-4 % per PCG iteration at working sets of 26-61 MB (Venice shape: 18.1 -> 13.3 us of vector phases), 1.2-1.4 % LOSS at
// 104-190 MB (Final-13682 shape on one or two GPUs).  Keeping part of the Jacobian itself resident (evict_last) does not speed the
// product up (24 / 48 MB neutral, 80 / 110 MB slower).  SKERES_L2_KEEP_MB overrides: megabytes of tiles kept, negative = plain copies.
constexpr double kL2KeepMB = 24.0;
constexpr double kL2PolicyMaxWorkingSetMB = 80.0;

BaSolver::BaSolver(const sk_solver_options& opt, cudaStream_t stream, BaLayoutHost&& layout, double* user_params,
                   int64_t user_n, LossSpec loss, BaLayoutDevice* dev, int functor_id)
    : LmSolver(opt, stream), H_(std::move(layout)), user_(user_params), user_n_(user_n), loss_(loss), functor_(functor_id) {
  const auto& H = H_;
  cudaStream_t s = stream_;
  if (dev != nullptr && dev->valid) {
    // built on the device: nothing to upload
    d_tile_obs_ = std::move(dev->tile_obs); d_tile_pt_ = std::move(dev->tile_pt); d_tile_seg_ = std::move(dev->tile_seg);
    d_pt_ptr_ = std::move(dev->pt_ptr); d_obs_slot_ = std::move(dev->obs_slot); d_obs_ptl_ = std::move(dev->obs_ptl);
    d_seg_perm_ = std::move(dev->seg_perm); d_seg_ptr_ = std::move(dev->seg_ptr); d_seg_cam_ = std::move(dev->seg_cam);
    d_cam_seg_ptr_ = std::move(dev->cam_seg_ptr); d_cam_seg_ = std::move(dev->cam_seg); d_seg_pos_ = std::move(dev->seg_pos);
    d_obs_ = std::move(dev->obs); d_tile_np_ = std::move(dev->tile_np);
    d_cam_off_ = std::move(dev->cam_off); d_pt_off_ = std::move(dev->pt_off);
    d_gp_begin_.alloc(1); d_gp_count_.alloc(1); d_gp_point_.alloc(1);
    chunk_pt_.alloc(6);
  } else {
  d_tile_obs_.upload(H.tile_obs, s); d_tile_pt_.upload(H.tile_pt, s); d_tile_seg_.upload(H.tile_seg, s);
  d_pt_ptr_.upload(H.pt_ptr, s); d_obs_slot_.upload(H.obs_slot, s); d_obs_ptl_.upload(H.obs_ptl, s);
  d_seg_perm_.upload(H.seg_perm, s); d_seg_ptr_.upload(H.seg_ptr, s); d_seg_cam_.upload(H.seg_cam, s);
  d_cam_seg_ptr_.upload(H.cam_seg_ptr, s); d_cam_seg_.upload(H.cam_seg, s);
  std::vector<int> seg_pos((size_t)H.n_segs);
  for (int t = 0; t < H.n_segs; ++t) seg_pos[H.cam_seg[t]] = t;
  d_seg_pos_.upload(seg_pos, s);
  d_obs_.alloc((size_t)2 * H.n_obs); d_obs_.upload(H.obs_src, (size_t)2 * H.n_obs, s);   // possibly the caller's own array
  // device encoding of the tile kind: > 0 points of a regular tile, < 0 chunk tile of a long track (ordinal = -v - 1)
  std::vector<int> tile_np_enc((size_t)H.n_tiles);
  for (int t = 0; t < H.n_tiles; ++t) tile_np_enc[t] = H.tile_chunk[t] >= 0 ? -(H.tile_chunk[t] + 1) : H.tile_np[t];
  d_tile_np_.upload(tile_np_enc, s);
  d_gp_begin_.upload(H.gp_tile_begin, s); d_gp_count_.upload(H.gp_tile_count, s); d_gp_point_.upload(H.gp_point, s);
  chunk_pt_.alloc((size_t)6 * std::max(H.n_chunks, 1));
  std::vector<long long> co(H.cam_offset.begin(), H.cam_offset.end()), po(H.pt_offset.begin(), H.pt_offset.end());
  d_cam_off_.upload(co, s); d_pt_off_.upload(po, s);
  SK_CUDA(cudaStreamSynchronize(s));   // host vectors above are temporaries / pageable
  }
  L_.n_obs = H.n_obs; L_.n_pts = H.n_pts; L_.n_cams = H.n_cams; L_.n_tiles = H.n_tiles; L_.n_segs = H.n_segs;
  L_.max_seg_tile = std::max(H.max_seg_tile, 1); L_.max_pt_tile = std::max(H.max_pt_tile, 1);
  L_.tile_obs = d_tile_obs_.p; L_.tile_pt = d_tile_pt_.p; L_.tile_seg = d_tile_seg_.p; L_.pt_ptr = d_pt_ptr_.p;
  L_.obs_slot = d_obs_slot_.p; L_.obs_ptl = d_obs_ptl_.p; L_.seg_perm = d_seg_perm_.p; L_.seg_ptr = d_seg_ptr_.p;
  L_.seg_cam = d_seg_cam_.p; L_.cam_seg_ptr = d_cam_seg_ptr_.p; L_.cam_seg = d_cam_seg_.p; L_.seg_pos = d_seg_pos_.p;
  L_.obs = reinterpret_cast<const double2*>(d_obs_.p);
  L_.n_giant = H.n_giant; L_.n_chunks = H.n_chunks; L_.tile_np = d_tile_np_.p;
  L_.gp_tile_begin = d_gp_begin_.p; L_.gp_tile_count = d_gp_count_.p; L_.gp_point = d_gp_point_.p;
  L_.tile_rec = nullptr; L_.rec_stride = L_.rec_sp = L_.rec_pp = L_.rec_sc = 0;
  { const char* e = getenv("SKERES_MATVEC"); L_.matvec_classic = (e != nullptr && e[0] == 'c') ? 1 : 0; }
  // per-point / per-segment sums of the product: serial chains (default: measured 1.5 % faster back to back, r02) or the chunked
  // two-level sums (SKERES_MATVEC_SUMS=chunked)
  { const char* e = getenv("SKERES_MATVEC_SUMS"); L_.matvec_serial_sums = (e != nullptr && e[0] == 'c') ? 0 : 1; }
  L_.l2_keep_tiles = -1;               // set with the tile records (build_tile_records)
  const int64_t nc = (int64_t)9 * H.n_cams, n = nc + (int64_t)3 * H.n_pts;
  allocate(n, nc);
  J2_.alloc((size_t)2 * kJPlanes * std::max(H.n_obs, 1)); r2_.alloc((size_t)2 * std::max(H.n_obs, 1));
  einv_.alloc((size_t)6 * std::max(H.n_pts, 1));
  seg_a_.alloc((size_t)9 * std::max(H.n_segs, 1)); seg_b_.alloc((size_t)9 * std::max(H.n_segs, 1));
  tile_cost_.alloc(std::max(H.n_tiles, 1)); tile_mcc_.alloc(std::max(H.n_tiles, 1));
  tile_cost_.zero(s); tile_mcc_.zero(s);
  rhs_.alloc(nc + 1); px_.alloc(nc);   // rhs_[nc]: the ranks' summed Schur set-up failure flag (multi-GPU)
  pr_.alloc(nc); pp_.alloc(nc); pz_.alloc(nc); ybuf_.alloc(nc);
  pcg_.alloc(1); pcg_h_.alloc(1);
  pcg_part_.alloc(4 * kMaxPartials);
  pcg_.zero(s);
  if (comm_ && comm_->world > 1 && !(lst_is_explicit(opt.linear_solver_type))) peer_allreduce_create(comm_, (size_t)nc, s, &peer_);   // collective
  if (!lst_is_explicit(opt.linear_solver_type)) flag_pcg_ = pcg_.p;          // a fatal linear-solver outcome reaches every rank (lm_kernels.cuh: SB_FLAG_LIN)
  if (peer_.ok) {
    flag_peer_error_ = peer_.win.error;
    // which ranks contribute to which camera: bit r of cam_mask[c] = rank r holds observations of camera c (collective)
    std::vector<double> touch((size_t)H.n_cams, 0.0);
    for (int c = 0; c < H.n_cams; ++c) if (H.cam_seg_ptr[c + 1] > H.cam_seg_ptr[c]) touch[c] = (double)(1u << comm_->rank);
    DBuf<double> d_touch(touch.size());
    d_touch.upload(touch, s);
    comm_allreduce_sum(comm_, d_touch.p, touch.size(), s);         // disjoint bits: the sum is the union, exact in a double
    d_touch.download(touch.data(), touch.size(), s);
    SK_CUDA(cudaStreamSynchronize(s));
    std::vector<unsigned char> mask(touch.size());
    for (size_t c = 0; c < touch.size(); ++c) mask[c] = (unsigned char)(unsigned)touch[c];
    peer_.cam_mask.upload(mask, s);
    SK_CUDA(cudaStreamSynchronize(s));
    const char* e = getenv("SKERES_PEER_GATHER");                  // development: SKERES_PEER_GATHER=all reads every rank's window
    peer_.win.cam_mask = (e != nullptr && e[0] == 'a') ? nullptr : peer_.cam_mask.p;
    // virtual blocks of 8 cameras this rank contributes to: the fused solve skips the others when it reduces its own partials
    std::vector<unsigned char> own((size_t)cdiv(H.n_cams, 8), 0);
    for (int c = 0; c < H.n_cams; ++c) if ((mask[c] >> comm_->rank) & 1u) own[(size_t)c / 8] = 1;
    peer_.vb_own.upload(own, s);
    SK_CUDA(cudaStreamSynchronize(s));
    peer_.win.vb_own = peer_.win.cam_mask != nullptr ? peer_.vb_own.p : nullptr;   // without masks every rank's window is read everywhere
  }
  SK_REQUIRE(cdiv(H.n_cams, 8) <= kMaxPartials, SK_ERR_UNSUPPORTED, "more than %d cameras", kMaxPartials * 256 / 9);
  const int lst = opt.linear_solver_type;
  explicit_schur_ = (lst == SK_DENSE_SCHUR || lst == SK_SPARSE_SCHUR);
  if (!explicit_schur_) {
    const auto t0 = std::chrono::steady_clock::now();
    build_tile_records();
    {
      // L2 residency of the product's operands: SKERES_L2_KEEP_MB megabytes of tiles (Jacobian planes + record + (E^T E)^-1
      // blocks) are copied evict_last, the rest evict_first; negative: plain copies
      const char* e = getenv("SKERES_L2_KEEP_MB");
      const double working_set_mb = (72.0 * H.n_segs + (648.0 + 720.0) * H.n_cams) / 1e6;   // partials + M^-1 blocks + ten camera-sized vectors
      const double mb = e != nullptr ? atof(e) : (working_set_mb <= kL2PolicyMaxWorkingSetMB ? kL2KeepMB : -1.0);
      if (mb >= 0.0 && H.n_tiles > 0) {
        const double per_tile = (double)kJPlanes * kTileObs * 16 + L_.rec_stride + 48.0 * H.n_pts / H.n_tiles;
        L_.l2_keep_tiles = (int)std::min<double>(H.n_tiles, mb * 1e6 / per_tile);
      }
    }
    { const char* e = getenv("SKERES_MATVEC_TMAP");   // development: SKERES_MATVEC_TMAP=0 keeps one bulk copy per plane
      if (!(e && e[0] == '0')) have_tmapJ_ = make_jacobian_tensor_map(reinterpret_cast<const double2*>(J2_.p), H_.n_obs, &tmapJ_); }
    if (getenv("SKERES_TRACE_HOST")) fprintf(stderr, "[skeres] tile records: %.3f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  }
  if (!explicit_schur_) {
    SK_REQUIRE(opt.preconditioner_type == SK_SCHUR_JACOBI || opt.preconditioner_type == SK_JACOBI || opt.preconditioner_type == SK_IDENTITY,
               SK_ERR_UNSUPPORTED, "ITERATIVE_SCHUR supports the JACOBI, SCHUR_JACOBI and IDENTITY preconditioners on the device (got %d)",
               opt.preconditioner_type);
    seg_M_.alloc((size_t)45 * std::max(H.n_segs, 1)); M45_.alloc((size_t)45 * H.n_cams); Minv_.alloc((size_t)81 * H.n_cams);
    // The whole PCG loop as one persistent kernel (pcg_fused.cu) unless the problem has tracks longer than a tile, a development
    // variant of the product is selected, or (multi-GPU) the peer window is unavailable; SKERES_PCG=sequence forces the kernel
    // sequence (bit-identical with SKERES_PCG_WPC=1; measured 0.2303 ms per PCG iteration against 0.2211 fused, profiles/r02_v2_*).
    const char* e = getenv("SKERES_PCG");
    const bool multi = comm_ && comm_->world > 1;
    fused_pcg_ = !(e != nullptr && e[0] == 's') && L_.matvec_classic == 0 && L_.matvec_serial_sums == 1 && (!multi || peer_.ok) &&
                 pcg_solve_supported(L_, have_tmapJ_);
    grid_bar_.alloc(1); phase_ns_.alloc(2);
    grid_bar_.zero(s); phase_ns_.zero(s);
  } else {
    SK_REQUIRE(nc <= 16384, SK_ERR_UNSUPPORTED,
               "DENSE_SCHUR / SPARSE_SCHUR keep the reduced camera matrix dense on the device: at most 1820 cameras (got %d); use ITERATIVE_SCHUR",
               H.n_cams);
    SK_REQUIRE(comm_ == nullptr || comm_->world == 1, SK_ERR_UNSUPPORTED, "explicit Schur solvers run on one GPU; use ITERATIVE_SCHUR with a communicator");
    build_pair_lists();
    S_.alloc((size_t)nc * nc);
    seg_M_.alloc((size_t)45 * std::max(H.n_segs, 1)); M45_.alloc((size_t)45 * H.n_cams);
  }
  release_host_layout();
}

// Per-tile metadata records of the implicit-Schur product (ba_tile_rec.h), packed on the host and uploaded once.
void BaSolver::build_tile_records() {
  const char* e = getenv("SKERES_TILE_REC");               // development / tests: SKERES_TILE_REC=host packs them on the host
  if (e != nullptr && e[0] == 'h' && !H_.obs_slot.empty()) {
    TileRecDims dims; std::vector<unsigned char> rec;
    sk::build_tile_records(H_, &dims, &rec);
    d_tile_rec_.upload(rec, stream_);
    SK_CUDA(cudaStreamSynchronize(stream_));
    L_.tile_rec = d_tile_rec_.p; L_.rec_stride = dims.stride; L_.rec_sp = dims.sp; L_.rec_pp = dims.pp; L_.rec_sc = dims.sc;
    return;
  }
  const TileRecDims dims = tile_rec_dims(std::max(H_.max_seg_tile, 1), std::max(H_.max_pt_tile, 1));
  d_tile_rec_.alloc((size_t)std::max(H_.n_tiles, 1) * dims.stride);
  launch_ba_build_tile_records(L_, d_tile_rec_.p, dims, stream_);
  L_.tile_rec = d_tile_rec_.p; L_.rec_stride = dims.stride; L_.rec_sp = dims.sp; L_.rec_pp = dims.pp; L_.rec_sc = dims.sc;
}

void BaSolver::load_state() {
  n_real_matvecs_ = 0;
  if (phase_ns_.n) phase_ns_.zero(stream_);
  if (peer_.ok) SK_CUDA(cudaMemsetAsync(peer_.win.error, 0, sizeof(int), stream_));   // a time-out is sticky within one solve only
  KScope k(prof_, SK_KF_LM, 2);
  k_gather_blocks<<<cdiv((int64_t)L_.n_cams * 9, 256), 256, 0, stream_>>>(L_.n_cams, 9, d_cam_off_.p, user_, x_.p);
  if (L_.n_pts) k_gather_blocks<<<cdiv((int64_t)L_.n_pts * 3, 256), 256, 0, stream_>>>(L_.n_pts, 3, d_pt_off_.p, user_, x_.p + nc_);
  check_launch("k_gather_blocks");
}

void BaSolver::store_state() {
  if (comm_ && comm_->world > 1 && !local_blocks_) {
    // every rank owns a slice of the points; publish them all through a zero-padded sum
    DBuf<double> tmp((size_t)user_n_);
    tmp.zero(stream_);
    if (comm_->rank == 0) k_scatter_blocks<<<cdiv((int64_t)L_.n_cams * 9, 256), 256, 0, stream_>>>(L_.n_cams, 9, d_cam_off_.p, x_.p, tmp.p);
    if (L_.n_pts) k_scatter_blocks<<<cdiv((int64_t)L_.n_pts * 3, 256), 256, 0, stream_>>>(L_.n_pts, 3, d_pt_off_.p, x_.p + nc_, tmp.p);
    comm_allreduce_sum(comm_, tmp.p, (size_t)user_n_, stream_);
    // only the blocks of this problem are copied back (other entries of the user array are untouched)
    DBuf<long long> all_pt;   // all point offsets are not known locally: copy the touched ranges rank by rank
    // cameras
    k_gather_blocks<<<cdiv((int64_t)L_.n_cams * 9, 256), 256, 0, stream_>>>(L_.n_cams, 9, d_cam_off_.p, tmp.p, x_.p);
    k_scatter_blocks<<<cdiv((int64_t)L_.n_cams * 9, 256), 256, 0, stream_>>>(L_.n_cams, 9, d_cam_off_.p, x_.p, user_);
    // points: the global list of point offsets was kept on the host
    if (!all_pt_off_.empty()) {
      std::vector<long long> ap(all_pt_off_.begin(), all_pt_off_.end());
      all_pt.upload(ap, stream_);
      DBuf<double> buf(ap.size() * 3);
      k_gather_blocks<<<cdiv((int64_t)ap.size() * 3, 256), 256, 0, stream_>>>((int)ap.size(), 3, all_pt.p, tmp.p, buf.p);
      k_scatter_blocks<<<cdiv((int64_t)ap.size() * 3, 256), 256, 0, stream_>>>((int)ap.size(), 3, all_pt.p, buf.p, user_);
      SK_CUDA(cudaStreamSynchronize(stream_));
    }
    check_launch("store_state");
    SK_CUDA(cudaStreamSynchronize(stream_));
    return;
  }
  KScope k(prof_, SK_KF_LM, 2);
  k_scatter_blocks<<<cdiv((int64_t)L_.n_cams * 9, 256), 256, 0, stream_>>>(L_.n_cams, 9, d_cam_off_.p, x_.p, user_);
  if (L_.n_pts) k_scatter_blocks<<<cdiv((int64_t)L_.n_pts * 3, 256), 256, 0, stream_>>>(L_.n_pts, 3, d_pt_off_.p, x_.p + nc_, user_);
  check_launch("k_scatter_blocks");
}

void BaSolver::fill_summary(sk_solver_summary_data* d) {
  d->num_parameter_blocks = total_param_blocks_;
  d->num_parameters = total_params_;
  d->num_residual_blocks = total_obs_;
  d->num_residuals = 2 * total_obs_;
  // matvec launches that did work: launches issued after PCG termination inside a batch return at once and
  // would dilute the per-launch average the roofline figure is built from
  if (!explicit_schur_) d->kernel_launches[SK_KF_SCHUR_MATVEC] = n_real_matvecs_;
  if (fused_pcg_ && prof_.enabled) {
    // where the fused solves spent their time, on the device clock of their first CTA (pcg_fused.cu)
    unsigned long long ns[2] = {0, 0};
    SK_CUDA(cudaMemcpyAsync(ns, phase_ns_.p, sizeof(ns), cudaMemcpyDeviceToHost, stream_));
    SK_CUDA(cudaStreamSynchronize(stream_));
    d->kernel_ms[SK_KF_SCHUR_MATVEC] += 1e-6 * (double)ns[0];
    d->kernel_ms[SK_KF_PCG_VECTOR] += 1e-6 * (double)ns[1];
  }
}

// The iteration count of the LM iteration's linear solve has come back with the state block: the fused solve ran
// its + its / 10 products (and as many peer-window exchanges) without the host counting them.
void BaSolver::note_linear_iterations(int its) {
  if (!fused_in_flight_) return;
  fused_in_flight_ = false;
  const int64_t products = (int64_t)its + its / kResidualResetPeriod;
  n_real_matvecs_ += products;
  if (peer_.ok) peer_.seq += (unsigned long long)products;
}

ReduceJob BaSolver::cost_job() { return {tile_cost_.p, L_.n_tiles, SB_COST, 0}; }

void BaSolver::eval_jacobian(bool scale_valid, bool store, const int* guard) {
  {
    KScope k(prof_, SK_KF_EVALUATE_JACOBIAN, 3 + (L_.n_giant ? 1 : 0));
    launch_ba_evaluate(L_, x_.p, scale_valid ? scale_.p : nullptr, loss_, true, store, reinterpret_cast<double2*>(J2_.p),
                       reinterpret_cast<double2*>(r2_.p), grad_.p, cnorm2_.p, seg_a_.p, seg_b_.p, tile_cost_.p,
                       chunk_pt_.p, &st_.p->eval_failed, guard, stream_, functor_);
    launch_cam_reduce(L_, 9, seg_a_.p, grad_.p, guard, stream_);
    launch_cam_reduce(L_, 9, seg_b_.p, cnorm2_.p, guard, stream_);
  }
  if (comm_ && comm_->world > 1) {
    KScope k(prof_, SK_KF_COMM, 1);
    comm_group_start(comm_);
    comm_allreduce_sum(comm_, grad_.p, (size_t)nc_, stream_);
    comm_allreduce_sum(comm_, cnorm2_.p, (size_t)nc_, stream_);
    comm_group_end(comm_);
  }
}

void BaSolver::eval_cost(const double* xv, const int* guard) {
  KScope k(prof_, SK_KF_EVALUATE_COST);
  launch_ba_evaluate(L_, xv, nullptr, loss_, false, false, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, tile_cost_.p,
                     nullptr, &st_.p->eval_failed, guard, stream_, functor_);
}

// seg_a_ = segment partials of S_local * v, where v = `in` or (pcg_dir) the PCG direction z + beta p.
// Multi-GPU: the partials are reduced per camera and exchanged -- through the peer window (the consumer kernels gather it
// themselves: returns nullptr) or, as a fallback, into ybuf_ with an NCCL allreduce (returns ybuf_).  Single GPU: returns
// nullptr = "reduce the segment partials yourself".
const double* BaSolver::matvec(const double* in, bool pcg_dir, const int* guard) {
  {
    KScope k(prof_, SK_KF_SCHUR_MATVEC, 1 + (L_.n_giant ? 1 : 0));
    launch_ba_matvec(L_, reinterpret_cast<const double2*>(J2_.p), in, pcg_dir ? pz_.p : nullptr, pcg_dir ? pcg_.p : nullptr, einv_.p,
                     seg_a_.p, guard, stream_, have_tmapJ_ ? &tmapJ_ : nullptr);
  }
  if (comm_ && comm_->world > 1) {
    if (peer_.ok) {
      // the exchange is fused into the kernels on either side: this rank's reduced vector goes into its peer window and is
      // published; the consumer (k_pcg_reduce / k_pcg_resid2, same sequence number) adds all ranks' windows in rank order
      ++peer_.seq;
      KScope k(prof_, SK_KF_PCG_VECTOR);
      launch_cam_reduce9_warp(L_, seg_a_.p, nullptr, guard, stream_, &peer_.win, peer_.seq);
      return nullptr;
    }
    { KScope k(prof_, SK_KF_PCG_VECTOR); launch_cam_reduce9_warp(L_, seg_a_.p, ybuf_.p, guard, stream_); }
    KScope k(prof_, SK_KF_COMM);
    comm_allreduce_sum(comm_, ybuf_.p, (size_t)nc_, stream_);
    return ybuf_.p;
  }
  return nullptr;
}

ReduceJob BaSolver::linear_solve(const PcgDev** pcg_out) {
  const double2* J2 = reinterpret_cast<const double2*>(J2_.p);
  const double2* r2 = reinterpret_cast<const double2*>(r2_.p);
  int* lin_error = &st_.p->lin_error;
  {
    KScope k(prof_, SK_KF_SCHUR_SETUP, 3 + (L_.n_giant ? 1 : 0));
    const bool jacobi = !explicit_schur_ && opt_.preconditioner_type == SK_JACOBI;
    launch_ba_schur_setup(L_, J2, r2, D_.p, einv_.p, seg_b_.p, seg_M_.p, lin_error, jacobi, stream_);
    launch_cam_reduce(L_, 9, seg_b_.p, rhs_.p, nullptr, stream_);
    launch_cam_reduce(L_, 45, seg_M_.p, M45_.p, nullptr, stream_);
  }
  const bool multi = comm_ && comm_->world > 1;
  if (multi) {
    // the set-up failure flag (a point's E^T E not positive definite on SOME rank) travels with the right-hand side: every rank
    // must skip or run the PCG loop alike, or the per-iteration exchange loses its partners
    KScope k(prof_, SK_KF_COMM, 2);
    k_store_flag<<<1, 1, 0, stream_>>>(lin_error, rhs_.p + nc_);
    comm_group_start(comm_);
    comm_allreduce_sum(comm_, rhs_.p, (size_t)nc_ + 1, stream_);
    comm_allreduce_sum(comm_, M45_.p, (size_t)45 * L_.n_cams, stream_);
    comm_group_end(comm_);
  }
  if (explicit_schur_) {
    explicit_schur_solve();
    *pcg_out = nullptr;
  } else {
    const bool schur_jacobi = opt_.preconditioner_type == SK_SCHUR_JACOBI || opt_.preconditioner_type == SK_JACOBI;   // block preconditioner
    if (schur_jacobi) {
      KScope k(prof_, SK_KF_SCHUR_SETUP);
      launch_ba_precond_invert(L_, M45_.p, D_.p, Minv_.p, lin_error, stream_);
    }
    pcg_solve(schur_jacobi ? Minv_.p : nullptr, multi ? rhs_.p + nc_ : nullptr);
    *pcg_out = pcg_.p;
  }
  {
    KScope k(prof_, SK_KF_BACK_SUBSTITUTE, 2 + (L_.n_giant ? 1 : 0));
    launch_negate(nc_, px_.p, step_.p, stream_);
    launch_ba_back_substitute(L_, J2, r2, px_.p, einv_.p, step_.p, tile_mcc_.p, stream_);
  }
  return {tile_mcc_.p, L_.n_tiles, SB_MCC, 0};
}

// ConjugateGradientsSolver::Solve on the implicit Schur complement; all scalars stay on the device
// (pcg_kernels.cu).  The host enqueues iterations in batches and polls the state block between batches;
// kernels of iterations after termination return immediately.
void BaSolver::pcg_solve(const double* Minv, const double* global_lin_flag) {
  const int nb = pcg_blocks(L_.n_cams);
  PcgParams pp{opt_.min_linear_solver_iterations, opt_.max_linear_solver_iterations, opt_.eta};
  const int* active = &pcg_.p->active;
  double* part_bb = pcg_part_.p;
  double* part_rho = pcg_part_.p + kMaxPartials;
  double* part_pq = pcg_part_.p + 2 * kMaxPartials;
  double* part_Q = pcg_part_.p + 3 * kMaxPartials;
  {
    KScope k(prof_, SK_KF_PCG_VECTOR, 2);
    launch_pcg_begin(L_.n_cams, rhs_.p, Minv, px_.p, pr_.p, pz_.p, part_bb, part_rho, pcg_.p, &st_.p->lin_error, global_lin_flag,
                     fused_pcg_ ? grid_bar_.p : nullptr, stream_);
  }
  if (fused_pcg_) {
    // One launch for the whole solve; nothing is read back here -- the outcome travels with the LM iteration's state block.
    PcgSolveArgs a{};
    a.L = L_; a.J2 = reinterpret_cast<const double2*>(J2_.p); a.einv = einv_.p; a.seg_y = seg_a_.p;
    a.D = D_.p; a.Minv = Minv; a.b = rhs_.p; a.x = px_.p; a.p = pp_.p; a.r = pr_.p; a.z = pz_.p;
    a.part_pq = part_pq; a.part_Q = part_Q; a.part_rho = part_rho; a.st = pcg_.p; a.prm = pp; a.reset_period = kResidualResetPeriod;
    a.grid_bar = grid_bar_.p; a.phase_ns = prof_.enabled ? phase_ns_.p : nullptr;
    if (peer_.ok) { a.win = peer_.win; a.seq_base = peer_.seq; }
    KScope k(prof_, SK_KF_PCG_SOLVE);
    launch_pcg_solve(a, have_tmapJ_ ? &tmapJ_ : nullptr, stream_);
    fused_in_flight_ = true;
    return;
  }
  const int kBatch = 8, kResetPeriod = kResidualResetPeriod;
  int it = 0;
  bool done = false;
  while (!done && it < pp.max_iterations) {
    for (int b = 0; b < kBatch && it < pp.max_iterations; ++b) {
      ++it;
      // the head step of iterations 2.. runs at the tail of the previous iteration's update / resid2 kernel
      if (it == 1) { KScope k(prof_, SK_KF_PCG_VECTOR); launch_pcg_head(pcg_.p, part_rho, part_pq, part_Q, nb, pp, 0, stream_); }
      const double* y = matvec(pp_.p, true, active);
      const int recompute = (it % kResetPeriod == 0) ? 1 : 0;
      {
        KScope k(prof_, SK_KF_PCG_VECTOR, 2);
        launch_pcg_reduce(L_, seg_a_.p, y, D_.p, pz_.p, pp_.p, part_pq, pcg_.p, stream_, peer_.ok ? &peer_.win : nullptr, peer_.seq);
        launch_pcg_update(L_.n_cams, Minv, rhs_.p, px_.p, pp_.p, pr_.p, pz_.p, part_pq, recompute, part_Q, part_rho, pcg_.p, pp, stream_);
      }
      if (recompute) {
        const double* yx = matvec(px_.p, false, active);
        KScope k(prof_, SK_KF_PCG_VECTOR);
        launch_pcg_resid2(L_, seg_a_.p, yx, D_.p, Minv, rhs_.p, px_.p, pr_.p, pz_.p, part_Q, part_rho, pcg_.p, part_pq, pp, stream_,
                          peer_.ok ? &peer_.win : nullptr, peer_.seq);
      }
    }
    SK_CUDA(cudaMemcpyAsync(pcg_h_.p, pcg_.p, sizeof(PcgDev), cudaMemcpyDeviceToHost, stream_));
    SK_CUDA(cudaStreamSynchronize(stream_));
    prof_.collect();
    done = pcg_h_.p->active == 0;
  }
  // A peer-window time-out (a rank stopped, or the ranks lost lockstep) is not thrown here: the rank that saw it has marked its
  // solve LIN_FATAL on the device, the flag reaches every rank with the next scalar allreduce (SB_FLAG_LIN) and all of them
  // terminate with FAILURE together -- an exception on one rank would leave the others waiting in that allreduce.
  const int its = pcg_h_.p->iter;
  n_real_matvecs_ += (its + its / kResetPeriod) * (L_.n_giant ? 2 : 1);
}

double BaSolver::time_linear_operator(int reps) {
  SK_REQUIRE(!explicit_schur_, SK_ERR_UNSUPPORTED, "the implicit Schur product exists for ITERATIVE_SCHUR solvers only");
  SK_REQUIRE(n_jac_evals_ > 0, SK_ERR_INVALID_ARGUMENT, "sk_solver_time_schur_product: run sk_solver_minimize first (no linearisation yet)");
  // development: SKERES_TIME_MODE selects what surrounds the product (why is it slower inside the PCG loop than back to back?)
  //   0 back to back | 1 a one-CTA kernel between products | 2 the second-level camera reduction between products
  //   3 the camera reduction alone | 4 back to back with the PCG-style input (z + beta p) | 5 back to back, events around every launch
  const int mode = [] { const char* e = getenv("SKERES_TIME_MODE"); return e ? atoi(e) : 0; }();
  cudaEvent_t a, b;
  SK_CUDA(cudaEventCreate(&a)); SK_CUDA(cudaEventCreate(&b));
  launch_fill(nc_, 1.0, pp_.p, stream_);
  launch_fill(nc_, 0.5, pz_.p, stream_);
  const double2* J2 = reinterpret_cast<const double2*>(J2_.p);
  const CUtensorMap* tm = have_tmapJ_ ? &tmapJ_ : nullptr;
  auto product = [&]() {
    if (mode == 4) launch_ba_matvec(L_, J2, pp_.p, pz_.p, pcg_.p, einv_.p, seg_a_.p, nullptr, stream_, tm);
    else if (mode != 3) launch_ba_matvec(L_, J2, pp_.p, nullptr, nullptr, einv_.p, seg_a_.p, nullptr, stream_, tm);
    if (mode == 1) launch_fill(1, 0.0, ybuf_.p, stream_);
    if (mode == 2 || mode == 3) launch_cam_reduce(L_, 9, seg_a_.p, ybuf_.p, nullptr, stream_);
  };
  for (int i = 0; i < 3; ++i) product();
  double total_ms = 0.0;
  if (mode == 5) {
    std::vector<cudaEvent_t> ev((size_t)2 * reps);
    for (auto& e : ev) SK_CUDA(cudaEventCreate(&e));
    for (int i = 0; i < reps; ++i) { SK_CUDA(cudaEventRecord(ev[2 * i], stream_)); product(); SK_CUDA(cudaEventRecord(ev[2 * i + 1], stream_)); }
    SK_CUDA(cudaStreamSynchronize(stream_));
    for (int i = 0; i < reps; ++i) { float ms = 0.f; cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]); total_ms += ms; }
    for (auto& e : ev) cudaEventDestroy(e);
  } else {
    SK_CUDA(cudaEventRecord(a, stream_));
    for (int i = 0; i < reps; ++i) product();
    SK_CUDA(cudaEventRecord(b, stream_));
    SK_CUDA(cudaStreamSynchronize(stream_));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    total_ms = ms;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  return total_ms / reps;
}

// The host copy of the layout is hundreds of MB at Venice scale and nothing reads its per-observation arrays once the device
// copies exist.  Returning it to the OS takes up to 80 ms, so it is handed to a detached thread (plain host memory, no CUDA
// calls) -- at the END OF CONSTRUCTION, where it overlaps the first LM iterations, not at destruction, where it delayed the
// caller and (through the process's mmap lock) whatever else the tear-down unmaps.  Counts and the camera table stay.
void BaSolver::release_host_layout() {
  if ((size_t)H_.n_obs < (size_t)1 << 20) return;            // small: freed in place with the other members
  try {
    auto* drop = new BaLayoutHost(std::move(H_));
    H_ = BaLayoutHost{};
    H_.n_obs = drop->n_obs; H_.n_pts = drop->n_pts; H_.n_cams = drop->n_cams; H_.n_tiles = drop->n_tiles; H_.n_segs = drop->n_segs;
    H_.max_seg_tile = drop->max_seg_tile; H_.max_pt_tile = drop->max_pt_tile; H_.n_giant = drop->n_giant; H_.n_chunks = drop->n_chunks;
    H_.input_was_sorted = drop->input_was_sorted;
    H_.cam_offset = drop->cam_offset;                         // exchange_local_totals
    std::thread([drop] { delete drop; }).detach();
  } catch (...) {}                                            // no thread: *drop is leaked at worst, H_ keeps what it has
}

BaSolver::~BaSolver() { peer_allreduce_destroy(&peer_); }

void BaSolver::exchange_local_totals() {
  local_blocks_ = true;
  const int world = comm_ ? comm_->world : 1;
  if (world == 1) return;
  double check = 0.0;                                        // exact in a double: < 2^24 cameras x offsets mod 2^20
  for (int64_t o : H_.cam_offset) check += (double)(o & 0xfffff);
  std::vector<double> h = {(double)L_.n_obs, (double)L_.n_pts, (double)L_.n_cams, check};
  DBuf<double> d(h.size());
  d.upload(h, stream_);
  comm_allreduce_sum(comm_, d.p, h.size(), stream_);
  std::vector<double> g(h.size());
  d.download(g.data(), g.size(), stream_);
  SK_CUDA(cudaStreamSynchronize(stream_));
  SK_REQUIRE(g[2] == h[2] * world && g[3] == h[3] * world, SK_ERR_INVALID_ARGUMENT,
             "rank-local residual blocks: the ranks do not agree on the camera blocks (this rank has %d; declare every camera on "
             "every rank with sk_problem_add_parameter_blocks)", L_.n_cams);
  total_obs_ = (int64_t)g[0];
  total_param_blocks_ = (int64_t)L_.n_cams + (int64_t)g[1];
  total_params_ = 9 * (int64_t)L_.n_cams + 3 * (int64_t)g[1];
}

void BaSolver::fill_totals(int64_t total_obs, int64_t total_blocks, int64_t total_params, std::vector<int64_t>&& all_pt_off) {
  total_obs_ = total_obs; total_param_blocks_ = total_blocks; total_params_ = total_params; all_pt_off_ = std::move(all_pt_off);
}

}  // namespace sk
