// ba_solver.cuh — bundle-adjustment back end of the LM driver (see ba_solver.cu).
#pragma once
#include "ba_kernels.cuh"
#include "ba_layout_device.cuh"
#include "lm_solver.cuh"

namespace sk {

class BaSolver : public LmSolver {
 public:
  // user_params: device pointer of the user's parameter DoubleArray (all blocks live in it).
  // dev != nullptr && dev->valid: the layout arrays are already on the device (ba_layout_device.cu) and are adopted; `layout`
  // then carries counts, per-tile maxima and the camera table only.
  // functor_id: the cost functor of every residual block -- the built-in SnavelyReprojectionError or a functor of the same
  // shape (2; 9, 3; 2 constants) registered from source, which runs in the same tile kernel (user_functor.cuh).
  BaSolver(const sk_solver_options& opt, cudaStream_t stream, BaLayoutHost&& layout, double* user_params,
           int64_t user_n, LossSpec loss, BaLayoutDevice* dev = nullptr, int functor_id = SK_FUNCTOR_SNAVELY_REPROJECTION_ERROR);
  ~BaSolver() override;
  void fill_totals(int64_t total_obs, int64_t total_blocks, int64_t total_params, std::vector<int64_t>&& all_pt_off);
  // Rank-local ingestion (sk_solver_options.residual_blocks_are_local): sums the per-rank observation / point counts for
  // the summary and checks that every rank built the same camera table.  Collective: every rank must call it.
  void exchange_local_totals();
  double time_linear_operator(int reps) override;

  // ---- test / debug access (sk_debug_* entry points) ---------------------------------------------
  const BaDev& layout() const { return L_; }

 protected:
  void eval_jacobian(bool scale_valid, bool store, const int* guard) override;
  void eval_cost(const double* xv, const int* guard) override;
  ReduceJob cost_job() override;
  ReduceJob linear_solve(const PcgDev** pcg_out) override;
  void load_state() override;
  void store_state() override;
  void fill_summary(sk_solver_summary_data* d) override;
  void note_linear_iterations(int iterations) override;

 private:
  const double* matvec(const double* in, bool pcg_dir, const int* guard);
  void pcg_solve(const double* Minv, const double* global_lin_flag);
  void build_pair_lists();
  void release_host_layout();
  void build_tile_records();
  void explicit_schur_solve();

  BaLayoutHost H_;
  BaDev L_{};
  double* user_; int64_t user_n_;
  LossSpec loss_;
  int functor_;                        // cost functor of the residual blocks (tile evaluation kernel)
  bool explicit_schur_ = false;
  PeerAllreduce peer_;                 // multi-GPU: NVLink peer window for the per-PCG-iteration exchange (comm.cuh)
  bool local_blocks_ = false;          // this rank was given only its own residual blocks (no publication of foreign points)
  int64_t total_obs_ = 0, total_param_blocks_ = 0, total_params_ = 0;
  std::vector<int64_t> all_pt_off_;
  DBuf<int> d_tile_obs_, d_tile_pt_, d_tile_seg_, d_pt_ptr_, d_seg_ptr_, d_seg_cam_, d_cam_seg_ptr_, d_cam_seg_, d_seg_pos_;
  DBuf<int> d_tile_np_, d_gp_begin_, d_gp_count_, d_gp_point_;   // long tracks (ba_layout.h)
  CUtensorMap tmapJ_{}; bool have_tmapJ_ = false;                // 2-D TMA view of J2_ (prefetching matvec)
  DBuf<unsigned char> d_tile_rec_;                               // per-tile metadata records (prefetching matvec)
  DBuf<double> chunk_pt_;                                        // [n_chunks][6] point partials of chunk tiles
  DBuf<unsigned short> d_obs_slot_, d_obs_ptl_, d_seg_perm_;
  DBuf<double> d_obs_;
  DBuf<long long> d_cam_off_, d_pt_off_;
  DBuf<double> J2_, r2_, einv_, seg_a_, seg_b_, seg_M_, M45_, Minv_, tile_cost_, tile_mcc_;
  DBuf<double> rhs_, px_, pr_, pp_, pz_, ybuf_, pcg_part_, S_;
  DBuf<PcgDev> pcg_;
  HBuf<PcgDev> pcg_h_;
  // fused PCG solve (pcg_fused.cu): the default for ITERATIVE_SCHUR; SKERES_PCG=sequence keeps the kernel sequence
  bool fused_pcg_ = false, fused_in_flight_ = false;
  DBuf<unsigned int> grid_bar_;
  DBuf<unsigned long long> phase_ns_;
  // explicit Schur: for every camera pair (c1 < c2) sharing points, the list of observation pairs
  DBuf<int> pair_ptr_, pair_c1_, pair_c2_, pair_o1_, pair_o2_, pair_pt_;
  int n_pair_groups_ = 0;
 public:
  int64_t n_real_matvecs_ = 0;   // matvec launches that did work (not skipped by the PCG guard)
};

}  // namespace sk
