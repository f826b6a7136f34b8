// comm.cuh — NCCL communicator, loaded lazily with dlopen so that single-GPU use never touches
// libnccl and the library has no link-time dependency on it (inside a torch process the already
// loaded libnccl.so.2 is reused).
#pragma once
#include <vector>

#include "common.cuh"

struct sk_comm {
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
  std::vector<void*> peer_cache;   // sk::PeerAllreduce* of destroyed solvers, kept mapped for the next solver of the same size
};

namespace sk {

void comm_get_unique_id(char* id128);
sk_comm* comm_create(const char* id128, int rank, int world);
void comm_destroy(sk_comm* c);
// In-place sum / max allreduce of `count` doubles on `stream`; no-op when c == nullptr or world == 1.
void comm_allreduce_sum(sk_comm* c, double* buf, size_t count, cudaStream_t stream);
void comm_allreduce_max(sk_comm* c, double* buf, size_t count, cudaStream_t stream);
extern double g_comm_host_seconds;   // development trace (SKERES_TRACE_HOST)
extern long g_comm_calls;
void comm_group_start(sk_comm* c);

// ---- peer window: allreduce of a camera-sized vector through NVLink peer memory, fused into the kernels on either side ----
// Every rank owns one device allocation, mapped into every other rank with CUDA IPC:
//   double data[2][stride]            the rank's contribution, double-buffered by the parity of the sequence number
//   u64    flags[kMaxPeers][2]        flags[r][parity] = sequence number of the last contribution rank r has published
//   int    error                      set when a wait timed out (a rank died or the ranks lost lockstep)
// Producer kernel (k_cam_reduce9_warp): writes its data, then its last CTA stores the sequence number into flags[me][parity] of
// EVERY rank (release, system scope).  Consumer kernels (k_pcg_reduce, k_pcg_resid2): wait until their own flags[r][parity]
// reach the sequence number for every r, then add the ranks' contributions in rank order -- the same bits on every rank.
constexpr int kMaxPeers = 8;
struct PeerWindow {
  double* data[kMaxPeers];                  // base of rank r's window (own: the local pointer)
  unsigned long long* flags[kMaxPeers];     // flags array inside rank r's window
  int* error;                               // own error word
  unsigned int* done_count;                 // own: CTAs of the running producer kernel that have written their part
  long long stride;                         // doubles per parity slot
  unsigned long long timeout_ns;            // bound of one wait for the peers (SKERES_PEER_TIMEOUT_S, default 60 s)
  const unsigned char* cam_mask;            // [n_cams] bit r: rank r holds observations of the camera, i.e. contributes to its sums;
                                            // the gather reads a camera only from those ranks (nullptr: from all)
  const unsigned char* vb_own;              // [ceil(n_cams / 8)] != 0: this rank holds observations of a camera of that virtual block of 8
                                            // cameras (pcg_device.cuh), i.e. has something to contribute for it (nullptr: assume all)
  int rank, world;                          // world == 0: no peer window (single GPU or NCCL path)
};
struct PeerAllreduce {
  PeerWindow win{};
  void* opened[kMaxPeers] = {};
  DBuf<double> mem;
  DBuf<unsigned char> cam_mask;             // see PeerWindow::cam_mask
  DBuf<unsigned char> vb_own;               // see PeerWindow::vb_own
  unsigned long long seq = 0;               // sequence number of the last exchange
  bool ok = false;
  sk_comm* owner = nullptr;                 // the communicator whose cache takes the window back
  size_t count = 0;                         // doubles per contribution it was created for
};
// Collective over `c`: allocates the window for `count` doubles and exchanges the IPC handles (through one NCCL allreduce).
// Leaves pa->ok == false (NCCL allreduce stays in use) when peer mapping is unavailable or SKERES_PEER_ALLREDUCE=0.
// Windows are recycled: peer_allreduce_destroy hands a mapped window back to its communicator and the next create of the same
// size takes it without any collective -- closing an IPC mapping synchronises the ranks and took 0.2-0.5 s of every
// sk_solver_destroy (profiles/r02_v9_*), more than the whole set-up of a solver.  Every rank must create and destroy solvers in
// the same order (they do: a solve is collective), so that all of them hit or miss the cache alike.  comm_destroy closes them.
void peer_allreduce_create(sk_comm* c, size_t count, cudaStream_t stream, PeerAllreduce* pa);
void peer_allreduce_destroy(PeerAllreduce* pa);
void comm_group_end(sk_comm* c);

}  // namespace sk
