// comm.cuh — NCCL communicator, loaded lazily with dlopen so that single-GPU use never touches
// libnccl and the library has no link-time dependency on it (inside a torch process the already
// loaded libnccl.so.2 is reused).
#pragma once
#include "common.cuh"

struct sk_comm {
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
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
void comm_group_end(sk_comm* c);

}  // namespace sk
