// comm.cu — lazy NCCL binding (see comm.cuh).  Only the handful of entry points the point-partitioned
// solver needs: unique id, communicator, sum/max allreduce of doubles, grouping.
#include "comm.cuh"

#include <dlfcn.h>

#include <chrono>
#include <mutex>

namespace sk {

namespace {

// Minimal declarations matching nccl.h (2.x ABI); the library is resolved at run time.
constexpr int kNcclUniqueIdBytes = 128;
struct NcclUniqueId { char internal[kNcclUniqueIdBytes]; };
typedef int ncclResult;
constexpr int kNcclSum = 0, kNcclMax = 2, kNcclFloat64 = 8;

struct Api {
  void* handle = nullptr;
  ncclResult (*GetUniqueId)(NcclUniqueId*) = nullptr;
  ncclResult (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
  ncclResult (*CommDestroy)(void*) = nullptr;
  ncclResult (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  ncclResult (*GroupStart)() = nullptr;
  ncclResult (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult) = nullptr;
};

Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.handle) break; }
    if (!a.handle) return;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.handle, "ncclAllReduce");
    a.GroupStart = (decltype(a.GroupStart))dlsym(a.handle, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.handle, "ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
  });
  if (!a.handle || !a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy || !a.GroupStart || !a.GroupEnd)
    throw Error(SK_ERR_NCCL, "libnccl.so.2 could not be loaded (multi-GPU solves need NCCL)");
  return a;
}

void check(ncclResult r, const char* what) {
  if (r != 0) {
    Api& a = api();
    throw Error(SK_ERR_NCCL, fmt("%s failed: %s", what, a.GetErrorString ? a.GetErrorString(r) : "?"));
  }
}

}  // namespace

void comm_get_unique_id(char* id128) {
  static_assert(SK_COMM_UNIQUE_ID_BYTES == kNcclUniqueIdBytes, "unique id size");
  NcclUniqueId id;
  check(api().GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(id128, id.internal, kNcclUniqueIdBytes);
}

sk_comm* comm_create(const char* id128, int rank, int world) {
  SK_REQUIRE(world >= 1 && rank >= 0 && rank < world, SK_ERR_INVALID_ARGUMENT, "bad rank %d / world size %d", rank, world);
  NcclUniqueId id;
  memcpy(id.internal, id128, kNcclUniqueIdBytes);
  sk_comm* c = new sk_comm;
  c->rank = rank; c->world = world;
  try {
    check(api().CommInitRank(&c->nccl_comm, world, id, rank), "ncclCommInitRank");
  } catch (...) { delete c; throw; }
  return c;
}

void comm_destroy(sk_comm* c) {
  if (!c) return;
  if (c->nccl_comm) api().CommDestroy(c->nccl_comm);
  delete c;
}

double g_comm_host_seconds = 0.0;   // development trace: host wall time spent enqueuing collectives
long g_comm_calls = 0;
static double host_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void comm_allreduce_sum(sk_comm* c, double* buf, size_t count, cudaStream_t stream) {
  if (!c || c->world == 1 || count == 0) return;
  const double t0 = host_now();
  check(api().AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, c->nccl_comm, stream), "ncclAllReduce(sum)");
  g_comm_host_seconds += host_now() - t0; ++g_comm_calls;
}
void comm_allreduce_max(sk_comm* c, double* buf, size_t count, cudaStream_t stream) {
  if (!c || c->world == 1 || count == 0) return;
  check(api().AllReduce(buf, buf, count, kNcclFloat64, kNcclMax, c->nccl_comm, stream), "ncclAllReduce(max)");
}
void comm_group_start(sk_comm* c) { if (c && c->world > 1) check(api().GroupStart(), "ncclGroupStart"); }
void comm_group_end(sk_comm* c) { if (c && c->world > 1) check(api().GroupEnd(), "ncclGroupEnd"); }

}  // namespace sk
