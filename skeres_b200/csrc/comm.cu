// comm.cu — lazy NCCL binding (see comm.cuh).  Only the handful of entry points the point-partitioned
// solver needs: unique id, communicator, sum/max allreduce of doubles, grouping.
#include "comm.cuh"

#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

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

static void peer_allreduce_close(PeerAllreduce* pa);
void comm_destroy(sk_comm* c) {
  if (!c) return;
  for (void* p : c->peer_cache) { PeerAllreduce* pa = static_cast<PeerAllreduce*>(p); peer_allreduce_close(pa); delete pa; }
  c->peer_cache.clear();
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

// ---- peer window ---------------------------------------------------------------------------------------------------------
static void peer_allreduce_close(PeerAllreduce* pa) {
  for (int r = 0; r < kMaxPeers; ++r)
    if (pa->opened[r]) { cudaIpcCloseMemHandle(pa->opened[r]); pa->opened[r] = nullptr; }
  pa->ok = false; pa->win = PeerWindow{};
}

void peer_allreduce_create(sk_comm* c, size_t count, cudaStream_t stream, PeerAllreduce* pa) {
  pa->ok = false; pa->win = PeerWindow{}; pa->seq = 0; pa->owner = nullptr; pa->count = count;
  if (!c || c->world <= 1 || c->world > kMaxPeers) return;
  { const char* e = getenv("SKERES_PEER_ALLREDUCE"); if (e && e[0] == '0') return; }
  for (size_t i = 0; i < c->peer_cache.size(); ++i) {       // a window of this size left by an earlier solver: no collective needed
    PeerAllreduce* cached = static_cast<PeerAllreduce*>(c->peer_cache[i]);
    if (cached->count != count) continue;
    c->peer_cache.erase(c->peer_cache.begin() + (long)i);
    *pa = std::move(*cached);
    delete cached;
    pa->win.cam_mask = nullptr; pa->win.vb_own = nullptr;
    return;
  }
  const int world = c->world, rank = c->rank;
  const size_t stride = (count + 15) & ~(size_t)15;
  const size_t flag_doubles = (size_t)kMaxPeers * 2, tail_doubles = 2;            // flags, then error + done_count
  pa->mem.alloc(2 * stride + flag_doubles + tail_doubles);
  pa->mem.zero(stream);                                                            // flags = 0 before anybody can publish
  // exchange the IPC handles: rank r's 64 bytes travel as 64 doubles (one byte each: exact under a sum of zeros) plus a
  // "can export" marker, in one NCCL allreduce -- which also orders every rank's zeroing before the first exchange
  cudaIpcMemHandle_t mine;
  const bool exported = cudaIpcGetMemHandle(&mine, pa->mem.p) == cudaSuccess;
  if (!exported) cudaGetLastError();
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  const size_t per = 65;
  std::vector<double> h((size_t)world * per, 0.0);
  for (size_t b = 0; b < 64; ++b) h[rank * per + b] = (double)reinterpret_cast<const unsigned char*>(&mine)[b];
  h[rank * per + 64] = exported ? 1.0 : 0.0;
  DBuf<double> d(h.size());
  d.upload(h, stream);
  comm_allreduce_sum(c, d.p, h.size(), stream);
  d.download(h.data(), h.size(), stream);
  SK_CUDA(cudaStreamSynchronize(stream));
  bool all = true;
  for (int r = 0; r < world; ++r) all = all && h[r * per + 64] == 1.0;
  // every rank takes the same decision from the same data; a mapping failure below is reported to the others as well
  double mapped = all ? 1.0 : 0.0;
  if (all) {
    for (int r = 0; r < world && mapped == 1.0; ++r) {
      if (r == rank) { pa->opened[r] = nullptr; pa->win.data[r] = pa->mem.p; continue; }
      cudaIpcMemHandle_t hr;
      for (size_t b = 0; b < 64; ++b) reinterpret_cast<unsigned char*>(&hr)[b] = (unsigned char)h[r * per + b];
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mapped = 0.0; break; }
      pa->opened[r] = ptr; pa->win.data[r] = static_cast<double*>(ptr);
    }
  }
  std::vector<double> m = {mapped};
  d.upload(m.data(), 1, stream);
  comm_allreduce_sum(c, d.p, 1, stream);
  d.download(m.data(), 1, stream);
  SK_CUDA(cudaStreamSynchronize(stream));
  if (m[0] != (double)world) { peer_allreduce_close(pa); return; }
  for (int r = 0; r < world; ++r) pa->win.flags[r] = reinterpret_cast<unsigned long long*>(pa->win.data[r] + 2 * stride);
  pa->win.error = reinterpret_cast<int*>(pa->mem.p + 2 * stride + flag_doubles);
  pa->win.done_count = reinterpret_cast<unsigned int*>(pa->win.error + 1);
  pa->win.stride = (long long)stride; pa->win.rank = rank; pa->win.world = world;
  { const char* e = getenv("SKERES_PEER_TIMEOUT_S"); const double sec = e ? atof(e) : 60.0;
    pa->win.timeout_ns = (unsigned long long)((sec > 0.0 ? sec : 60.0) * 1e9); }
  pa->ok = true; pa->owner = c;
  if (getenv("SKERES_TRACE_HOST")) fprintf(stderr, "[skeres] rank %d: peer window of %zu doubles mapped on %d ranks\n", rank, count, world);
}

void peer_allreduce_destroy(PeerAllreduce* pa) {
  if (pa->ok && pa->owner != nullptr && pa->owner->peer_cache.size() < 4) {     // keep it mapped for the next solver of this size
    sk_comm* c = pa->owner;
    c->peer_cache.push_back(new PeerAllreduce(std::move(*pa)));
    pa->ok = false; pa->win = PeerWindow{};
    for (int r = 0; r < kMaxPeers; ++r) pa->opened[r] = nullptr;
    return;
  }
  peer_allreduce_close(pa);
}

}  // namespace sk
