// common.cuh — error plumbing, device buffers, launch accounting shared by all translation units.
#pragma once
#include <cuda_runtime.h>

#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/skeres.h"

namespace sk {

// Thrown inside the library, converted to an sk_status at the C boundary (never crosses it).
struct Error : std::runtime_error {
  int status;
  Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};

inline std::string fmt(const char* f, ...) {
  char buf[1024];
  va_list ap; va_start(ap, f); vsnprintf(buf, sizeof buf, f, ap); va_end(ap);
  return buf;
}

#define SK_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      throw ::sk::Error(SK_ERR_CUDA, ::sk::fmt("%s failed: %s (%s:%d)", #expr,                \
                                               cudaGetErrorString(e_), __FILE__, __LINE__));  \
  } while (0)

#define SK_REQUIRE(cond, status, ...)                                  \
  do { if (!(cond)) throw ::sk::Error((status), ::sk::fmt(__VA_ARGS__)); } while (0)

// Caching device allocator behind DBuf: freed blocks are kept per (device, rounded size) and handed to the next request of
// that size, so that building a solver after another one was destroyed does not pay cudaMalloc / cudaFree of GB-sized
// buffers again (measured 0.08-0.40 s per solver on a B200).  sk_release_cached_memory() returns everything to the driver;
// SKERES_POOL_MAX_GB (default 16) bounds what is kept; SKERES_POOL_MAX_GB=0 disables caching.
// A block is filed under the device it was ALLOCATED on (recorded by DBuf), whatever device is current when it is released --
// a handle destroyed from another thread, or after sk_set_device to another GPU, must not hand device-A memory to a request
// on device B.  A released block may still be read or written by kernels in flight (exception paths, handles destroyed right
// after an asynchronous call), and the pool has no stream to order the next owner behind them: the owning device is
// synchronised before the block is cached or freed.  Releases happen on tear-down paths only, never inside a solve.
class DevicePool {
 public:
  static DevicePool& get() { static DevicePool* p = new DevicePool; return *p; }   // never destroyed: outlives the CUDA context teardown
  static size_t rounded(size_t bytes) { const size_t g = bytes >= (1u << 20) ? (size_t)(2u << 20) : (size_t)512; return (bytes + g - 1) / g * g; }
  void* take(size_t bytes, int* dev_out) {
    const size_t r = rounded(bytes);
    int dev = 0; cudaGetDevice(&dev);
    *dev_out = dev;
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = free_.find({dev, r});
      if (it != free_.end()) { void* p = it->second; free_.erase(it); cached_ -= r; return p; }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, r);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); release_all(); e = cudaMalloc(&p, r); }   // give cached blocks back and retry
    if (e != cudaSuccess) throw Error(SK_ERR_CUDA, fmt("cudaMalloc of %zu bytes failed: %s", r, cudaGetErrorString(e)));
    return p;
  }
  void give(void* p, size_t bytes, int dev) {
    const size_t r = rounded(bytes);
    int cur = 0; cudaGetDevice(&cur);
    if (cur != dev) cudaSetDevice(dev);
    const auto t0 = std::chrono::steady_clock::now();
    cudaDeviceSynchronize();                           // nothing in flight may still touch the block (see above)
    const auto t1 = std::chrono::steady_clock::now();
    bool keep = false;
    {
      std::lock_guard<std::mutex> g(mu_);
      if (cached_ + r <= max_cached_) { free_.insert({{dev, r}, p}); cached_ += r; keep = true; }
    }
    if (!keep) cudaFree(p);
    t_sync += std::chrono::duration<double>(t1 - t0).count();
    t_free += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    ++n_give;
    if (cur != dev) cudaSetDevice(cur);
  }
  void release_all() {
    std::multimap<std::pair<int, size_t>, void*> drop;
    { std::lock_guard<std::mutex> g(mu_); drop.swap(free_); cached_ = 0; }
    int cur = 0; cudaGetDevice(&cur);
    for (auto& kv : drop) { cudaSetDevice(kv.first.first); cudaFree(kv.second); }
    cudaSetDevice(cur);
  }
  size_t cached_bytes() { std::lock_guard<std::mutex> g(mu_); return cached_; }
  double t_sync = 0.0, t_free = 0.0; long n_give = 0;   // development trace (SKERES_TRACE_HOST): where tear-down time goes

 private:
  DevicePool() { const char* e = std::getenv("SKERES_POOL_MAX_GB"); max_cached_ = (size_t)((e ? atof(e) : 16.0) * (double)(1ull << 30)); }
  std::mutex mu_;
  std::multimap<std::pair<int, size_t>, void*> free_;
  size_t cached_ = 0, max_cached_ = 0;
};

// Device buffer with RAII; all device memory of the library is owned through these.
template <class T>
struct DBuf {
  T* p = nullptr; size_t n = 0; int dev = 0;            // dev: the device the block lives on
  DBuf() = default;
  explicit DBuf(size_t count) { alloc(count); }
  DBuf(const DBuf&) = delete; DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), n(o.n), dev(o.dev) { o.p = nullptr; o.n = 0; }
  DBuf& operator=(DBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; dev = o.dev; o.p = nullptr; o.n = 0; } return *this; }
  ~DBuf() { release(); }
  void release() { if (p) DevicePool::get().give(p, n * sizeof(T), dev); p = nullptr; n = 0; }
  void alloc(size_t count) {
    release(); n = count;
    if (count) p = static_cast<T*>(DevicePool::get().take(count * sizeof(T), &dev));
  }
  void zero(cudaStream_t s) { if (n) SK_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void upload(const T* h, size_t count, cudaStream_t s) { if (count) SK_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s)); }
  void upload(const std::vector<T>& h, cudaStream_t s) { if (n < h.size()) alloc(h.size()); upload(h.data(), h.size(), s); }
  void download(T* h, size_t count, cudaStream_t s) const { if (count) SK_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s)); }
};

// Pinned host buffer (readback targets).  Freed blocks are kept per size and reused: cudaFreeHost unmaps pinned pages and was
// seen to stall a solver's tear-down for up to 0.45 s when another thread was returning large host vectors to the OS at the
// same time (both need the process's mmap lock); the buffers are a few hundred bytes each.
class PinnedPool {
 public:
  static PinnedPool& get() { static PinnedPool* p = new PinnedPool; return *p; }
  void* take(size_t bytes) {
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = free_.find(bytes);
      if (it != free_.end()) { void* p = it->second; free_.erase(it); return p; }
    }
    void* p = nullptr;
    SK_CUDA(cudaMallocHost(&p, bytes));
    return p;
  }
  void give(void* p, size_t bytes) {
    std::lock_guard<std::mutex> g(mu_);
    if (free_.size() < 256) { free_.insert({bytes, p}); return; }
    cudaFreeHost(p);
  }
 private:
  std::mutex mu_;
  std::multimap<size_t, void*> free_;
};
template <class T>
struct HBuf {
  T* p = nullptr; size_t n = 0;
  HBuf() = default;
  explicit HBuf(size_t count) { alloc(count); }
  HBuf(const HBuf&) = delete; HBuf& operator=(const HBuf&) = delete;
  ~HBuf() { release(); }
  void release() { if (p) PinnedPool::get().give(p, n * sizeof(T)); p = nullptr; n = 0; }
  void alloc(size_t count) { release(); n = count; if (count) p = static_cast<T*>(PinnedPool::get().take(count * sizeof(T))); }
};

// Per-solve launch accounting + optional CUDA-event timing per kernel family.
struct Profiler {
  bool enabled = false;
  unsigned family_mask = ~0u;   // which kernel families get CUDA events when enabled (profile_kernels = 2: matvec only)
  cudaStream_t stream = nullptr;
  int64_t launches[SK_KF_COUNT] = {0};
  double ms[SK_KF_COUNT] = {0};
  struct Pending { int family; cudaEvent_t a, b; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; SK_CUDA(cudaEventCreate(&e)); return e;
  }
  // Called after the stream has been synchronised.
  void collect() {
    for (auto& p : pending) {
      float t = 0; cudaEventElapsedTime(&t, p.a, p.b); ms[p.family] += t;
      pool.push_back(p.a); pool.push_back(p.b);
    }
    pending.clear();
  }
  ~Profiler() { for (auto& p : pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); } for (auto e : pool) cudaEventDestroy(e); }
};

// Scope covering `count` kernel launches of one family.
struct KScope {
  Profiler& pr; int family; cudaEvent_t a = nullptr;
  KScope(Profiler& p, int fam, int count = 1) : pr(p), family(fam) {
    pr.launches[fam] += count;
    if (pr.enabled && ((pr.family_mask >> fam) & 1u)) { a = pr.get(); cudaEventRecord(a, pr.stream); }
  }
  ~KScope() {
    if (a != nullptr) { cudaEvent_t b = pr.get(); cudaEventRecord(b, pr.stream); pr.pending.push_back({family, a, b}); }
  }
};

inline void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw Error(SK_ERR_CUDA, fmt("kernel launch %s failed: %s", what, cudaGetErrorString(e)));
}

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace sk
