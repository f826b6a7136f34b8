// common.cuh — error plumbing, device buffers, launch accounting shared by all translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
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

// Device buffer with RAII; all device memory of the library is owned through these.
template <class T>
struct DBuf {
  T* p = nullptr; size_t n = 0;
  DBuf() = default;
  explicit DBuf(size_t count) { alloc(count); }
  DBuf(const DBuf&) = delete; DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DBuf& operator=(DBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
  ~DBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  void alloc(size_t count) {
    release(); n = count;
    if (count) SK_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
  }
  void zero(cudaStream_t s) { if (n) SK_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void upload(const T* h, size_t count, cudaStream_t s) { if (count) SK_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s)); }
  void upload(const std::vector<T>& h, cudaStream_t s) { if (n < h.size()) alloc(h.size()); upload(h.data(), h.size(), s); }
  void download(T* h, size_t count, cudaStream_t s) const { if (count) SK_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s)); }
};

// Pinned host buffer (readback targets).
template <class T>
struct HBuf {
  T* p = nullptr; size_t n = 0;
  HBuf() = default;
  explicit HBuf(size_t count) { alloc(count); }
  HBuf(const HBuf&) = delete; HBuf& operator=(const HBuf&) = delete;
  ~HBuf() { if (p) cudaFreeHost(p); }
  void alloc(size_t count) { if (p) cudaFreeHost(p); p = nullptr; n = count; if (count) SK_CUDA(cudaMallocHost((void**)&p, count * sizeof(T))); }
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
