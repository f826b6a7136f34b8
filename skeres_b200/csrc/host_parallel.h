// host_parallel.h — static-chunk parallel loop on the host (ingestion and layout construction).
#pragma once
#include <algorithm>
#include <cstdint>
#include <exception>
#include <memory>
#include <type_traits>
#include <utility>
#include <mutex>
#include <thread>
#include <vector>

namespace sk {

// Allocator whose resize() leaves new elements uninitialised: a vector that is filled by a parallel copy right after resize()
// must not be zeroed by one thread first (80 MB of page faults and stores at 5M residual blocks: 30 ms of the end-to-end path).
template <class T>
struct DefaultInitAllocator : std::allocator<T> {
  template <class U> struct rebind { using other = DefaultInitAllocator<U>; };
  using std::allocator<T>::allocator;
  template <class U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void*>(p)) U; }
  template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
template <class T> using RawVector = std::vector<T, DefaultInitAllocator<T>>;

// Static-chunk parallel loop over [0, n) on the host (results do not depend on the thread count).
template <class F>
inline void parallel_for(int64_t n, F f) {
  int nt = (int)std::min<int64_t>(std::max(1u, std::thread::hardware_concurrency()), 32);
  nt = (int)std::min<int64_t>(nt, std::max<int64_t>(1, n / 64));
  if (nt <= 1) { f(0, n); return; }
  std::vector<std::thread> th;
  std::exception_ptr err;
  std::mutex mu;
  for (int k = 0; k < nt; ++k) {
    const int64_t a = n * k / nt, b = n * (k + 1) / nt;
    th.emplace_back([&, a, b] { try { f(a, b); } catch (...) { std::lock_guard<std::mutex> g(mu); err = std::current_exception(); } });
  }
  for (auto& t : th) t.join();
  if (err) std::rethrow_exception(err);
}


}  // namespace sk
