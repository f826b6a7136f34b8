// host_parallel.h — static-chunk parallel loop on the host (ingestion and layout construction).
#pragma once
#include <algorithm>
#include <cstdint>
#include <exception>
#include <mutex>
#include <thread>
#include <vector>

namespace sk {

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
