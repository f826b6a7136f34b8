// Development micro-benchmark (not part of the product): what read bandwidth can each load mechanism deliver for the
// matvec kernel's access pattern -- per 256-observation tile, 12 separate 4 KB runs (one per Jacobian plane), trivial
// math, almost no writes?  Build: make -C tools/membench ; run on a B200: tools/membench/mvprobe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
constexpr int T = 256, P = 12;

__device__ __forceinline__ void sink(double s, double* out, size_t i) { if (s == 123.456789) out[i] = s; }

// K1: flat grid-stride stream over the whole buffer
__global__ void __launch_bounds__(256) k_flat(const double2* __restrict__ J, size_t n, double* out) {
  double s = 0;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; const size_t st = (size_t)gridDim.x * blockDim.x;
  for (; i + 7 * st < n; i += 8 * st) {
    double2 a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = __ldcs(J + i + k * st);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k].x + a[k].y;
  }
  for (; i < n; i += st) { double2 a = __ldcs(J + i); s += a.x + a.y; }
  sink(s, out, threadIdx.x);
}

// K2: one CTA per tile, 12 plane loads per thread (what k_ba_matvec does); occupancy set through dynamic smem
__global__ void __launch_bounds__(T) k_tile_ldg(const double2* __restrict__ J, size_t O, double* out) {
  extern __shared__ double sm[];
  const size_t i = (size_t)blockIdx.x * T + threadIdx.x;
  double2 a[P];
#pragma unroll
  for (int k = 0; k < P; ++k) a[k] = __ldcs(J + k * O + i);
  double s = 0;
#pragma unroll
  for (int k = 0; k < P; ++k) s += a[k].x + a[k].y;
  if (s == 123.456789) sm[threadIdx.x] = s;
  sink(s, out, i);
}

__device__ __forceinline__ void cp_async16(void* d, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(d)), "l"(g) : "memory");
}
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// K3: persistent, per-thread cp.async into a ring of S tile buffers (each thread copies and reads its own 12 vectors)
template <int S>
__global__ void __launch_bounds__(T) k_tile_cpasync(const double2* __restrict__ J, size_t O, int n_tiles, double* out) {
  extern __shared__ double2 buf[];     // [S][P][T]
  const int tid = threadIdx.x;
  const int my = ((int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto issue = [&](int it) {
    if (it < my) {
      const size_t i = (size_t)(blockIdx.x + it * gridDim.x) * T + tid;
#pragma unroll
      for (int k = 0; k < P; ++k) cp_async16(buf + ((it % S) * P + k) * T + tid, J + k * O + i);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  for (int it = 0; it < S - 1; ++it) issue(it);
  double s = 0;
  for (int it = 0; it < my; ++it) {
    issue(it + S - 1);                 // slot (it-1)%S: consumed by this thread in the previous iteration
    cp_wait<S - 1>();
#pragma unroll
    for (int k = 0; k < P; ++k) { const double2 a = buf[((it % S) * P + k) * T + tid]; s += a.x + a.y; }
  }
  sink(s, out, tid);
}

// K4: persistent, TMA bulk copies (12 x 4 KB per tile) issued by one thread into a ring of S stages, mbarrier completion
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
               "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}

template <int S>
__global__ void __launch_bounds__(T) k_tile_tma(const double2* __restrict__ J, size_t O, int n_tiles, double* out) {
  extern __shared__ __align__(128) unsigned char raw[];
  double2* buf = reinterpret_cast<double2*>(raw);                                  // [S][P][T]
  uint64_t* full = reinterpret_cast<uint64_t*>(raw + (size_t)S * P * T * sizeof(double2));   // [S]
  const int tid = threadIdx.x;
  const int my = ((int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(full + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int it) {           // thread 0 only
    if (it < my) {
      const int st = it % S;
      const size_t i = (size_t)(blockIdx.x + it * gridDim.x) * T;
      mbar_expect_tx(full + st, P * T * (unsigned)sizeof(double2));
#pragma unroll
      for (int k = 0; k < P; ++k) bulk_g2s(buf + (st * P + k) * T, J + k * O + i, T * (unsigned)sizeof(double2), full + st);
    }
  };
  if (tid == 0) for (int it = 0; it < S - 1; ++it) issue(it);
  double s = 0;
  for (int it = 0; it < my; ++it) {
    const int st = it % S;
    __syncthreads();                   // everyone is done with slot (it-1)%S
    if (tid == 0) issue(it + S - 1);
    mbar_wait(full + st, (unsigned)((it / S) & 1));
#pragma unroll
    for (int k = 0; k < P; ++k) { const double2 a = buf[(st * P + k) * T + tid]; s += a.x + a.y; }
  }
  sink(s, out, tid);
}


// Latency probes (one warp): dependent chains of N operations, cycles per operation by clock64().
__global__ void k_latency(double* out, long long* cyc, int N) {
  __shared__ double sh[1024];
  __shared__ unsigned short perm[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) { sh[i] = 1.0 + i * 1e-9; perm[i] = (unsigned short)((i * 37) & 1023); }
  __syncwarp();
  double a = out[0], b = out[1];
  long long t0 = clock64();
  for (int i = 0; i < N; ++i) a += b;                                   // DADD chain
  long long t1 = clock64();
  for (int i = 0; i < N; ++i) a = fma(a, b, b);                         // DFMA chain
  long long t2 = clock64();
  for (int i = 0; i < N; ++i) a += sh[(i + threadIdx.x) & 1023];        // independent LDS feeding a DADD chain (no unroll hints)
  long long t3 = clock64();
  for (int i = 0; i < N; ++i) a += sh[perm[(i + threadIdx.x) & 1023]];  // LDS index -> LDS value -> DADD (segment-sum pattern)
  long long t4 = clock64();
  float f = (float)out[2], g = (float)out[3];
  for (int i = 0; i < N; ++i) f += g;                                   // FADD chain for comparison
  long long t5 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
  out[4 + threadIdx.x] = a + f;
}

template <class F>
static double time_ms(F launch, int reps = 20) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  CK(cudaGetLastError());
  float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main() {
  const int n_tiles = 19777; const size_t O = (size_t)n_tiles * T; const size_t n = O * P;
  double2* J; double* out; CK(cudaMalloc(&J, n * sizeof(double2))); CK(cudaMalloc(&out, O * sizeof(double)));
  CK(cudaMemset(J, 0, n * sizeof(double2)));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  {
    long long* cyc; CK(cudaMalloc(&cyc, 8 * sizeof(long long))); CK(cudaMemset(out, 0, 64 * sizeof(double)));
    const int N = 4096;
    k_latency<<<1, 32>>>(out, cyc, N); CK(cudaDeviceSynchronize());
    long long h[5]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    printf("latency per dependent op (cycles): DADD %.1f  DFMA %.1f  LDS+DADD %.1f  LDS->LDS->DADD %.1f  FADD %.1f\n",
           h[0] / (double)N, h[1] / (double)N, h[2] / (double)N, h[3] / (double)N, h[4] / (double)N);
    if (getenv("MVPROBE_LATENCY_ONLY")) return 0;
  }
  const double gb = n * sizeof(double2) / 1e9;
  printf("bytes per pass %.3f GB, %d SMs\n", gb, sms);
  auto report = [&](const char* name, double ms) { printf("%-44s %8.4f ms  %7.1f GB/s\n", name, ms, gb / ms * 1e3); fflush(stdout); };
  for (int per : {4, 8, 16}) { char nm[64]; snprintf(nm, 64, "flat grid-stride, %d CTAs/SM", per); report(nm, time_ms([&] { k_flat<<<sms * per, 256>>>(J, n, out); })); }
  for (int per : {1, 2, 3, 4, 6, 8}) {
    const size_t smem = per == 8 ? 0 : (size_t)(227 * 1024 / per - 1024);
    CK(cudaFuncSetAttribute(k_tile_ldg, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    char nm[64]; snprintf(nm, 64, "tile LDG x12, %d CTAs/SM (smem-capped)", per);
    report(nm, time_ms([&] { k_tile_ldg<<<n_tiles, T, smem>>>(J, O, out); }));
  }
#define CPA(S_, PER_) { const size_t smem = (size_t)S_ * P * T * sizeof(double2); \
    CK(cudaFuncSetAttribute(k_tile_cpasync<S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    char nm[64]; snprintf(nm, 64, "persistent cp.async, %d stages x %d CTAs/SM", S_, PER_); \
    report(nm, time_ms([&] { k_tile_cpasync<S_><<<sms * PER_, T, smem>>>(J, O, n_tiles, out); })); }
  CPA(2, 2) CPA(2, 1) CPA(4, 1) CPA(1, 4) CPA(3, 1)
#define TMA(S_, PER_) { const size_t smem = (size_t)S_ * P * T * sizeof(double2) + 64; \
    CK(cudaFuncSetAttribute(k_tile_tma<S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    char nm[64]; snprintf(nm, 64, "persistent TMA bulk, %d stages x %d CTAs/SM", S_, PER_); \
    report(nm, time_ms([&] { k_tile_tma<S_><<<sms * PER_, T, smem>>>(J, O, n_tiles, out); })); }
  TMA(2, 2) TMA(2, 1) TMA(3, 1) TMA(4, 1) TMA(1, 4)
  return 0;
}
