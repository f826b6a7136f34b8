// user_functor.cu — run-time compiled cost functors (see user_functor.cuh).
//
// sk_functor_register_source hands over the CUDA source of ONE functor template
//     template <class T> __device__ bool NAME(const double* consts, T const* const* x, T* residuals);
// (x[k] = parameter block k -- the `x: Array[T]*` of CostFunctor.apply, core/.../CostFunctor.scala:46-50).  The library wraps it
// into a translation unit made of its own jet.cuh (the device Jet<N> with spire's operator set), the evaluation argument
// blocks (eval_abi.cuh) and two kernels with the sizes baked in:
//     sk_user_evaluate_single   AutoDiffCostFunction.evaluate for one residual block (AutodiffCostFunction.scala:74-134)
//     sk_user_dense_evaluate    the residual blocks of this functor inside a DENSE_QR problem (dense_kernels.cu: k_dense_evaluate)
// and -- for a functor of the bundle-adjustment shape (2 residuals; blocks of 9 and 3; 2 constants), e.g. another camera model --
//     sk_user_ba_evaluate_jac / _cost   the TILE evaluation kernel of the Schur solvers (ba_evaluate.cuh: the body k_ba_evaluate
//                               runs for the built-in SnavelyReprojectionError), so that such a functor runs on the hot path
// compiles it with NVRTC for sm_100a (libnvrtc is loaded lazily with dlopen: no link-time dependency, and nothing is loaded
// unless a functor is registered) and loads the cubin with cudaLibraryLoadData.  Seeding order, Jacobian layout and the
// success convention are those of the built-in functors, so the reference's own AutodiffCostFuntionSpec vectors are reproduced
// exactly by functors given as strings (tests/test_gpu_parity.py::test_user_functors_from_source).
#include "user_functor.cuh"

#include <dlfcn.h>

#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "ba_dev.cuh"
#include "common.cuh"

namespace sk {

namespace {

#include "build/embedded_sources.inc"   // kSrcSkeresH, kSrcJetCuh, kSrcEvalAbi, kSrcBaTileObs, kSrcBaDev, kSrcBaTile, kSrcBaEvaluate

// ---- NVRTC through dlopen -----------------------------------------------------------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct NvrtcApi {
  void* h = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NvrtcApi& nvrtc() {
  static NvrtcApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
      a.h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (a.h) break;
    }
    if (!a.h) return;
    auto sym = [&](const char* n) { return dlsym(a.h, n); };
    a.CreateProgram = reinterpret_cast<decltype(a.CreateProgram)>(sym("nvrtcCreateProgram"));
    a.CompileProgram = reinterpret_cast<decltype(a.CompileProgram)>(sym("nvrtcCompileProgram"));
    a.GetCUBINSize = reinterpret_cast<decltype(a.GetCUBINSize)>(sym("nvrtcGetCUBINSize"));
    a.GetCUBIN = reinterpret_cast<decltype(a.GetCUBIN)>(sym("nvrtcGetCUBIN"));
    a.GetProgramLogSize = reinterpret_cast<decltype(a.GetProgramLogSize)>(sym("nvrtcGetProgramLogSize"));
    a.GetProgramLog = reinterpret_cast<decltype(a.GetProgramLog)>(sym("nvrtcGetProgramLog"));
    a.DestroyProgram = reinterpret_cast<decltype(a.DestroyProgram)>(sym("nvrtcDestroyProgram"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("nvrtcGetErrorString"));
  });
  SK_REQUIRE(a.h != nullptr && a.CreateProgram && a.CompileProgram && a.GetCUBINSize && a.GetCUBIN && a.GetProgramLogSize && a.GetProgramLog &&
                 a.DestroyProgram,
             SK_ERR_UNSUPPORTED, "libnvrtc.so.12 could not be loaded: functors given as source need the CUDA run-time compiler");
  return a;
}

struct Loaded { cudaLibrary_t lib = nullptr; cudaKernel_t single = nullptr, dense = nullptr, ba_jac = nullptr, ba_cost = nullptr; size_t ba_smem[2] = {0, 0}; };
struct UserFunctor {
  FunctorInfo info{};
  std::string name;
  std::vector<char> cubin;
  std::map<int, Loaded> per_device;            // the module is loaded on a device at first use
};
std::mutex g_mu;
std::map<int, std::unique_ptr<UserFunctor>> g_functors;
int g_next_id = kUserFunctorBase;

bool valid_identifier(const char* s) {
  if (s == nullptr || !(isalpha((unsigned char)s[0]) || s[0] == '_')) return false;
  for (const char* p = s; *p; ++p) if (!(isalnum((unsigned char)*p) || *p == '_')) return false;
  return true;
}

// The bundle-adjustment shape: SchurEliminator<2, 3, 9> with the observation as the two constants.
bool ba_shaped(const FunctorInfo& fi) { return fi.nres == 2 && fi.nblk == 2 && fi.sizes[0] == 9 && fi.sizes[1] == 3 && fi.nconsts == 2; }

// The translation unit handed to NVRTC.
std::string generate(const char* name, const char* source, const FunctorInfo& fi) {
  std::string s;
  s += "#include \"../../include/skeres.h\"\n#include \"eval_abi.cuh\"\n#include \"jet.cuh\"\n";
  if (ba_shaped(fi)) s += "#include \"ba_evaluate.cuh\"\n";
  s += "using sk::Jet;\n";
  s += "#line 1 \"user_functor_source\"\n";
  s += source;
  s += "\n#line 1 \"skeres_generated\"\nnamespace sk_user {\n";
  s += fmt("constexpr int NRES = %d, NBLK = %d, NTOT = %d;\n", fi.nres, fi.nblk, fi.ntot);
  s += "__device__ constexpr int SIZES[NBLK] = {";
  for (int k = 0; k < fi.nblk; ++k) s += fmt("%s%d", k ? ", " : "", fi.sizes[k]);
  s += "};\n__device__ constexpr int OFFS[NBLK] = {";
  for (int k = 0, o = 0; k < fi.nblk; o += fi.sizes[k], ++k) s += fmt("%s%d", k ? ", " : "", o);
  s += "};\n";
  // x: the concatenated parameter values; jac: nullptr (residuals only, AutodiffCostFunction.scala:80) or NRES x NTOT row-major over
  // the concatenated parameters.  Seeding: Jet(x_i, i) per scalar parameter in block order (:95-107).
  s += std::string("__device__ __forceinline__ bool evaluate(const double* c, const double* x, double* res, double* jac) {\n"
                   "  if (jac == nullptr) {\n"
                   "    const double* xp[NBLK];\n"
                   "    for (int k = 0; k < NBLK; ++k) xp[k] = x + OFFS[k];\n"
                   "    return ::") + name + "<double>(c, xp, res);\n"
       "  }\n"
       "  typedef sk::Jet<NTOT> J;\n"
       "  J jx[NTOT], jr[NRES];\n"
       "  const J* xp[NBLK];\n"
       "  for (int i = 0; i < NTOT; ++i) jx[i] = J(x[i], i);\n"
       "  for (int k = 0; k < NBLK; ++k) xp[k] = jx + OFFS[k];\n"
       "  if (!::" + name + "<J>(c, xp, jr)) return false;\n"
       "  for (int q = 0; q < NRES; ++q) { res[q] = jr[q].a; for (int i = 0; i < NTOT; ++i) jac[q * NTOT + i] = jr[q].v[i]; }\n"
       "  return true;\n"
       "}\n"
       "}  // namespace sk_user\n";
  s += R"SKGEN(
extern "C" __global__ void sk_user_evaluate_single(sk::EvalArgs a, int* ok_out) {
  using namespace sk_user;
  double xx[NTOT], res[NRES], jac[NRES * NTOT];
  int t = 0;
  for (int k = 0; k < NBLK; ++k)
    for (int c = 0; c < SIZES[k]; ++c) xx[t++] = a.params[k][c];
  const bool ok = evaluate(a.consts, xx, res, a.has_jac ? jac : nullptr);
  *ok_out = ok ? 1 : 0;
  if (!ok) return;
  for (int q = 0; q < NRES; ++q) a.residuals[q] = res[q];
  if (!a.has_jac) return;
  for (int k = 0; k < NBLK; ++k) {
    if (a.jac[k] == nullptr) continue;
    int col = 0;
    for (int q = 0; q < NRES; ++q)
      for (int p = 0; p < SIZES[k]; ++p) a.jac[k][col++] = jac[q * NTOT + OFFS[k] + p];
  }
}

// 128 threads per CTA, thread i of the grid = residual block i (the slots of other functors are left alone).
extern "C" __global__ void sk_user_dense_evaluate(int functor_id, int with_jac, int nrb, const sk::DenseRb* __restrict__ rbs,
                                                  const double* __restrict__ x, double* __restrict__ J, int m, double* __restrict__ b,
                                                  double* __restrict__ block_cost, int* fail_flag, const int* guard) {
  using namespace sk_user;
  if (guard != nullptr && *guard == 0) return;
  __shared__ double red[128];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double cost = 0.0;
  if (i < nrb && rbs[i].functor == functor_id) {
    const sk::DenseRb rb = rbs[i];
    double xx[NTOT], res[NRES], jac[NRES * NTOT];
    int t = 0;
    for (int k = 0; k < NBLK; ++k)
      for (int c = 0; c < SIZES[k]; ++c) xx[t++] = x[rb.col[k] + c];
    const bool ok = evaluate(rb.consts, xx, res, with_jac ? jac : nullptr);
    if (!ok) atomicOr(fail_flag, 1);
    double sq = 0.0;
    for (int q = 0; q < NRES; ++q) sq += res[q] * res[q];
    double rho[3];
    sk::LossSpec ls{rb.loss_type, rb.loss_a, rb.loss_b};
    sk::loss_evaluate(ls, sq, rho);
    cost = 0.5 * rho[0];
    if (!(cost == cost)) atomicOr(fail_flag, 1);
    if (with_jac) {
      const sk::Corrector corr(sq, rho);
      corr.correct_jacobian(NRES, NTOT, NTOT, res, jac);
      corr.correct_residuals(NRES, res);
      for (int k = 0; k < NBLK; ++k)
        for (int c = 0; c < SIZES[k]; ++c)
          for (int q = 0; q < NRES; ++q) J[(size_t)(rb.col[k] + c) * m + rb.row + q] = jac[q * NTOT + OFFS[k] + c];
      for (int q = 0; q < NRES; ++q) b[rb.row + q] = res[q];
    }
  }
  red[threadIdx.x] = cost;                 // fixed-order sum over the CTA: deterministic
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int k = 0; k < (int)blockDim.x; ++k) tot += red[k];
    block_cost[blockIdx.x] += tot;
  }
}
)SKGEN";
  if (ba_shaped(fi)) s += R"SKGEN(
// The functor as the tile kernels of the Schur solvers see it (ba_evaluate.cuh): constants = the observed (x, y), blocks = (camera 9, point 3).
namespace sk_user {
struct BaFunctor {
  static __device__ __forceinline__ bool residual(const double* cam, const double* pt, double ox, double oy, double* res) {
    const double c[2] = {ox, oy};
    double xx[12];
    for (int k = 0; k < 9; ++k) xx[k] = cam[k];
    for (int k = 0; k < 3; ++k) xx[9 + k] = pt[k];
    return evaluate(c, xx, res, nullptr);
  }
  static __device__ __forceinline__ bool residual_jacobian(const double* cam, const double* pt, double ox, double oy, double* res,
                                                           double* F, double* E) {
    const double c[2] = {ox, oy};
    double xx[12], jac[24];
    for (int k = 0; k < 9; ++k) xx[k] = cam[k];
    for (int k = 0; k < 3; ++k) xx[9 + k] = pt[k];
    if (!evaluate(c, xx, res, jac)) return false;
    for (int q = 0; q < 2; ++q) {
      for (int k = 0; k < 9; ++k) F[q * 9 + k] = jac[q * 12 + k];
      for (int k = 0; k < 3; ++k) E[q * 3 + k] = jac[q * 12 + 9 + k];
    }
    return true;
  }
};
}  // namespace sk_user
extern "C" __global__ void __launch_bounds__(sk::kTileObs, 2)
sk_user_ba_evaluate_jac(sk::BaDev L, const double* __restrict__ x, const double* __restrict__ scale, sk::LossSpec loss, int write_j,
                        double2* __restrict__ J2, double2* __restrict__ r2, double* __restrict__ grad, double* __restrict__ cnorm2,
                        double* __restrict__ seg_g, double* __restrict__ seg_n, double* __restrict__ tile_cost,
                        double* __restrict__ chunk_pt, int* fail_flag, const int* guard) {
  sk::ba_evaluate_tile<true, sk_user::BaFunctor>(L, x, scale, loss, write_j, J2, r2, grad, cnorm2, seg_g, seg_n, tile_cost, chunk_pt, fail_flag, guard);
}
extern "C" __global__ void __launch_bounds__(sk::kTileObs)
sk_user_ba_evaluate_cost(sk::BaDev L, const double* __restrict__ x, const double* __restrict__ scale, sk::LossSpec loss, int write_j,
                         double2* __restrict__ J2, double2* __restrict__ r2, double* __restrict__ grad, double* __restrict__ cnorm2,
                         double* __restrict__ seg_g, double* __restrict__ seg_n, double* __restrict__ tile_cost,
                         double* __restrict__ chunk_pt, int* fail_flag, const int* guard) {
  sk::ba_evaluate_tile<false, sk_user::BaFunctor>(L, x, scale, loss, write_j, J2, r2, grad, cnorm2, seg_g, seg_n, tile_cost, chunk_pt, fail_flag, guard);
}
)SKGEN";
  return s;
}

Loaded& loaded_on_current_device(UserFunctor& f) {
  int dev = 0;
  SK_CUDA(cudaGetDevice(&dev));
  auto it = f.per_device.find(dev);
  if (it != f.per_device.end()) return it->second;
  Loaded l;
  SK_CUDA(cudaLibraryLoadData(&l.lib, f.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  SK_CUDA(cudaLibraryGetKernel(&l.single, l.lib, "sk_user_evaluate_single"));
  SK_CUDA(cudaLibraryGetKernel(&l.dense, l.lib, "sk_user_dense_evaluate"));
  if (ba_shaped(f.info)) {
    SK_CUDA(cudaLibraryGetKernel(&l.ba_jac, l.lib, "sk_user_ba_evaluate_jac"));
    SK_CUDA(cudaLibraryGetKernel(&l.ba_cost, l.lib, "sk_user_ba_evaluate_cost"));
  }
  return f.per_device.emplace(dev, l).first->second;
}

UserFunctor& find(int id) {
  auto it = g_functors.find(id);
  SK_REQUIRE(it != g_functors.end(), SK_ERR_INVALID_ARGUMENT, "functor id %d is not a registered run-time functor", id);
  return *it->second;
}

}  // namespace

int register_user_functor(const char* name, const char* source, int nres, int nblk, const int* sizes, int nconsts) {
  SK_REQUIRE(valid_identifier(name), SK_ERR_INVALID_ARGUMENT, "sk_functor_register_source: the functor name must be a C identifier");
  SK_REQUIRE(source != nullptr && sizes != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_functor_register_source: null argument");
  SK_REQUIRE(nres >= 1 && nres <= 16 && nblk >= 1 && nblk <= SK_MAX_PARAMETER_BLOCKS && nconsts >= 0 && nconsts <= SK_MAX_CONSTS,
             SK_ERR_INVALID_ARGUMENT, "sk_functor_register_source: 1..16 residuals, 1..%d parameter blocks, 0..%d constants", SK_MAX_PARAMETER_BLOCKS, SK_MAX_CONSTS);
  auto f = std::make_unique<UserFunctor>();
  f->name = name;
  FunctorInfo& fi = f->info;
  fi.nres = nres; fi.nblk = nblk; fi.nconsts = nconsts; fi.ntot = 0;
  for (int k = 0; k < SK_MAX_PARAMETER_BLOCKS; ++k) fi.sizes[k] = 0;
  for (int k = 0; k < nblk; ++k) {
    SK_REQUIRE(sizes[k] >= 1 && sizes[k] <= 32, SK_ERR_INVALID_ARGUMENT, "sk_functor_register_source: block sizes must be in 1..32");
    fi.sizes[k] = sizes[k]; fi.ntot += sizes[k];
  }
  SK_REQUIRE(fi.ntot <= 32, SK_ERR_INVALID_ARGUMENT, "sk_functor_register_source: at most 32 scalar parameters per residual block (Jet<32>)");
  const std::string tu = generate(name, source, fi);
  NvrtcApi& rt = nvrtc();
  const char* header_names[] = {"../../include/skeres.h", "eval_abi.cuh", "jet.cuh", "stdint.h", "stddef.h", "math.h",
                                "ba_tile_obs.h", "ba_dev.cuh", "ba_tile.cuh", "ba_evaluate.cuh"};
  const char* header_srcs[] = {kSrcSkeresH, kSrcEvalAbi, kSrcJetCuh,
                               "typedef signed char int8_t; typedef short int16_t; typedef int int32_t; typedef long long int64_t;\n"
                               "typedef unsigned char uint8_t; typedef unsigned short uint16_t; typedef unsigned int uint32_t; typedef unsigned long long uint64_t;\n",
                               "", "", kSrcBaTileObs, kSrcBaDev, kSrcBaTile, kSrcBaEvaluate};
  nvrtcProgram prog = nullptr;
  int r = rt.CreateProgram(&prog, tu.c_str(), "skeres_user_functor.cu", 10, header_srcs, header_names);
  SK_REQUIRE(r == 0, SK_ERR_INTERNAL, "nvrtcCreateProgram failed: %s", rt.GetErrorString ? rt.GetErrorString(r) : "?");
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device", "--fmad=true"};
  r = rt.CompileProgram(prog, 4, opts);
  std::string log;
  size_t log_n = 0;
  if (rt.GetProgramLogSize(prog, &log_n) == 0 && log_n > 1) { log.resize(log_n); rt.GetProgramLog(prog, &log[0]); }
  if (r != 0) {
    rt.DestroyProgram(&prog);
    throw Error(SK_ERR_INVALID_ARGUMENT, fmt("functor '%s' does not compile (%s):\n%s", name, rt.GetErrorString ? rt.GetErrorString(r) : "?", log.c_str()));
  }
  size_t n = 0;
  r = rt.GetCUBINSize(prog, &n);
  if (r == 0 && n > 0) { f->cubin.resize(n); r = rt.GetCUBIN(prog, f->cubin.data()); }
  rt.DestroyProgram(&prog);
  SK_REQUIRE(r == 0 && !f->cubin.empty(), SK_ERR_INTERNAL, "nvrtcGetCUBIN failed for functor '%s'", name);
  if (const char* dir = getenv("SKERES_DUMP_USER_CUBIN")) {   // development: <dir>/<name>.cubin for cuobjdump -res-usage / -sass
    if (FILE* fp = fopen(fmt("%s/%s.cubin", dir, name).c_str(), "wb")) { fwrite(f->cubin.data(), 1, f->cubin.size(), fp); fclose(fp); }
  }
  std::lock_guard<std::mutex> g(g_mu);
  const int id = g_next_id++;
  fi.id = id;
  g_functors[id] = std::move(f);
  return id;
}

bool user_functor_info(int id, FunctorInfo* out) {
  std::lock_guard<std::mutex> g(g_mu);
  auto it = g_functors.find(id);
  if (it == g_functors.end()) return false;
  *out = it->second->info;
  return true;
}

void launch_user_evaluate_single(int id, const EvalArgs& a, int* ok_out, cudaStream_t s) {
  std::lock_guard<std::mutex> g(g_mu);
  Loaded& l = loaded_on_current_device(find(id));
  EvalArgs args = a;
  void* params[] = {&args, &ok_out};
  SK_CUDA(cudaLaunchKernel(reinterpret_cast<const void*>(l.single), dim3(1), dim3(1), params, 0, s));
}

bool user_functor_runs_on_tiles(int id) {
  std::lock_guard<std::mutex> g(g_mu);
  auto it = g_functors.find(id);
  return it != g_functors.end() && ba_shaped(it->second->info);
}

void launch_user_ba_evaluate(int id, bool with_jacobian, size_t smem, const BaDev& L, const double* x, const double* scale, LossSpec loss,
                             int write_j, double2* J2, double2* r2, double* grad, double* cnorm2, double* seg_g, double* seg_n,
                             double* tile_cost, double* chunk_pt, int* fail_flag, const int* guard, cudaStream_t s) {
  std::lock_guard<std::mutex> g(g_mu);
  UserFunctor& f = find(id);
  SK_REQUIRE(ba_shaped(f.info), SK_ERR_UNSUPPORTED, "functor %d does not have the bundle-adjustment shape (2; 9, 3; 2 constants)", id);
  Loaded& l = loaded_on_current_device(f);
  cudaKernel_t k = with_jacobian ? l.ba_jac : l.ba_cost;
  size_t& cfg = l.ba_smem[with_jacobian ? 1 : 0];
  if (smem > 48 * 1024 && smem > cfg) {
    SK_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(k), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfg = smem;
  }
  BaDev Lc = L;
  void* params[] = {&Lc, &x, &scale, &loss, &write_j, &J2, &r2, &grad, &cnorm2, &seg_g, &seg_n, &tile_cost, &chunk_pt, &fail_flag, &guard};
  SK_CUDA(cudaLaunchKernel(reinterpret_cast<const void*>(k), dim3(L.n_tiles), dim3(kTileObs), params, smem, s));
}

void launch_user_dense_evaluate(int id, bool with_jacobian, int nrb, const DenseRb* rbs, const double* x, double* J, int m, double* b,
                                double* block_cost, int* fail_flag, const int* guard, cudaStream_t s) {
  std::lock_guard<std::mutex> g(g_mu);
  Loaded& l = loaded_on_current_device(find(id));
  int wj = with_jacobian ? 1 : 0;
  void* params[] = {&id, &wj, &nrb, &rbs, &x, &J, &m, &b, &block_cost, &fail_flag, &guard};
  SK_CUDA(cudaLaunchKernel(reinterpret_cast<const void*>(l.dense), dim3(cdiv(nrb, 128)), dim3(128), params, 0, s));
}

}  // namespace sk
