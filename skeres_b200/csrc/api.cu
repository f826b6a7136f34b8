// api.cu — the extern "C" surface declared in include/skeres.h.  Nothing but plain pointers and
// sizes crosses this boundary; every failure becomes an sk_status + sk_last_error().
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "ba_solver.cuh"
#include "host_parallel.h"
#include "batch.cuh"
#include "comm.cuh"
#include "dense_solver.cuh"
#include "jet.cuh"
#include "user_functor.cuh"

using namespace sk;

// ---- handle types -----------------------------------------------------------------------------------
struct sk_double_array { DBuf<double> d; int64_t n = 0; };
struct sk_loss_function { LossSpec spec; };
struct sk_cost_function { int functor_id; FunctorInfo info; double consts[SK_MAX_CONSTS]; };

struct ResidualGroup {              // n residual blocks of one functor / loss over one array per block slot
  int functor_id; FunctorInfo info; LossSpec loss;
  std::vector<sk_double_array*> arrays;   // per residual block per parameter block (size n * nblk)
  RawVector<double> consts;                // n * nconsts   (RawVector: filled by parallel copies, never zeroed first)
  RawVector<int64_t> offsets;              // n * nblk
  int64_t n = 0;
};
struct DeclaredBlocks { sk_double_array* array; int32_t size; std::vector<int64_t> offsets; };   // AddParameterBlock
struct sk_problem { std::vector<ResidualGroup> groups; std::vector<DeclaredBlocks> declared; int64_t num_residual_blocks = 0, num_residuals = 0; };

struct sk_bal_problem {
  int32_t n_cam = 0, n_pt = 0, n_obs = 0;
  std::vector<int32_t> cam_idx, pt_idx;
  std::vector<double> obs;
  sk_double_array* params = nullptr;
};

namespace {

thread_local std::string g_last_error;
thread_local int g_device = 0;
std::string g_log_name;

int fail(int status, const std::string& msg) { g_last_error = msg; return status; }

#define SK_API_BEGIN try {
#define SK_API_END                                                          \
  return SK_OK;                                                             \
  }                                                                         \
  catch (const Error& e) { return fail(e.status, e.what()); }               \
  catch (const std::bad_alloc&) { return fail(SK_ERR_INTERNAL, "out of host memory"); } \
  catch (const std::exception& e) { return fail(SK_ERR_INTERNAL, e.what()); }

void ensure_device() {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    throw Error(SK_ERR_CUDA, "no CUDA device available: libskeres has no CPU fallback");
  }
  SK_CUDA(cudaSetDevice(g_device));
}

double wall() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Sizes of a functor: one of the built-in device functors (jet.cuh) or one compiled at run time from source (user_functor.cu).
bool lookup_functor(int id, FunctorInfo* fi) { return id >= kUserFunctorBase ? user_functor_info(id, fi) : functor_info(id, fi); }
const char* const kNoFunctor = "functor id %d is neither a built-in device functor nor one registered from source (sk_functor_register_source); "
                               "arbitrary JVM functors cannot run on the GPU and there is no CPU fallback";

// AutoDiffCostFunction.evaluate (AutodiffCostFunction.scala:74-134) for one residual block.
__global__ void k_evaluate_single(EvalArgs a, int* ok_out) {
  FunctorInfo fi;
  functor_info(a.functor, &fi);
  double xx[SK_MAX_TOTAL_PARAMS], res[SK_MAX_RESIDUALS], jac[SK_MAX_RESIDUALS * SK_MAX_TOTAL_PARAMS];
  int t = 0;
  for (int k = 0; k < fi.nblk; ++k)
    for (int c = 0; c < fi.sizes[k]; ++c) xx[t++] = a.params[k][c];
  const bool ok = evaluate_functor(a.functor, a.consts, xx, res, a.has_jac ? jac : nullptr);
  *ok_out = ok ? 1 : 0;
  if (!ok) return;
  for (int q = 0; q < fi.nres; ++q) a.residuals[q] = res[q];            // :91 / :113
  if (!a.has_jac) return;
  int off = 0;
  for (int k = 0; k < fi.nblk; ++k) {                                   // :115-130
    const int ni = fi.sizes[k];
    if (a.jac[k] != nullptr) {
      int col = 0;
      for (int q = 0; q < fi.nres; ++q)
        for (int p = 0; p < ni; ++p) a.jac[k][col++] = jac[q * fi.ntot + off + p];
    }
    off += ni;
  }
}

__global__ void k_loss_evaluate(LossSpec l, double s, double* rho) { loss_evaluate(l, s, rho); }

void check_block(const sk_double_array* a, int64_t off, int size, const char* what) {
  SK_REQUIRE(a != nullptr, SK_ERR_INVALID_ARGUMENT, "%s: null DoubleArray", what);
  SK_REQUIRE(off >= 0 && off + size <= a->n, SK_ERR_INVALID_ARGUMENT, "%s: block [%lld, %lld) outside array of %lld doubles", what,
             (long long)off, (long long)(off + size), (long long)a->n);
}

}  // namespace

extern "C" {

const char* sk_last_error(void) { return g_last_error.c_str(); }
int sk_abi_version(void) { return SKERES_ABI_VERSION; }
int sk_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
  return count;
}
int sk_set_device(int ordinal) {
  SK_API_BEGIN
  int count = sk_device_count();
  SK_REQUIRE(ordinal >= 0 && ordinal < count, SK_ERR_CUDA, "device %d not available (%d visible)", ordinal, count);
  g_device = ordinal;
  SK_CUDA(cudaSetDevice(ordinal));
  SK_API_END
}
void sk_init_google_logging(const char* name) { g_log_name = name ? name : ""; }

// ---- DoubleArray -------------------------------------------------------------------------------------
int sk_double_array_create(int64_t n, sk_double_array** out) {
  SK_API_BEGIN
  SK_REQUIRE(out != nullptr && n >= 0, SK_ERR_INVALID_ARGUMENT, "sk_double_array_create: bad arguments");
  ensure_device();
  std::unique_ptr<sk_double_array> a(new sk_double_array);
  a->n = n; a->d.alloc((size_t)std::max<int64_t>(n, 1));
  SK_CUDA(cudaMemset(a->d.p, 0, sizeof(double) * (size_t)std::max<int64_t>(n, 1)));
  *out = a.release();
  SK_API_END
}
int sk_double_array_destroy(sk_double_array* a) { SK_API_BEGIN delete a; SK_API_END }
int64_t sk_double_array_size(const sk_double_array* a) { return a ? a->n : -1; }
int sk_double_array_upload(sk_double_array* a, int64_t offset, const double* host, int64_t n) {
  SK_API_BEGIN
  check_block(a, offset, (int)0, "sk_double_array_upload");
  SK_REQUIRE(n >= 0 && offset + n <= a->n && (host != nullptr || n == 0), SK_ERR_INVALID_ARGUMENT, "sk_double_array_upload: range outside array");
  ensure_device();
  if (n) SK_CUDA(cudaMemcpy(a->d.p + offset, host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
  SK_API_END
}
int sk_double_array_download(const sk_double_array* a, int64_t offset, double* host, int64_t n) {
  SK_API_BEGIN
  check_block(a, offset, 0, "sk_double_array_download");
  SK_REQUIRE(n >= 0 && offset + n <= a->n && (host != nullptr || n == 0), SK_ERR_INVALID_ARGUMENT, "sk_double_array_download: range outside array");
  ensure_device();
  if (n) SK_CUDA(cudaMemcpy(host, a->d.p + offset, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
  SK_API_END
}
int sk_double_array_get(const sk_double_array* a, int64_t i, double* out) { return sk_double_array_download(a, i, out, 1); }
int sk_double_array_set(sk_double_array* a, int64_t i, double value) { return sk_double_array_upload(a, i, &value, 1); }
void* sk_double_array_device_ptr(sk_double_array* a) { return a ? a->d.p : nullptr; }
int sk_double_array_copy(sk_double_array* dst, int64_t dst_offset, const sk_double_array* src, int64_t src_offset, int64_t n) {
  SK_API_BEGIN
  check_block(dst, dst_offset, 0, "sk_double_array_copy");
  check_block(src, src_offset, 0, "sk_double_array_copy");
  SK_REQUIRE(n >= 0 && dst_offset + n <= dst->n && src_offset + n <= src->n, SK_ERR_INVALID_ARGUMENT, "sk_double_array_copy: range outside array");
  ensure_device();
  if (n) SK_CUDA(cudaMemcpy(dst->d.p + dst_offset, src->d.p + src_offset, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice));
  SK_API_END
}

// ---- LossFunction ------------------------------------------------------------------------------------
static int make_loss(int type, double a, sk_loss_function** out, double b = 0.0) {
  SK_API_BEGIN
  SK_REQUIRE(out != nullptr, SK_ERR_INVALID_ARGUMENT, "loss factory: null output");
  if (type == SK_LOSS_TOLERANT) SK_REQUIRE(a >= 0.0 && b > 0.0, SK_ERR_INVALID_ARGUMENT, "TolerantLoss needs a >= 0 and b > 0");   // Ceres CHECKs
  else SK_REQUIRE(type == SK_LOSS_TRIVIAL || a > 0.0, SK_ERR_INVALID_ARGUMENT, "loss scale must be positive");
  *out = new sk_loss_function{LossSpec{type, a, b}};
  SK_API_END
}
int sk_loss_trivial(sk_loss_function** out) { return make_loss(SK_LOSS_TRIVIAL, 0.0, out); }
int sk_loss_huber(double a, sk_loss_function** out) { return make_loss(SK_LOSS_HUBER, a, out); }
int sk_loss_cauchy(double a, sk_loss_function** out) { return make_loss(SK_LOSS_CAUCHY, a, out); }
int sk_loss_soft_l_one(double, sk_loss_function**) { return fail(SK_ERR_UNSUPPORTED, "SoftLOneLoss has no device implementation (ceres.i:172); registered losses: trivial, huber, cauchy, tolerant"); }
int sk_loss_tukey(double, sk_loss_function**) { return fail(SK_ERR_UNSUPPORTED, "TukeyLoss has no device implementation (ceres.i:174); registered losses: trivial, huber, cauchy, tolerant"); }
int sk_loss_tolerant(double a, double b, sk_loss_function** out) { return make_loss(SK_LOSS_TOLERANT, a, out, b); }
int sk_loss_destroy(sk_loss_function* l) { SK_API_BEGIN delete l; SK_API_END }
int sk_loss_evaluate(const sk_loss_function* loss, double s, double rho[3]) {
  SK_API_BEGIN
  SK_REQUIRE(rho != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_loss_evaluate: null output");
  ensure_device();
  DBuf<double> d(3);
  const LossSpec spec = loss ? loss->spec : LossSpec{SK_LOSS_TRIVIAL, 0.0};
  k_loss_evaluate<<<1, 1>>>(spec, s, d.p);
  check_launch("k_loss_evaluate");
  SK_CUDA(cudaMemcpy(rho, d.p, 3 * sizeof(double), cudaMemcpyDeviceToHost));
  SK_API_END
}

// ---- CostFunction ------------------------------------------------------------------------------------
int sk_functor_info(int functor_id, int* num_residuals, int* num_parameter_blocks, int block_sizes[SK_MAX_PARAMETER_BLOCKS],
                    int* num_consts) {
  SK_API_BEGIN
  FunctorInfo fi;
  SK_REQUIRE(lookup_functor(functor_id, &fi), SK_ERR_UNSUPPORTED, kNoFunctor, functor_id);
  if (num_residuals) *num_residuals = fi.nres;
  if (num_parameter_blocks) *num_parameter_blocks = fi.nblk;
  if (num_consts) *num_consts = fi.nconsts;
  if (block_sizes) for (int i = 0; i < SK_MAX_PARAMETER_BLOCKS; ++i) block_sizes[i] = fi.sizes[i];
  SK_API_END
}
int sk_functor_register_source(const char* name, const char* cuda_source, int num_residuals, int num_parameter_blocks, const int* block_sizes,
                               int num_consts, int* out_functor_id) {
  SK_API_BEGIN
  SK_REQUIRE(out_functor_id != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_functor_register_source: null output");
  *out_functor_id = register_user_functor(name, cuda_source, num_residuals, num_parameter_blocks, block_sizes, num_consts);
  SK_API_END
}
int sk_cost_function_create(int functor_id, const double* consts, int num_consts, sk_cost_function** out) {
  SK_API_BEGIN
  SK_REQUIRE(out != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_cost_function_create: null output");
  FunctorInfo fi;
  SK_REQUIRE(lookup_functor(functor_id, &fi), SK_ERR_UNSUPPORTED, kNoFunctor, functor_id);
  SK_REQUIRE(num_consts == fi.nconsts && (consts != nullptr || num_consts == 0), SK_ERR_INVALID_ARGUMENT,
             "functor %d takes %d constants, got %d", functor_id, fi.nconsts, num_consts);
  std::unique_ptr<sk_cost_function> f(new sk_cost_function);
  f->functor_id = functor_id; f->info = fi;
  for (int i = 0; i < SK_MAX_CONSTS; ++i) f->consts[i] = i < num_consts ? consts[i] : 0.0;
  *out = f.release();
  SK_API_END
}
int sk_cost_function_destroy(sk_cost_function* f) { SK_API_BEGIN delete f; SK_API_END }
int sk_cost_function_num_residuals(const sk_cost_function* f) { return f ? f->info.nres : -1; }

int sk_cost_function_evaluate(const sk_cost_function* f, const sk_double_pointer* parameters, sk_double_pointer residuals,
                              const sk_double_pointer* jacobians, int* ok) {
  SK_API_BEGIN
  SK_REQUIRE(f != nullptr && parameters != nullptr && ok != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_cost_function_evaluate: null argument");
  ensure_device();
  EvalArgs a{};
  a.functor = f->functor_id; a.has_jac = jacobians != nullptr;
  for (int i = 0; i < SK_MAX_CONSTS; ++i) a.consts[i] = f->consts[i];
  for (int k = 0; k < f->info.nblk; ++k) {
    check_block(parameters[k].array, parameters[k].offset, f->info.sizes[k], "parameters");
    a.params[k] = parameters[k].array->d.p + parameters[k].offset;
    a.jac[k] = nullptr;
    if (jacobians != nullptr && jacobians[k].array != nullptr) {
      check_block(jacobians[k].array, jacobians[k].offset, f->info.nres * f->info.sizes[k], "jacobians");
      a.jac[k] = jacobians[k].array->d.p + jacobians[k].offset;
    }
  }
  check_block(residuals.array, residuals.offset, f->info.nres, "residuals");
  a.residuals = residuals.array->d.p + residuals.offset;
  DBuf<int> d_ok(1);
  if (f->functor_id >= kUserFunctorBase) launch_user_evaluate_single(f->functor_id, a, d_ok.p, nullptr);
  else k_evaluate_single<<<1, 1>>>(a, d_ok.p);
  check_launch("k_evaluate_single");
  SK_CUDA(cudaMemcpy(ok, d_ok.p, sizeof(int), cudaMemcpyDeviceToHost));
  SK_API_END
}

int sk_cost_function_evaluate_host(const sk_cost_function* f, double const* const* parameters, double* residuals,
                                   double** jacobians, int* ok) {
  SK_API_BEGIN
  SK_REQUIRE(f != nullptr && parameters != nullptr && residuals != nullptr && ok != nullptr, SK_ERR_INVALID_ARGUMENT,
             "sk_cost_function_evaluate_host: null argument");
  ensure_device();
  const FunctorInfo& fi = f->info;
  sk_double_array buf;
  buf.n = fi.ntot + fi.nres + fi.nres * fi.ntot;
  buf.d.alloc((size_t)buf.n);
  std::vector<double> h((size_t)buf.n, 0.0);
  std::vector<sk_double_pointer> pp(fi.nblk), jp(fi.nblk);
  int off = 0;
  for (int k = 0; k < fi.nblk; ++k) {
    SK_REQUIRE(parameters[k] != nullptr, SK_ERR_INVALID_ARGUMENT, "parameters[%d] is null", k);
    for (int c = 0; c < fi.sizes[k]; ++c) h[off + c] = parameters[k][c];
    pp[k] = {&buf, off};
    off += fi.sizes[k];
  }
  const int res_off = off; off += fi.nres;
  for (int k = 0; k < fi.nblk; ++k) {
    jp[k] = {(jacobians != nullptr && jacobians[k] != nullptr) ? &buf : nullptr, off};
    off += fi.nres * fi.sizes[k];
  }
  SK_CUDA(cudaMemcpy(buf.d.p, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
  const int st = sk_cost_function_evaluate(f, pp.data(), sk_double_pointer{&buf, res_off}, jacobians ? jp.data() : nullptr, ok);
  if (st != SK_OK) throw Error(st, g_last_error);
  SK_CUDA(cudaMemcpy(h.data(), buf.d.p, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
  if (*ok) {
    for (int q = 0; q < fi.nres; ++q) residuals[q] = h[res_off + q];
    if (jacobians != nullptr)
      for (int k = 0; k < fi.nblk; ++k)
        if (jacobians[k] != nullptr)
          for (int e = 0; e < fi.nres * fi.sizes[k]; ++e) jacobians[k][e] = h[jp[k].offset + e];
  }
  SK_API_END
}

// ---- Problem -------------------------------------------------------------------------------------------
int sk_problem_create(sk_problem** out) {
  SK_API_BEGIN
  SK_REQUIRE(out != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_problem_create: null output");
  *out = new sk_problem;
  SK_API_END
}
int sk_problem_destroy(sk_problem* p) { SK_API_BEGIN delete p; SK_API_END }

static ResidualGroup& group_for(sk_problem* p, int functor_id, const FunctorInfo& fi, LossSpec loss) {
  if (!p->groups.empty()) {
    ResidualGroup& g = p->groups.back();
    if (g.functor_id == functor_id && g.loss.type == loss.type && g.loss.a == loss.a && g.loss.b == loss.b) return g;
  }
  p->groups.emplace_back();
  ResidualGroup& g = p->groups.back();
  g.functor_id = functor_id; g.info = fi; g.loss = loss;
  return g;
}

int sk_problem_add_residual_block(sk_problem* p, const sk_cost_function* cost, const sk_loss_function* loss,
                                  const sk_double_pointer* blocks, int num_blocks, sk_residual_block_id* id) {
  SK_API_BEGIN
  SK_REQUIRE(p != nullptr && cost != nullptr && blocks != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_problem_add_residual_block: null argument");
  const FunctorInfo& fi = cost->info;
  SK_REQUIRE(num_blocks == fi.nblk, SK_ERR_INVALID_ARGUMENT, "cost function expects %d parameter blocks, got %d", fi.nblk, num_blocks);
  for (int k = 0; k < fi.nblk; ++k) {
    check_block(blocks[k].array, blocks[k].offset, fi.sizes[k], "addResidualBlock");
    for (int j = 0; j < k; ++j)
      SK_REQUIRE(!(blocks[j].array == blocks[k].array && blocks[j].offset == blocks[k].offset), SK_ERR_INVALID_ARGUMENT,
                 "duplicate parameter block in one residual block");
  }
  const LossSpec ls = loss ? loss->spec : LossSpec{SK_LOSS_TRIVIAL, 0.0};
  ResidualGroup& g = group_for(p, cost->functor_id, fi, ls);
  for (int k = 0; k < fi.nblk; ++k) { g.arrays.push_back(blocks[k].array); g.offsets.push_back(blocks[k].offset); }
  for (int c = 0; c < fi.nconsts; ++c) g.consts.push_back(cost->consts[c]);
  g.n += 1;
  if (id) *id = p->num_residual_blocks;
  p->num_residual_blocks += 1; p->num_residuals += fi.nres;
  SK_API_END
}

int sk_problem_add_residual_blocks(sk_problem* p, int functor_id, int64_t n, const double* consts, const sk_loss_function* loss,
                                   sk_double_array* array, const int64_t* block_offsets, sk_residual_block_id* first_id) {
  SK_API_BEGIN
  SK_REQUIRE(p != nullptr && array != nullptr && block_offsets != nullptr && n >= 0, SK_ERR_INVALID_ARGUMENT,
             "sk_problem_add_residual_blocks: null argument");
  FunctorInfo fi;
  SK_REQUIRE(lookup_functor(functor_id, &fi), SK_ERR_UNSUPPORTED, kNoFunctor, functor_id);
  SK_REQUIRE(consts != nullptr || fi.nconsts == 0 || n == 0, SK_ERR_INVALID_ARGUMENT, "missing constants");
  {
    int64_t bad = -1;                                          // lowest offending residual block, whatever the thread count
    std::mutex mu;
    parallel_for(n, [&](int64_t a, int64_t b) {
      for (int64_t i = a; i < b; ++i)
        for (int k = 0; k < fi.nblk; ++k) {
          const int64_t off = block_offsets[i * fi.nblk + k];
          if (off < 0 || off + fi.sizes[k] > array->n) { std::lock_guard<std::mutex> g(mu); if (bad < 0 || i < bad) bad = i; return; }
        }
    });
    if (bad >= 0)
      for (int k = 0; k < fi.nblk; ++k) {
        const int64_t off = block_offsets[bad * fi.nblk + k];
        if (off < 0 || off + fi.sizes[k] > array->n)
          throw Error(SK_ERR_INVALID_ARGUMENT, fmt("residual block %lld: parameter block %d at offset %lld is outside the array of %lld doubles",
                                                   (long long)bad, k, (long long)off, (long long)array->n));
      }
  }
  const LossSpec ls = loss ? loss->spec : LossSpec{SK_LOSS_TRIVIAL, 0.0};
  p->groups.emplace_back();
  ResidualGroup& g = p->groups.back();
  g.functor_id = functor_id; g.info = fi; g.loss = ls; g.n = n;
  g.arrays.assign(1, array);            // size-1 == "all blocks in this one array"
  // the problem owns copies (the caller may reuse its arrays); copied by all host threads -- at 5M residual blocks the
  // two single-threaded copies were 40 ms of the end-to-end path
  g.offsets.resize((size_t)n * fi.nblk);
  if (fi.nconsts) g.consts.resize((size_t)n * fi.nconsts);
  parallel_for(n, [&](int64_t a, int64_t b) {
    std::memcpy(g.offsets.data() + a * fi.nblk, block_offsets + a * fi.nblk, sizeof(int64_t) * (size_t)(b - a) * fi.nblk);
    if (fi.nconsts) std::memcpy(g.consts.data() + a * fi.nconsts, consts + a * fi.nconsts, sizeof(double) * (size_t)(b - a) * fi.nconsts);
  });
  if (first_id) *first_id = p->num_residual_blocks;
  p->num_residual_blocks += n; p->num_residuals += n * fi.nres;
  SK_API_END
}

int sk_problem_add_parameter_blocks(sk_problem* p, sk_double_array* array, int64_t n, const int64_t* offsets, int32_t block_size) {
  SK_API_BEGIN
  SK_REQUIRE(p != nullptr && array != nullptr && (offsets != nullptr || n == 0) && n >= 0 && block_size > 0, SK_ERR_INVALID_ARGUMENT,
             "sk_problem_add_parameter_blocks: invalid argument");
  for (int64_t i = 0; i < n; ++i)
    SK_REQUIRE(offsets[i] >= 0 && offsets[i] + block_size <= array->n, SK_ERR_INVALID_ARGUMENT,
               "parameter block %lld at offset %lld (size %d) is outside the array of %lld doubles", (long long)i, (long long)offsets[i],
               block_size, (long long)array->n);
  p->declared.push_back({array, block_size, std::vector<int64_t>(offsets, offsets + n)});
  SK_API_END
}

int64_t sk_problem_num_residual_blocks(const sk_problem* p) { return p ? p->num_residual_blocks : -1; }
int64_t sk_problem_num_residuals(const sk_problem* p) { return p ? p->num_residuals : -1; }

static sk_double_array* group_array(const ResidualGroup& g, int64_t i, int k) {
  return g.arrays.size() == 1 ? g.arrays[0] : g.arrays[(size_t)i * g.info.nblk + k];
}

static void count_blocks(const sk_problem* p, int64_t* nblocks, int64_t* nparams) {
  std::map<std::pair<const sk_double_array*, int64_t>, int> seen;
  for (auto& g : p->groups)
    for (int64_t i = 0; i < g.n; ++i)
      for (int k = 0; k < g.info.nblk; ++k) seen[{group_array(g, i, k), g.offsets[(size_t)i * g.info.nblk + k]}] = g.info.sizes[k];
  for (auto& d : p->declared)
    for (int64_t off : d.offsets) seen[{d.array, off}] = d.size;
  *nblocks = (int64_t)seen.size(); *nparams = 0;
  for (auto& kv : seen) *nparams += kv.second;
}
int64_t sk_problem_num_parameter_blocks(const sk_problem* p) { if (!p) return -1; int64_t a, b; count_blocks(p, &a, &b); return a; }
int64_t sk_problem_num_parameters(const sk_problem* p) { if (!p) return -1; int64_t a, b; count_blocks(p, &a, &b); return b; }

// ---- Options / Summary ---------------------------------------------------------------------------------
void sk_solver_options_init(sk_solver_options* o) {
  if (!o) return;
  memset(o, 0, sizeof *o);
  o->minimizer_type = SK_TRUST_REGION; o->trust_region_strategy_type = SK_LEVENBERG_MARQUARDT;
  o->linear_solver_type = SK_SPARSE_NORMAL_CHOLESKY; o->preconditioner_type = SK_JACOBI;
  o->max_num_iterations = 50; o->max_num_consecutive_invalid_steps = 5;
  o->min_linear_solver_iterations = 0; o->max_linear_solver_iterations = 500;
  o->jacobi_scaling = 1; o->minimizer_progress_to_stdout = 0; o->num_threads = 1; o->profile_kernels = 0;
  o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
  o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8; o->eta = 1e-1;
  o->max_solver_time_in_seconds = 1e9; o->comm = nullptr; o->residual_blocks_are_local = 0;
}
int sk_solver_summary_create(sk_solver_summary** out) {
  SK_API_BEGIN
  SK_REQUIRE(out != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_solver_summary_create: null output");
  *out = new sk_solver_summary;
  (*out)->data.termination_type = SK_FAILURE;
  (*out)->message = "ceres::Solve was not called.";
  SK_API_END
}
int sk_solver_summary_destroy(sk_solver_summary* s) { SK_API_BEGIN delete s; SK_API_END }
int sk_solver_summary_get(const sk_solver_summary* s, sk_solver_summary_data* out) {
  SK_API_BEGIN
  SK_REQUIRE(s != nullptr && out != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_solver_summary_get: null argument");
  *out = s->data;
  SK_API_END
}
int sk_solver_summary_iterations(const sk_solver_summary* s, sk_iteration_summary* out, int32_t capacity, int32_t* count) {
  SK_API_BEGIN
  SK_REQUIRE(s != nullptr && count != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_solver_summary_iterations: null argument");
  *count = (int32_t)s->rows.size();
  for (int32_t i = 0; i < capacity && i < *count && out != nullptr; ++i) out[i] = s->rows[i];
  SK_API_END
}
const char* sk_solver_summary_message(const sk_solver_summary* s) { return s ? s->message.c_str() : ""; }
const char* sk_solver_summary_brief_report(sk_solver_summary* s) { if (!s) return ""; format_reports(s); return s->brief.c_str(); }
const char* sk_solver_summary_full_report(sk_solver_summary* s) { if (!s) return ""; format_reports(s); return s->full.c_str(); }
int sk_solver_summary_is_solution_usable(const sk_solver_summary* s) {
  if (!s) return 0;
  const int t = s->data.termination_type;
  return t == SK_CONVERGENCE || t == SK_NO_CONVERGENCE || t == SK_USER_SUCCESS;
}

// ---- Solve -----------------------------------------------------------------------------------------------
static bool is_schur(int t) { return t == SK_DENSE_SCHUR || t == SK_SPARSE_SCHUR || t == SK_ITERATIVE_SCHUR; }

static std::unique_ptr<LmSolver> prepare_ba(const sk_solver_options& opt, sk_problem* p, cudaStream_t stream) {
  // All residual blocks must be SnavelyReprojectionError(2; 9, 3) -- or ONE functor of that shape registered from source
  // (sk_functor_register_source: it is compiled into the same tile kernel) -- over one DoubleArray with one loss:
  // the SchurEliminator<2, 3, 9> shape (cameras = f-blocks, points = e-blocks).
  sk_double_array* array = nullptr;
  LossSpec loss{SK_LOSS_TRIVIAL, 0.0};
  int64_t n = 0;
  bool first = true;
  int functor = SK_FUNCTOR_SNAVELY_REPROJECTION_ERROR;
  for (auto& g : p->groups) {
    if (g.n == 0) continue;
    SK_REQUIRE(g.functor_id == SK_FUNCTOR_SNAVELY_REPROJECTION_ERROR || (g.functor_id >= kUserFunctorBase && user_functor_runs_on_tiles(g.functor_id)),
               SK_ERR_UNSUPPORTED,
               "Schur-type linear solvers are implemented for bundle adjustment only: SnavelyReprojectionError residual blocks, or a functor "
               "of the same shape (2 residuals; blocks of 9 and 3; 2 constants) registered from source; functor %d found", g.functor_id);
    if (first) functor = g.functor_id;
    SK_REQUIRE(g.functor_id == functor, SK_ERR_UNSUPPORTED, "one cost functor per bundle adjustment problem (functors %d and %d found)", functor, g.functor_id);
    for (size_t a = 0; a < g.arrays.size(); ++a) {
      if (array == nullptr) array = g.arrays[a];
      SK_REQUIRE(g.arrays[a] == array, SK_ERR_UNSUPPORTED, "bundle adjustment parameter blocks must live in one DoubleArray");
    }
    if (first) { loss = g.loss; first = false; }
    SK_REQUIRE(g.loss.type == loss.type && g.loss.a == loss.a && g.loss.b == loss.b, SK_ERR_UNSUPPORTED, "one loss function per bundle adjustment problem");
    n += g.n;
  }
  SK_REQUIRE(n > 0, SK_ERR_INVALID_ARGUMENT, "problem has no residual blocks");
  const bool trace = getenv("SKERES_TRACE_HOST") != nullptr;   // development: where the preprocessor's time goes
  const double tp0 = wall();
  BaLayoutHost H;
  // Rank-local mode: the problem holds this rank's share only, so the layout is built as on one GPU; the camera table is
  // made the same on every rank by the declared (AddParameterBlock) camera blocks.
  const bool local = opt.residual_blocks_are_local != 0;
  const int rank = (opt.comm && !local) ? opt.comm->rank : 0, world = (opt.comm && !local) ? opt.comm->world : 1;
  std::vector<int64_t> declared_cams;
  if (local)
    for (auto& d : p->declared)
      if (d.size == 9 && d.array == array) declared_cams.insert(declared_cams.end(), d.offsets.begin(), d.offsets.end());
  const std::vector<int64_t>* extra = declared_cams.empty() ? nullptr : &declared_cams;
  const ResidualGroup* single = nullptr;                    // one bulk group (the usual case): read its arrays in place
  for (auto& g : p->groups) if (g.n == n) single = &g;
  double tp1 = tp0;
  // The usual case -- one bulk group, one GPU or rank-local blocks, implicit Schur solver -- is built by kernels from the uploaded
  // residual-block table (ba_layout_device.cu); anything that path does not cover falls through to the host builder.
  BaLayoutDevice dev;
  const bool try_device = single != nullptr && world == 1 && opt.linear_solver_type == SK_ITERATIVE_SCHUR;
  if (try_device && build_ba_layout_device(n, single->offsets.data(), single->consts.data(), extra, stream, &H, &dev)) {
  } else if (single != nullptr) {
    H = BaLayoutHost{};
    build_ba_layout(n, single->offsets.data(), single->offsets.data() + 1, single->consts.data(), rank, world, &H, 2, extra);
  } else {
    std::vector<int64_t> cam_off((size_t)n), pt_off((size_t)n);
    std::vector<double> obs((size_t)2 * n);
    int64_t at = 0;
    for (auto& g : p->groups)
      for (int64_t i = 0; i < g.n; ++i, ++at) {
        cam_off[at] = g.offsets[2 * i]; pt_off[at] = g.offsets[2 * i + 1];
        obs[2 * at] = g.consts[2 * i]; obs[2 * at + 1] = g.consts[2 * i + 1];
      }
    tp1 = wall();
    build_ba_layout(n, cam_off.data(), pt_off.data(), obs.data(), rank, world, &H, 1, extra);
  }
  const double tp2 = wall();
  std::vector<int64_t> all_pt;
  int64_t total_points = H.n_pts;
  if (world > 1) {                 // computed once by the layout builder (was: a second O(n log n) sort of all observations)
    all_pt = std::move(H.all_pt_offset);
    total_points = (int64_t)all_pt.size();
  }
  const int64_t n_cams = H.n_cams;
  std::unique_ptr<BaSolver> solver(new BaSolver(opt, stream, std::move(H), array->d.p, array->n, loss, &dev, functor));
  solver->fill_totals(n, n_cams + total_points, 9 * n_cams + 3 * total_points, std::move(all_pt));
  if (local && opt.comm != nullptr && opt.comm->world > 1) solver->exchange_local_totals();
  if (trace) fprintf(stderr, "[skeres] preprocess: flatten %.3f s, layout %.3f s, device set-up %.3f s\n", tp1 - tp0, tp2 - tp1, wall() - tp2);
  return solver;
}

static std::unique_ptr<LmSolver> prepare_dense(const sk_solver_options& opt, sk_problem* p, cudaStream_t stream) {
  std::map<std::pair<sk_double_array*, int64_t>, std::pair<int, int>> blocks;   // -> (first column, size), program order
  std::vector<std::pair<sk_double_array*, int64_t>> order;
  std::vector<DenseRb> rbs;
  int row = 0, ncols = 0;
  for (auto& g : p->groups)
    for (int64_t i = 0; i < g.n; ++i) {
      DenseRb rb{};
      rb.functor = g.functor_id; rb.row = row; rb.loss_type = g.loss.type; rb.loss_a = g.loss.a; rb.loss_b = g.loss.b;
      for (int c = 0; c < g.info.nconsts; ++c) rb.consts[c] = g.consts[(size_t)i * g.info.nconsts + c];
      for (int k = 0; k < g.info.nblk; ++k) {
        auto key = std::make_pair(group_array(g, i, k), g.offsets[(size_t)i * g.info.nblk + k]);
        auto it = blocks.find(key);
        if (it == blocks.end()) {                       // Ceres Program order: first appearance
          it = blocks.emplace(key, std::make_pair(ncols, g.info.sizes[k])).first;
          order.push_back(key);
          ncols += g.info.sizes[k];
        } else {
          SK_REQUIRE(it->second.second == g.info.sizes[k], SK_ERR_INVALID_ARGUMENT, "parameter block used with two different sizes");
        }
        rb.col[k] = it->second.first;
      }
      rbs.push_back(rb);
      row += g.info.nres;
    }
  SK_REQUIRE(!rbs.empty(), SK_ERR_INVALID_ARGUMENT, "problem has no residual blocks");
  // overlapping blocks check
  {
    std::vector<std::tuple<sk_double_array*, int64_t, int>> all;
    for (auto& kv : blocks) all.emplace_back(kv.first.first, kv.first.second, kv.second.second);
    std::sort(all.begin(), all.end());
    for (size_t k = 1; k < all.size(); ++k)
      SK_REQUIRE(std::get<0>(all[k]) != std::get<0>(all[k - 1]) || std::get<1>(all[k]) >= std::get<1>(all[k - 1]) + std::get<2>(all[k - 1]),
                 SK_ERR_INVALID_ARGUMENT, "overlapping parameter blocks");
  }
  std::vector<double*> ptrs((size_t)ncols);
  for (auto& key : order) {
    const auto& cs = blocks[key];
    for (int c = 0; c < cs.second; ++c) ptrs[cs.first + c] = key.first->d.p + key.second + c;
  }
  return std::unique_ptr<LmSolver>(new DenseSolver(opt, stream, rbs, row, ptrs, (int)order.size()));
}

}  // extern "C" (reopened below)

struct sk_solver {
  cudaStream_t stream = nullptr;
  std::unique_ptr<LmSolver> impl;
  double preprocessor_s = 0.0;
  int device = 0;
  ~sk_solver() { impl.reset(); if (stream) cudaStreamDestroy(stream); }
};

extern "C" {

int sk_solver_create(const sk_solver_options* options, sk_problem* problem, sk_solver** out) {
  SK_API_BEGIN
  SK_REQUIRE(options != nullptr && problem != nullptr && out != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_solver_create: null argument");
  const double t0 = wall();
  const sk_solver_options& o = *options;
  SK_REQUIRE(o.minimizer_type == SK_TRUST_REGION, SK_ERR_UNSUPPORTED, "only the TRUST_REGION minimizer has a device implementation");
  SK_REQUIRE(o.trust_region_strategy_type == SK_LEVENBERG_MARQUARDT, SK_ERR_UNSUPPORTED, "only the LEVENBERG_MARQUARDT strategy has a device implementation");
  SK_REQUIRE(o.max_num_iterations >= 0 && o.initial_trust_region_radius > 0, SK_ERR_INVALID_ARGUMENT, "invalid solver options");
  ensure_device();
  std::unique_ptr<sk_solver> s(new sk_solver);
  s->device = g_device;
  SK_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  if (is_schur(o.linear_solver_type)) s->impl = prepare_ba(o, problem, s->stream);
  // DENSE_NORMAL_CHOLESKY and SPARSE_NORMAL_CHOLESKY -- the latter is Ceres' DEFAULT linear_solver_type, which HelloWorld.scala and
  // HelloWorldNumericDiff.scala never change -- solve the same damped normal equations as DENSE_QR: they run on the dense QR back
  // end (same step up to rounding; within that back end's size limits).
  else if (o.linear_solver_type == SK_DENSE_QR || o.linear_solver_type == SK_DENSE_NORMAL_CHOLESKY || o.linear_solver_type == SK_SPARSE_NORMAL_CHOLESKY)
    s->impl = prepare_dense(o, problem, s->stream);
  else throw Error(SK_ERR_UNSUPPORTED, fmt("linear_solver_type %d has no device implementation (supported: DENSE_QR, DENSE_NORMAL_CHOLESKY, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR)", o.linear_solver_type));
  s->preprocessor_s = wall() - t0;
  *out = s.release();
  SK_API_END
}

int sk_solver_minimize(sk_solver* solver, int32_t max_num_iterations_override, sk_solver_summary* summary) {
  SK_API_BEGIN
  SK_REQUIRE(solver != nullptr && summary != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_solver_minimize: null argument");
  const double t0 = wall();
  SK_CUDA(cudaSetDevice(solver->device));
  *summary = sk_solver_summary{};
  summary->data.termination_type = SK_FAILURE;
  summary->data.preprocessor_time_in_seconds = solver->preprocessor_s;
  solver->impl->minimize(summary, max_num_iterations_override);
  summary->data.total_time_in_seconds = wall() - t0 + solver->preprocessor_s;
  format_reports(summary);
  SK_API_END
}

int sk_solver_destroy(sk_solver* solver) {
  SK_API_BEGIN
  if (solver != nullptr && getenv("SKERES_TRACE_HOST")) {
    const double t0 = wall();
    solver->impl.reset();
    const double t1 = wall();
    if (solver->stream) { cudaStreamDestroy(solver->stream); solver->stream = nullptr; }
    const double t2 = wall();
    fprintf(stderr, "[skeres] sk_solver_destroy: solver %.3f s (pool: sync %.3f s, free %.3f s over %ld blocks), stream %.3f s\n", t1 - t0,
            DevicePool::get().t_sync, DevicePool::get().t_free, DevicePool::get().n_give, t2 - t1);
    DevicePool::get().t_sync = DevicePool::get().t_free = 0.0; DevicePool::get().n_give = 0;
  }
  delete solver;
  SK_API_END
}

int sk_solver_time_schur_product(sk_solver* solver, int32_t reps, double* out_ms_per_launch) {
  SK_API_BEGIN
  SK_REQUIRE(solver != nullptr && out_ms_per_launch != nullptr && reps > 0, SK_ERR_INVALID_ARGUMENT, "sk_solver_time_schur_product: bad argument");
  SK_CUDA(cudaSetDevice(solver->device));
  *out_ms_per_launch = solver->impl->time_linear_operator(reps);
  SK_API_END
}

int sk_solve(const sk_solver_options* options, sk_problem* problem, sk_solver_summary* summary) {
  if (summary == nullptr) return fail(SK_ERR_INVALID_ARGUMENT, "sk_solve: null argument");
  sk_solver* s = nullptr;
  const bool trace = getenv("SKERES_TRACE_HOST") != nullptr;
  const double t0 = wall();
  int st = sk_solver_create(options, problem, &s);
  if (st != SK_OK) { summary->data.termination_type = SK_FAILURE; summary->message = g_last_error; return st; }
  const double t1 = wall();
  st = sk_solver_minimize(s, -1, summary);
  const double t2 = wall();
  sk_solver_destroy(s);
  if (trace) fprintf(stderr, "[skeres] sk_solve: create %.3f s, minimize %.3f s, destroy %.3f s\n", t1 - t0, t2 - t1, wall() - t2);
  return st;
}

// ---- BAL reader ---------------------------------------------------------------------------------------------
int sk_bal_problem_from_file(const char* path, sk_bal_problem** out) {
  SK_API_BEGIN
  SK_REQUIRE(path != nullptr && out != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_bal_problem_from_file: null argument");
  FILE* f = std::fopen(path, "r");
  SK_REQUIRE(f != nullptr, SK_ERR_IO, "cannot open BAL file %s", path);
  std::unique_ptr<sk_bal_problem> b(new sk_bal_problem);
  struct Closer { FILE* f; ~Closer() { std::fclose(f); } } closer{f};
  SK_REQUIRE(std::fscanf(f, "%d %d %d", &b->n_cam, &b->n_pt, &b->n_obs) == 3, SK_ERR_IO, "%s: bad BAL header", path);   // :41-43
  SK_REQUIRE(b->n_cam > 0 && b->n_pt > 0 && b->n_obs > 0, SK_ERR_IO, "%s: bad BAL header", path);
  b->cam_idx.resize(b->n_obs); b->pt_idx.resize(b->n_obs); b->obs.resize((size_t)2 * b->n_obs);
  for (int i = 0; i < b->n_obs; ++i) {                                                                                  // :52-58
    SK_REQUIRE(std::fscanf(f, "%d %d %lf %lf", &b->cam_idx[i], &b->pt_idx[i], &b->obs[2 * i], &b->obs[2 * i + 1]) == 4, SK_ERR_IO,
               "%s: bad observation line %d", path, i);
    SK_REQUIRE(b->cam_idx[i] >= 0 && b->cam_idx[i] < b->n_cam && b->pt_idx[i] >= 0 && b->pt_idx[i] < b->n_pt, SK_ERR_IO,
               "%s: observation %d references camera %d / point %d out of range", path, i, b->cam_idx[i], b->pt_idx[i]);
  }
  const int64_t np = (int64_t)9 * b->n_cam + (int64_t)3 * b->n_pt;                                                      // :49
  std::vector<double> params((size_t)np);
  for (int64_t i = 0; i < np; ++i) SK_REQUIRE(std::fscanf(f, "%lf", &params[i]) == 1, SK_ERR_IO, "%s: bad parameter %lld", path, (long long)i);  // :60-62
  int st = sk_double_array_create(np, &b->params);
  if (st != SK_OK) throw Error(st, g_last_error);
  st = sk_double_array_upload(b->params, 0, params.data(), np);
  if (st != SK_OK) { sk_double_array_destroy(b->params); throw Error(st, g_last_error); }
  *out = b.release();
  SK_API_END
}
int sk_bal_problem_destroy(sk_bal_problem* b) { SK_API_BEGIN if (b) { delete b->params; delete b; } SK_API_END }
int32_t sk_bal_problem_num_cameras(const sk_bal_problem* b) { return b ? b->n_cam : -1; }
int32_t sk_bal_problem_num_points(const sk_bal_problem* b) { return b ? b->n_pt : -1; }
int32_t sk_bal_problem_num_observations(const sk_bal_problem* b) { return b ? b->n_obs : -1; }
sk_double_array* sk_bal_problem_parameters(sk_bal_problem* b) { return b ? b->params : nullptr; }
const int32_t* sk_bal_problem_camera_index(const sk_bal_problem* b) { return b ? b->cam_idx.data() : nullptr; }
const int32_t* sk_bal_problem_point_index(const sk_bal_problem* b) { return b ? b->pt_idx.data() : nullptr; }
const double* sk_bal_problem_observations(const sk_bal_problem* b) { return b ? b->obs.data() : nullptr; }
int sk_bal_problem_build(sk_bal_problem* b, const sk_loss_function* loss, sk_problem* problem) {
  SK_API_BEGIN
  SK_REQUIRE(b != nullptr && problem != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_bal_problem_build: null argument");
  std::vector<int64_t> off((size_t)2 * b->n_obs);
  for (int i = 0; i < b->n_obs; ++i) {                     // mutableCameraForObservation / mutablePointForObservation (:31-33)
    off[2 * (size_t)i] = (int64_t)9 * b->cam_idx[i];
    off[2 * (size_t)i + 1] = (int64_t)9 * b->n_cam + (int64_t)3 * b->pt_idx[i];
  }
  const int st = sk_problem_add_residual_blocks(problem, SK_FUNCTOR_SNAVELY_REPROJECTION_ERROR, b->n_obs, b->obs.data(), loss, b->params,
                                                off.data(), nullptr);
  if (st != SK_OK) throw Error(st, g_last_error);
  SK_API_END
}

// ---- batched curve fits ------------------------------------------------------------------------------------
int sk_bal_block_offsets(int64_t n_obs, const int32_t* camera_index, const int32_t* point_index, int32_t num_cameras, int32_t num_points,
                         int64_t* out_offsets) {
  SK_API_BEGIN
  SK_REQUIRE(n_obs >= 0 && (n_obs == 0 || (camera_index && point_index && out_offsets)) && num_cameras >= 0 && num_points >= 0,
             SK_ERR_INVALID_ARGUMENT, "sk_bal_block_offsets: invalid argument");
  int64_t bad = -1;
  std::mutex mu;
  parallel_for(n_obs, [&](int64_t a, int64_t b) {
    const int64_t pbase = (int64_t)9 * num_cameras;
    for (int64_t i = a; i < b; ++i) {
      const int32_t c = camera_index[i], p = point_index[i];
      if (c < 0 || c >= num_cameras || p < 0 || p >= num_points) { std::lock_guard<std::mutex> g(mu); if (bad < 0 || i < bad) bad = i; return; }
      out_offsets[2 * i] = (int64_t)9 * c; out_offsets[2 * i + 1] = pbase + (int64_t)3 * p;
    }
  });
  SK_REQUIRE(bad < 0, SK_ERR_INVALID_ARGUMENT, "observation %lld: camera index %d / point index %d outside [0, %d) / [0, %d)", (long long)bad,
             camera_index[bad], point_index[bad], num_cameras, num_points);
  SK_API_END
}

int sk_curve_fit_batch_solve(const sk_solver_options* options, int64_t n_problems, int32_t n_obs, const sk_double_array* x,
                             const sk_double_array* y, sk_double_array* mc, double* out_initial_cost, double* out_final_cost,
                             int32_t* out_num_iterations, int32_t* out_termination_type, sk_solver_summary* summary) {
  SK_API_BEGIN
  SK_REQUIRE(options != nullptr && x != nullptr && y != nullptr && mc != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_curve_fit_batch_solve: null argument");
  SK_REQUIRE(n_problems > 0 && n_obs > 0, SK_ERR_INVALID_ARGUMENT, "sk_curve_fit_batch_solve: empty batch");
  SK_REQUIRE(x->n >= n_problems * n_obs && y->n >= n_problems * n_obs && mc->n >= 2 * n_problems, SK_ERR_INVALID_ARGUMENT,
             "sk_curve_fit_batch_solve: arrays too small for %lld problems of %d observations", (long long)n_problems, n_obs);
  SK_REQUIRE(options->minimizer_type == SK_TRUST_REGION && options->trust_region_strategy_type == SK_LEVENBERG_MARQUARDT &&
             options->linear_solver_type == SK_DENSE_QR, SK_ERR_UNSUPPORTED, "batched curve fits run TRUST_REGION / LEVENBERG_MARQUARDT / DENSE_QR");
  ensure_device();
  sk_solver_summary local;
  curve_fit_batch_solve(*options, n_problems, n_obs, x->d.p, y->d.p, mc->d.p, out_initial_cost, out_final_cost, out_num_iterations,
                        out_termination_type, summary ? summary : &local);
  SK_API_END
}

// ---- communicator --------------------------------------------------------------------------------------------
int sk_comm_get_unique_id(char id[SK_COMM_UNIQUE_ID_BYTES]) { SK_API_BEGIN SK_REQUIRE(id != nullptr, SK_ERR_INVALID_ARGUMENT, "null id"); comm_get_unique_id(id); SK_API_END }
int sk_comm_create(const char id[SK_COMM_UNIQUE_ID_BYTES], int rank, int world_size, sk_comm** out) {
  SK_API_BEGIN
  SK_REQUIRE(id != nullptr && out != nullptr, SK_ERR_INVALID_ARGUMENT, "sk_comm_create: null argument");
  ensure_device();
  *out = comm_create(id, rank, world_size);
  SK_API_END
}
int sk_comm_destroy(sk_comm* c) { SK_API_BEGIN comm_destroy(c); SK_API_END }
int sk_comm_rank(const sk_comm* c) { return c ? c->rank : 0; }
int sk_comm_world_size(const sk_comm* c) { return c ? c->world : 1; }
int sk_partition_points(int64_t n_points, const int64_t* point_ptr, int world_size, int64_t* out_begin) {
  SK_API_BEGIN
  SK_REQUIRE(n_points >= 0 && point_ptr != nullptr && out_begin != nullptr && world_size >= 1, SK_ERR_INVALID_ARGUMENT, "sk_partition_points: bad arguments");
  partition_points(n_points, point_ptr, world_size, out_begin);
  SK_API_END
}
int sk_release_cached_memory(void) { SK_API_BEGIN DevicePool::get().release_all(); SK_API_END }
int64_t sk_cached_memory_bytes(void) { return (int64_t)DevicePool::get().cached_bytes(); }

}  // extern "C"
