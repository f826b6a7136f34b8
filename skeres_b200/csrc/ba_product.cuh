// ba_product.cuh — device side of the implicit-Schur product (ImplicitSchurComplement::RightMultiply), shared
// by the stand-alone product kernel (ba_kernels.cu: k_ba_matvec_tma) and the fused PCG solve (pcg_fused.cu: k_pcg_solve).
// Everything here is per-translation-unit (anonymous namespace): the two kernels run the same instructions in the same order
// and therefore produce the same bits.
#pragma once
#include "ba_kernels.cuh"
#include "ba_tile.cuh"

namespace sk {
namespace {

// ---- two-level sums of the implicit-Schur product --------------------------------------------------------------------
// Round 1 formed every per-point sum and every per-(segment, component) sum as ONE serial chain of dependent shared-memory
// loads (up to 16 resp. 40 long) walked by ~50 resp. ~150 of the 256 threads while the others waited at the next barrier:
// 45 % of the warp samples of the kernel sat behind those two barriers (profiles/r01_v9_matvec_tma_ncu_source_lines.txt).
// Here every sum is cut into fixed-size chunks (4 observations of a point, 8 of a segment), all threads of the CTA add one
// chunk each as a small tree from independent loads, and a second short pass adds the chunk sums of a point / segment in
// order.  The order of additions is fixed by the chunk tables alone, so the prefetching and the classic kernel still agree
// bit for bit.  w is staged [observation][3], v in the segment order [position][9] (odd strides: conflict-free writes), so
// that a chunk is one contiguous run and no permutation is read on the way.  Measured (profiles/r02_v1_summary.md): neutral
// against the serial chains (two more passes over shared memory, one more barrier), so the serial chains stay the default.
constexpr int VS = kSegRow;                // row stride of the segment-ordered staging of v
__device__ __forceinline__ RecView rec_view(const BaDev& L, const unsigned char* base) { return ::sk::rec_view(base, L.rec_sp, L.rec_pp, L.rec_sc); }

// pw[3 c + k] = sum over chunk c of w[.][k]; then u = (E^T E)^-1 (sum of the point's chunk sums).  Two barriers inside.
template <int NT = T>
__device__ __forceinline__ void point_sums_chunked(const RecView& R, int np, const double* w, double* pw, const double* ei,
                                                   double* u, int UP) {
  const int tid = threadIdx.x;
  const int n3 = 3 * (int)R.pcptr[np];
  for (int idx = tid; idx < n3; idx += NT) pw[idx] = point_chunk_sum(R, w, idx);
  __syncthreads();
  if (tid < np) {
    double a0, a1, a2;
    point_combine(R, pw, tid, a0, a1, a2);
    const double* m = ei + tid * 6;
    u[tid] = m[0] * a0 + m[1] * a1 + m[2] * a2;
    u[UP + tid] = m[1] * a0 + m[3] * a1 + m[4] * a2;
    u[2 * UP + tid] = m[2] * a0 + m[4] * a1 + m[5] * a2;
  }
  __syncthreads();
}

// ps[9 c + k] = sum over chunk c of the segment-ordered v[.][k]; then seg_y[spos[s]][k] = sum of the segment's chunk sums.
template <int NT = T>
__device__ __forceinline__ void seg_sums_chunked(const RecView& R, int ns, int sb, const double* vs, double* ps, double* seg_y) {
  const int tid = threadIdx.x;
  const int n9 = 9 * (int)R.scptr[ns];
  for (int idx = tid; idx < n9; idx += NT) ps[idx] = seg_chunk_sum(R, vs, idx);
  __syncthreads();
  for (int idx = tid; idx < ns * 9; idx += NT) { const int s = idx / 9; seg_y[(size_t)R.spos[s] * 9 + (idx - 9 * s)] = seg_combine(R, ps, idx); }
}

// Arithmetic of one observation inside the implicit-Schur product, with every rounding spelled out: the product kernels must
// agree bit for bit, so nothing is left to the compiler's choice of multiply-add contraction.  A row's dot products are FMA
// chains; what combines the two rows of an observation is one rounded product per row and one rounded sum (row 0 first).
__device__ __forceinline__ double two_rows(double a0, double b0, double a1, double b1) {   // a0 b0 + a1 b1, three roundings
  return __dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1));
}

__device__ __forceinline__ void l2_prefetch(const void* gsrc, unsigned bytes) {   // TMA prefetch into L2: no registers, no smem
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gsrc), "r"(bytes) : "memory");
}

// ---- TMA / mbarrier primitives ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// One box of a 2-D tensor map (inner coordinate c0, outer c1) into shared memory, completion on an mbarrier.
__device__ __forceinline__ void tma_box_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* b) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
               "r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
               "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
// The same copies with an L2 cache policy (createpolicy): evict_last keeps the lines resident across the products of a linear
// solve, evict_first marks a pure stream.
__device__ __forceinline__ unsigned long long l2_policy(bool keep) {
  unsigned long long pol;
  if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_box_2d_hint(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* b, unsigned long long pol) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;\n" ::
               "r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(smem_u32(b)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, unsigned bytes, unsigned long long* b, unsigned long long pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::
               "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)), "l"(pol) : "memory");
}

// ---- the persistent, fully prefetching product ------------------------------------------------------------------------
// Each CTA walks a strided list of tiles (tile blockIdx.x + it * gridDim.x) and, while it computes tile i from registers, the TMA
// engine streams EVERYTHING tile i+1 needs into shared memory: the 12 Jacobian planes (TMAP: two [12 planes][128 observations]
// boxes of a 2-D tensor map; else 12 bulk copies of <= 4 KB), the tile's metadata record (one bulk copy; packed per tile by
// build_tile_records, layout: RecView) and the (E^T E)^-1 blocks of its points (one bulk copy).  One thread issues the copies;
// completion is tracked by mbarriers, so the copy costs no LSU issue slots -- with per-thread cp.async the issue of the copies
// and the read-back took a third of a tile's time (profiles/r01_v5_matvec_*.md).  The gather of the input vector for tile i+1
// is started into registers while tile i runs its segment sums.
//
// The object keeps the ring position across passes, so that a kernel which runs many products (k_pcg_solve) can let the last
// tile of one pass start the copies of the first tile of the NEXT pass (`chain_next`): the Jacobian, the records and the
// (E^T E)^-1 blocks do not change during a linear solve, only the input vector does, and that one is gathered at the start of
// the next pass.  COHERENT: the input vector was written by other CTAs of the SAME launch -- read it at L2 (ld.global.cg).
template <bool TMAP, bool CHUNKED, bool COHERENT>
struct ProductPass {
  // Only the ring position lives across passes; the carve-up of the shared memory is recomputed by every pass into locals (a
  // kernel that keeps ten more pointers alive across its other phases runs out of registers at 2 CTAs / SM and the compiler
  // then re-derives them inside the inner loops: measured 0.316 instead of 0.206 ms per product, profiles/r02_v2_*).
  unsigned done;          // tiles this CTA has consumed so far, over all passes: tile `it` of a pass sits at ring position done + it
  bool chained;           // the first tile of the coming pass has already been issued (by the previous pass)

  struct Smem {
    double2* Jbuf; double* xs; double* v; double* w; double* u; double* ps; double* eibuf; unsigned char* recbuf;
    unsigned long long* bar_full; unsigned long long* bar_rec;
    int UP;
  };
  static __device__ __forceinline__ Smem carve(const BaDev& L, double* sm) {
    Smem m;
    m.Jbuf = reinterpret_cast<double2*>(sm);                 // [12][T] next tile's Jacobian (TMAP: [2][12][T/2])
    m.xs = sm + 2 * kJPlanes * T;                            // [max_seg][9]
    m.v = m.xs + ((L.max_seg_tile * 9 + 1) & ~1);            // [9][VLD]   (xs padded to an even count: 16-byte alignment below)
    m.w = m.v + 9 * VLD;                                     // [3][T]
    m.UP = (L.max_pt_tile + 1) & ~1;
    m.u = m.w + 3 * T;                                       // [3][UP]
    m.ps = m.u + 3 * m.UP;                                   // [seg_chunk_scratch]  chunk sums of the segment sums (CHUNKED)
    m.eibuf = m.ps + seg_chunk_scratch(L.max_seg_tile) + 1;  // 2 x [max_pt][6]   (+1: 9 * VLD is odd)
    m.recbuf = reinterpret_cast<unsigned char*>(m.eibuf + 2 * (size_t)L.max_pt_tile * 6);   // 2 x rec_stride bytes
    m.bar_full = reinterpret_cast<unsigned long long*>(m.recbuf + 2 * (size_t)L.rec_stride);   // [2] Jacobian + einv
    m.bar_rec = m.bar_full + 2;                                                                 // [2] record
    return m;
  }
  static __device__ __forceinline__ int tiles_of_cta(const BaDev& L) {
    return ((int)blockIdx.x < L.n_tiles) ? (L.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  }

  static size_t smem_bytes(const BaDev& L) {
    return sizeof(double2) * kJPlanes * T +
           sizeof(double) * ((size_t)((L.max_seg_tile * 9 + 1) & ~1) + 9 * VLD + 3 * T + 3 * (size_t)((L.max_pt_tile + 1) & ~1) + 1 +
                             (size_t)seg_chunk_scratch(L.max_seg_tile) + 2 * (size_t)L.max_pt_tile * 6) +
           2 * (size_t)L.rec_stride + 4 * sizeof(unsigned long long);
  }

  // Initialises the mbarriers; ends with a CTA barrier.
  __device__ __forceinline__ void init(const BaDev& L, double* sm) {
    done = 0; chained = false;
    if (threadIdx.x == 0) {
      const Smem m = carve(L, sm);
      mbar_init(m.bar_full, 1); mbar_init(m.bar_full + 1, 1); mbar_init(m.bar_rec, 1); mbar_init(m.bar_rec + 1, 1);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
  }

  static __device__ __forceinline__ Tile header(const BaDev& L, int it) {   // chunk tiles of long tracks are empty work items here
    Tile q = load_tile(L, blockIdx.x + it * gridDim.x);
    if (q.chunk >= 0) { q.no = 0; q.np = 0; q.ns = 0; }
    return q;
  }
  // thread 0: everything tile t needs, into ring slot `buf`
  static __device__ __forceinline__ void issue(const Smem& m, const CUtensorMap* tmapJ, const BaDev& L, const double2* J2, const double* einv,
                                               const Tile& q, int t, int buf) {
    double2* Jbuf = m.Jbuf; double* eibuf = m.eibuf; unsigned char* recbuf = m.recbuf;
    unsigned long long* bar_full = m.bar_full; unsigned long long* bar_rec = m.bar_rec;
    const size_t O = (size_t)L.n_obs;
    mbar_expect_tx(bar_rec + buf, (unsigned)L.rec_stride);
    if (L.l2_keep_tiles >= 0) {            // the same copies with an L2 policy: the first l2_keep_tiles tiles stay resident between products
      const unsigned long long pol = l2_policy(t < L.l2_keep_tiles);
      bulk_g2s_hint(recbuf + (size_t)buf * L.rec_stride, L.tile_rec + (size_t)t * L.rec_stride, (unsigned)L.rec_stride, bar_rec + buf, pol);
      if (TMAP) {
        const int boxes = (q.no + T / 2 - 1) / (T / 2);
        mbar_expect_tx(bar_full + buf, (unsigned)(boxes * kJPlanes * (T / 2) * 16 + q.np * 48));
        for (int h = 0; h < boxes; ++h) tma_box_2d_hint(Jbuf + h * kJPlanes * (T / 2), tmapJ, 2 * (q.ob + h * (T / 2)), 0, bar_full + buf, pol);
      } else {
        mbar_expect_tx(bar_full + buf, (unsigned)(q.no * kJPlanes * 16 + q.np * 48));
        if (q.no > 0) {
#pragma unroll
          for (int k = 0; k < kJPlanes; ++k) bulk_g2s_hint(Jbuf + k * T, J2 + k * O + q.ob, (unsigned)q.no * 16u, bar_full + buf, pol);
        }
      }
      if (q.np > 0) bulk_g2s_hint(eibuf + (size_t)buf * L.max_pt_tile * 6, einv + (size_t)q.pb * 6, (unsigned)q.np * 48u, bar_full + buf, pol);
      return;
    }
    bulk_g2s(recbuf + (size_t)buf * L.rec_stride, L.tile_rec + (size_t)t * L.rec_stride, (unsigned)L.rec_stride, bar_rec + buf);
    if (TMAP) {
      static_assert(T == 256, "the tensor-map box is 128 observations: two boxes per tile");
      const int boxes = (q.no + T / 2 - 1) / (T / 2);          // a box past the tile's end only brings the next tile's data
      mbar_expect_tx(bar_full + buf, (unsigned)(boxes * kJPlanes * (T / 2) * 16 + q.np * 48));
      for (int h = 0; h < boxes; ++h) tma_box_2d(Jbuf + h * kJPlanes * (T / 2), tmapJ, 2 * (q.ob + h * (T / 2)), 0, bar_full + buf);
    } else {
      mbar_expect_tx(bar_full + buf, (unsigned)(q.no * kJPlanes * 16 + q.np * 48));
      if (q.no > 0) {
#pragma unroll
        for (int k = 0; k < kJPlanes; ++k) bulk_g2s(Jbuf + k * T, J2 + k * O + q.ob, (unsigned)q.no * 16u, bar_full + buf);
      }
    }
    if (q.np > 0) bulk_g2s(eibuf + (size_t)buf * L.max_pt_tile * 6, einv + (size_t)q.pb * 6, (unsigned)q.np * 48u, bar_full + buf);
  }

  // One product: seg_y = segment partials of S_local * d with d = va (two == false) or va + beta * vb.
  // Element idx of a tile's input vector [ns][9] is fetched as its two RAW operands one tile ahead, behind the segment sums; the
  // multiply-add is left to the consumer on purpose: an arithmetic instruction placed right after the loads would make every warp
  // wait for that L2 round trip on the spot (measured: +34 us per product with the fused form, profiles/r02_v1_summary.md).
  __device__ __forceinline__ void run(const CUtensorMap* tmapJ, const BaDev& L, const double2* J2, const double* va, const double* vb,
                                      double beta, bool two, const double* einv, double* seg_y, bool chain_next, double* sm) {
    const int my_tiles = tiles_of_cta(L);
    if (my_tiles == 0) return;
    const int tid = threadIdx.x;
    const Smem m = carve(L, sm);
    double2* const Jbuf = m.Jbuf; double* const xs = m.xs; double* const v = m.v; double* const w = m.w; double* const u = m.u;
    double* const ps = m.ps; double* const eibuf = m.eibuf; unsigned char* const recbuf = m.recbuf;
    unsigned long long* const bar_full = m.bar_full; unsigned long long* const bar_rec = m.bar_rec;
    const int UP = m.UP;
    auto gather2 = [&](const RecView& R, int idx, double& a, double& b) {
      const int s = idx / 9, k = idx - s * 9;
      const size_t e = (size_t)R.scam[s] * 9 + k;
      a = COHERENT ? __ldcg(va + e) : va[e];
      b = two ? (COHERENT ? __ldcg(vb + e) : vb[e]) : 0.0;
    };
    auto combine = [&](double a, double b) { return two ? __fma_rn(beta, b, a) : a; };
    Tile q = header(L, 0);
    const Tile q0 = q;
    Tile qn = q;
    if (my_tiles > 1) qn = header(L, 1);
    if (!chained && tid == 0) issue(m, tmapJ, L, J2, einv, q, blockIdx.x, (int)(done & 1u));
    mbar_wait(bar_rec + (done & 1u), (done >> 1) & 1u);
    double xpre = 0.0, xpre2 = 0.0;                            // operands of element `tid` of the current tile's input vector
    if (tid < q.ns * 9) gather2(rec_view(L, recbuf + (size_t)(done & 1u) * L.rec_stride), tid, xpre, xpre2);
    for (int it = 0; it < my_tiles; ++it) {
      const unsigned rp = done + (unsigned)it;
      const int cur = (int)(rp & 1u);
      const unsigned par = (rp >> 1) & 1u;
      Tile qnn = qn;
      if (it + 2 < my_tiles) qnn = header(L, it + 2);          // plain loads, consumed in the next iteration
      const RecView R = rec_view(L, recbuf + (size_t)cur * L.rec_stride);
      const double* ei = eibuf + (size_t)cur * L.max_pt_tile * 6;
      const bool active = tid < q.no;
      mbar_wait(bar_full + cur, par);                          // this tile's Jacobian and (E^T E)^-1 have landed
      double2 Fv[9], Ev[3];
      int slot = 0, ptl = 0, rank = 0;
      if (active) {
#pragma unroll
        for (int k = 0; k < 9; ++k) Fv[k] = TMAP ? Jbuf[(tid >> 7) * kJPlanes * (T / 2) + k * (T / 2) + (tid & 127)] : Jbuf[k * T + tid];
#pragma unroll
        for (int k = 0; k < 3; ++k) Ev[k] = TMAP ? Jbuf[(tid >> 7) * kJPlanes * (T / 2) + (9 + k) * (T / 2) + (tid & 127)] : Jbuf[(9 + k) * T + tid];
        slot = R.slot[tid]; ptl = R.ptl[tid];
        if (CHUNKED) rank = R.srank[tid];
      }
      if (tid < q.ns * 9) xs[tid] = combine(xpre, xpre2);
      for (int idx = tid + T; idx < q.ns * 9; idx += T) { double a, b; gather2(R, idx, a, b); xs[idx] = combine(a, b); }   // more than 28 segments: the rest, not prefetched
      __syncthreads();                                         // xs complete; everyone has taken its Jacobian out of Jbuf
      if (tid == 0) {
        if (it + 1 < my_tiles) issue(m, tmapJ, L, J2, einv, qn, blockIdx.x + (it + 1) * gridDim.x, cur ^ 1);
        else if (chain_next) issue(m, tmapJ, L, J2, einv, q0, blockIdx.x, cur ^ 1);   // first tile of the next pass
      }
      double t0 = 0.0, t1 = 0.0;
      if (active) {
#pragma unroll
        for (int k = 0; k < 9; ++k) { const double xk = xs[slot * 9 + k]; t0 = __fma_rn(Fv[k].x, xk, t0); t1 = __fma_rn(Fv[k].y, xk, t1); }
#pragma unroll
        for (int k = 0; k < 3; ++k) w[CHUNKED ? tid * 3 + k : k * T + tid] = two_rows(Ev[k].x, t0, Ev[k].y, t1);
      }
      __syncthreads();
      if (CHUNKED) point_sums_chunked(R, q.np, w, v, ei, u, UP);
      else {
        if (tid < q.np) {
          const int b = R.pptr[tid], e = R.pptr[tid + 1];
          double a0 = 0.0, a1 = 0.0, a2 = 0.0;
          for (int j = b; j < e; ++j) { a0 += w[j]; a1 += w[T + j]; a2 += w[2 * T + j]; }
          const double* m = ei + tid * 6;
          u[tid] = m[0] * a0 + m[1] * a1 + m[2] * a2;
          u[UP + tid] = m[1] * a0 + m[3] * a1 + m[4] * a2;
          u[2 * UP + tid] = m[2] * a0 + m[4] * a1 + m[5] * a2;
        }
        __syncthreads();
      }
      if (active) {
        const double u0 = u[ptl], u1 = u[UP + ptl], u2 = u[2 * UP + ptl];
        const double s0 = __dsub_rn(t0, __fma_rn(Ev[2].x, u2, __fma_rn(Ev[1].x, u1, __dmul_rn(Ev[0].x, u0))));
        const double s1 = __dsub_rn(t1, __fma_rn(Ev[2].y, u2, __fma_rn(Ev[1].y, u1, __dmul_rn(Ev[0].y, u0))));
        double* vt = CHUNKED ? v + rank * VS : v + tid;
#pragma unroll
        for (int k = 0; k < 9; ++k) vt[CHUNKED ? k : k * VLD] = two_rows(Fv[k].x, s0, Fv[k].y, s1);
      }
      __syncthreads();
      if (it + 1 < my_tiles) {                                 // start the next tile's input gather behind the segment sums
        mbar_wait(bar_rec + (cur ^ 1), ((rp + 1u) >> 1) & 1u);
        if (tid < qn.ns * 9) gather2(rec_view(L, recbuf + (size_t)(cur ^ 1) * L.rec_stride), tid, xpre, xpre2);
      }
      if (CHUNKED) seg_sums_chunked(R, q.ns, q.sb, v, ps, seg_y);
      else {
        for (int idx = tid; idx < q.ns * 9; idx += T) {
          const int s = idx / 9, k = idx - s * 9;
          const int b = R.sptr[s], e = R.sptr[s + 1];
          const double* vk = v + k * VLD;
          double sum = 0.0;
          int pos = b;
          for (; pos + 4 <= e; pos += 4) {
            const int i0 = R.sperm[pos], i1 = R.sperm[pos + 1], i2 = R.sperm[pos + 2], i3 = R.sperm[pos + 3];
            const double x0 = vk[i0], x1 = vk[i1], x2 = vk[i2], x3 = vk[i3];
            sum += x0; sum += x1; sum += x2; sum += x3;
          }
          for (; pos < e; ++pos) sum += vk[R.sperm[pos]];
          seg_y[(size_t)R.spos[s] * 9 + k] = sum;
        }
      }
      q = qn; qn = qnn;
    }
    done += (unsigned)my_tiles; chained = chain_next;
  }

  // Before the kernel ends: a chained first tile that no pass will consume must have landed (no copy may outlive the CTA).
  __device__ __forceinline__ void drain(const BaDev& L, double* sm) {
    if (tiles_of_cta(L) == 0 || !chained) return;
    const Smem m = carve(L, sm);
    mbar_wait(m.bar_rec + (done & 1u), (done >> 1) & 1u);
    mbar_wait(m.bar_full + (done & 1u), (done >> 1) & 1u);
    chained = false;
  }
};

}  // namespace
}  // namespace sk
