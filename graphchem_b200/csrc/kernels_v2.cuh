// Second-generation row-wise kernels: same arithmetic and the same per-row summation order as
// kernels.cuh (the embedding-gradient kernel uses a different, equally fixed, order), restructured
// for memory-level parallelism:
//
//   * a row is owned by a GROUP of CH = D / (16 bytes) lanes (8, 16 or 32 -> 4, 2 or 1 rows per warp);
//   * the first four CSR entries of every row come from the packed batch's ELL-4 view (one int4
//     broadcast load, -1 padded): no pointer -> index -> data chain, no data-dependent loop, no
//     divergence between rows of different degree; rows with more than four entries finish in a
//     short CSR loop (rare for molecules: valence <= 4);
//   * all neighbour-row loads of a row are issued back to back and kept RAW (4 registers per
//     16-byte chunk) until they are consumed;
//   * max / tie counting runs on packed bf16x2 (exact), sums accumulate in fp32;
//   * all index arithmetic is 32-bit with compile-time D (no 64-bit divisions).
//
// Used when D / vec is 8, 16 or 32; other widths fall back to kernels.cuh.
#pragma once
#include "kernels.cuh"

namespace gcb {
namespace v2 {

constexpr int kThreads = 256;
template <int CH> constexpr int rows_per_block() { return (kThreads / 32) * (32 / CH); }
template <int CH> inline unsigned blocks_for(int rows) { return (unsigned)ceil_div(rows, rows_per_block<CH>()); }
// Gather kernels walk rows with a grid stride and prefetch the NEXT row's ELL/CSR indices while the
// current row's data is in flight; the grid is capped so that every thread sees several rows.
template <int CH> inline unsigned blocks_strided(int rows) {
  const unsigned need = blocks_for<CH>(rows), cap = (unsigned)kNumSMs * 8u;
  return need < cap ? need : cap;
}

#define GCB_V2_PROLOGUE(ROWS)                                                                        \
  constexpr int V = Vec<T>::N;                                                                       \
  constexpr int D = CH * V;                                                                          \
  constexpr int RPW = 32 / CH;                                                                       \
  const int lane = threadIdx.x & 31;                                                                 \
  const int lig = lane & (CH - 1);                                                                   \
  const int gig = lane / CH;                                                                         \
  const unsigned gmask = CH == 32 ? 0xffffffffu : (((1u << (CH & 31)) - 1u) << (gig * CH));          \
  const int row_raw = ((int)blockIdx.x * (kThreads / 32) + ((int)threadIdx.x >> 5)) * RPW + gig;     \
  if (row_raw >= (ROWS)) return;                                                                     \
  const int r = row_raw; /* `rev` (reverse traversal) made no measurable difference and is ignored */ \
  (void)rev;                                                                                         \
  const int c = lig * V;                                                                             \
  (void)D; (void)gmask; (void)c;

// 16-byte chunks are kept RAW (4 registers) between the load and the point of use, so that four
// neighbour rows in flight cost 16 registers instead of 32.
template <typename T> struct Raw;
template <> struct Raw<float> {
  static __device__ __forceinline__ uint4 load(const float* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ void to_float(const uint4& r, float (&v)[4]) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  }
  static __device__ __forceinline__ uint4 neg_inf() { return make_uint4(0xFF800000u, 0xFF800000u, 0xFF800000u, 0xFF800000u); }
  static __device__ __forceinline__ uint4 zero() { return make_uint4(0u, 0u, 0u, 0u); }
  static __device__ __forceinline__ uint32_t cnt1(uint32_t c, uint32_t q, uint32_t m) {
    return __float_as_uint(__uint_as_float(c) + (__uint_as_float(q) == __uint_as_float(m) ? 1.f : 0.f));
  }
  static __device__ __forceinline__ void count_eq(uint4& c, const uint4& q, const uint4& m) {
    c.x = cnt1(c.x, q.x, m.x); c.y = cnt1(c.y, q.y, m.y); c.z = cnt1(c.z, q.z, m.z); c.w = cnt1(c.w, q.w, m.w);
  }
  // bit i set <=> channel i of q equals channel i of m (channels in to_float order)
  static __device__ __forceinline__ uint32_t eq_bits(const uint4& q, const uint4& m) {
    return (__uint_as_float(q.x) == __uint_as_float(m.x) ? 1u : 0u) | (__uint_as_float(q.y) == __uint_as_float(m.y) ? 2u : 0u) |
           (__uint_as_float(q.z) == __uint_as_float(m.z) ? 4u : 0u) | (__uint_as_float(q.w) == __uint_as_float(m.w) ? 8u : 0u);
  }
  static __device__ __forceinline__ uint4 vmax(const uint4& a, const uint4& b) {
    return make_uint4(__float_as_uint(fmaxf(__uint_as_float(a.x), __uint_as_float(b.x))),
                      __float_as_uint(fmaxf(__uint_as_float(a.y), __uint_as_float(b.y))),
                      __float_as_uint(fmaxf(__uint_as_float(a.z), __uint_as_float(b.z))),
                      __float_as_uint(fmaxf(__uint_as_float(a.w), __uint_as_float(b.w))));
  }
};
template <> struct Raw<bf16> {
  static __device__ __forceinline__ uint4 load(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ void to_float(const uint4& r, float (&v)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
  }
  static __device__ __forceinline__ uint4 neg_inf() { return make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u); }
  static __device__ __forceinline__ uint4 zero() { return make_uint4(0u, 0u, 0u, 0u); }
  static __device__ __forceinline__ uint32_t cnt2(uint32_t c, uint32_t q, uint32_t m) {  // c += (q == m) per bf16 lane
    __nv_bfloat162 cc = *reinterpret_cast<__nv_bfloat162*>(&c), qq = *reinterpret_cast<__nv_bfloat162*>(&q),
                   mm = *reinterpret_cast<__nv_bfloat162*>(&m);
    cc = __hadd2(cc, __heq2(qq, mm));
    return *reinterpret_cast<uint32_t*>(&cc);
  }
  static __device__ __forceinline__ void count_eq(uint4& c, const uint4& q, const uint4& m) {
    c.x = cnt2(c.x, q.x, m.x); c.y = cnt2(c.y, q.y, m.y); c.z = cnt2(c.z, q.z, m.z); c.w = cnt2(c.w, q.w, m.w);
  }
  static __device__ __forceinline__ uint32_t eq2(uint32_t q, uint32_t m) {  // 2 bits: (lo == lo) | (hi == hi) << 1
    const unsigned k = __heq2_mask(*reinterpret_cast<__nv_bfloat162*>(&q), *reinterpret_cast<__nv_bfloat162*>(&m));
    return (k & 1u) | ((k >> 15) & 2u);
  }
  static __device__ __forceinline__ uint32_t eq_bits(const uint4& q, const uint4& m) {
    return eq2(q.x, m.x) | (eq2(q.y, m.y) << 2) | (eq2(q.z, m.z) << 4) | (eq2(q.w, m.w) << 6);
  }
  static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
    __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a), y = *reinterpret_cast<__nv_bfloat162*>(&b);
    __nv_bfloat162 m = __hmax2(x, y);
    return *reinterpret_cast<uint32_t*>(&m);
  }
  static __device__ __forceinline__ uint4 vmax(const uint4& a, const uint4& b) {  // exact in bf16
    return make_uint4(max2(a.x, b.x), max2(a.y, b.y), max2(a.z, b.z), max2(a.w, b.w));
  }
};
// acc_lo += bf16 low half of w, acc_hi += bf16 high half: Blackwell's mixed-precision add (PTX add.f32.bf16, SASS
// FHADD.BF16 with an .H1 operand selector) takes the packed halves as they are -- ONE instruction per element
// instead of unpack (shift / mask) + FADD, and the same fp32 result bit for bit (fp32(bf16) + fp32, round to nearest).
__device__ __forceinline__ void add_bf16x2(float& lo, float& hi, uint32_t w) {
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tadd.f32.bf16 %0, l, %0;\n\tadd.f32.bf16 %1, h, %1;\n\t}"
      : "+f"(lo), "+f"(hi)
      : "r"(w));
}
__device__ __forceinline__ void add_bf16_lo(float& acc, uint32_t w) {
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tadd.f32.bf16 %0, l, %0;\n\t}" : "+f"(acc) : "r"(w));
}
__device__ __forceinline__ void add_bf16_hi(float& acc, uint32_t w) {
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tadd.f32.bf16 %0, h, %0;\n\t}" : "+f"(acc) : "r"(w));
}
// acc[i] += channel i of the bf16 chunk r where bit i of `bits` is set (the amax backward routes a gradient only
// to the in-edges that attained the maximum): a predicate per bit and a predicated mixed-precision add
__device__ __forceinline__ void raw_add_masked(const uint4& r, uint32_t bits, float (&acc)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (bits & (1u << (2 * i))) add_bf16_lo(acc[2 * i], w[i]);
    if (bits & (2u << (2 * i))) add_bf16_hi(acc[2 * i + 1], w[i]);
  }
}
template <typename T, int V>
__device__ __forceinline__ void raw_add(const uint4& r, float (&acc)[V]) {
  if constexpr (sizeof(T) == 2) {
    static_assert(sizeof(T) != 2 || V == 8, "bf16 chunks hold eight channels");
    add_bf16x2(acc[0], acc[1], r.x); add_bf16x2(acc[2], acc[3], r.y);
    add_bf16x2(acc[4], acc[5], r.z); add_bf16x2(acc[6], acc[7], r.w);
  } else {
    float v[V];
    Raw<T>::to_float(r, v);
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] += v[i];
  }
}
template <typename T, int V>
__device__ __forceinline__ void raw_fma(float w, const uint4& r, float (&acc)[V]) {
  float v[V];
  Raw<T>::to_float(r, v);
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = fmaf(w, v[i], acc[i]);
}

template <typename T, int V>
__device__ __forceinline__ void vzero(float (&a)[V]) {
#pragma unroll
  for (int i = 0; i < V; ++i) a[i] = 0.f;
}

// Bond update: Bout[r] = dropout(act(P[r] + max_{e->r} Q[src e])), act(0) for rows without in-edges.
template <typename T, int CH, bool DROP, bool TIES>
__global__ void __launch_bounds__(kThreads, TIES ? 4 : 6)
bond_update_kernel(const T* __restrict__ PQ, const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_src,
                   const int32_t* __restrict__ in_eid, const int4* __restrict__ in_src4,
                   const int4* __restrict__ in_eid4, T* __restrict__ Bout, uint8_t* __restrict__ tiemask,
                   uint8_t* __restrict__ tiecnt, int M, int act, Drop dr, int rev) {
  GCB_V2_PROLOGUE(M)
  const int4 s4 = __ldg(in_src4 + r);
  const int e0 = __ldg(in_ptr + r), e1 = __ldg(in_ptr + r + 1);
  const uint4 praw = Raw<T>::load(PQ + (size_t)r * (2 * D) + c);
  const int sidx[4] = {s4.x, s4.y, s4.z, s4.w};
  uint4 q[4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (sidx[u] >= 0) q[u] = Raw<T>::load(PQ + (size_t)sidx[u] * (2 * D) + D + c);
  uint4 mraw = Raw<T>::neg_inf();
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (sidx[u] >= 0) mraw = Raw<T>::vmax(mraw, q[u]);
  for (int e = e0 + 4; e < e1; ++e)
    mraw = Raw<T>::vmax(mraw, Raw<T>::load(PQ + (size_t)__ldg(in_src + e) * (2 * D) + D + c));
  if (TIES && e1 > e0) {
    // Training: record, per in-edge, which channels attain the maximum (one byte per lane) and, per
    // channel, how many edges tie -- everything the amax backward needs, so that it never has to
    // gather Q again (torch splits the gradient evenly among tied maxima).
    // tiemask is indexed by dst-CSR position (rows write consecutive bytes: coalesced)
    uint4 craw = Raw<T>::zero();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (sidx[u] >= 0) {
        Raw<T>::count_eq(craw, q[u], mraw);
        tiemask[(size_t)(e0 + u) * CH + lig] = (uint8_t)Raw<T>::eq_bits(q[u], mraw);
      }
    }
    for (int e = e0 + 4; e < e1; ++e) {
      const uint4 qe = Raw<T>::load(PQ + (size_t)__ldg(in_src + e) * (2 * D) + D + c);
      Raw<T>::count_eq(craw, qe, mraw);
      tiemask[(size_t)e * CH + lig] = (uint8_t)Raw<T>::eq_bits(qe, mraw);
    }
    float cf[V];
    Raw<T>::to_float(craw, cf);
    uint8_t* cp = tiecnt + (size_t)r * D + c;
    if constexpr (V == 8) {
      uint2 pk;
      pk.x = (uint32_t)cf[0] | ((uint32_t)cf[1] << 8) | ((uint32_t)cf[2] << 16) | ((uint32_t)cf[3] << 24);
      pk.y = (uint32_t)cf[4] | ((uint32_t)cf[5] << 8) | ((uint32_t)cf[6] << 16) | ((uint32_t)cf[7] << 24);
      *reinterpret_cast<uint2*>(cp) = pk;
    } else {
      *reinterpret_cast<uint32_t*>(cp) =
          (uint32_t)cf[0] | ((uint32_t)cf[1] << 8) | ((uint32_t)cf[2] << 16) | ((uint32_t)cf[3] << 24);
    }
  }
  float o[V];
  if (e0 == e1) {
    const float z = act_fwd_t<T>(0.f, act);
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] = z;
  } else {
    float p[V], mx[V];
    Raw<T>::to_float(praw, p);
    Raw<T>::to_float(mraw, mx);
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] = p[i] + mx[i];
    act_fwd_vec<T, V>(o, act);
  }
  if (DROP) apply_dropout<T, V>(o, dr, (uint64_t)r * D + c);
  store_vec(Bout + (size_t)r * D + c, o);
}

// Atom aggregation: S[i] = [ sum_{e->i} A[src e] | sum_{e->i} Bnew[eid e] ]  (Bnew[e >= M] = act(0)).
// HALF = 0: only the A half, 1: only the B half, 2: both in one pass.  Two single-stream passes keep
// each kernel at <= 40 registers (full occupancy), which is faster than one two-stream pass.
// DROP doubles as the "generic" switch: dropout and mean aggregation are only compiled in when set.
template <typename T, int CH, bool DROP, int HALF>
__global__ void __launch_bounds__(kThreads, HALF == 2 ? 4 : 6)
atom_gather_kernel(const T* __restrict__ A, const T* __restrict__ Bnew, const int32_t* __restrict__ in_ptr,
                   const int32_t* __restrict__ in_src, const int32_t* __restrict__ in_eid,
                   const int4* __restrict__ in_src4, const int4* __restrict__ in_eid4, T* __restrict__ S, int N, int M,
                   int act, int aggr, Drop dr_bond, int rev) {
  GCB_V2_PROLOGUE(N)
  constexpr bool DO_A = HALF != 1, DO_B = HALF != 0;
  const int e0 = __ldg(in_ptr + r), e1 = __ldg(in_ptr + r + 1);
  float sa[V], sb[V];
  vzero<T, V>(sa);
  vzero<T, V>(sb);
  if (DO_A) {
    const int4 s4 = __ldg(in_src4 + r);
    const int sidx[4] = {s4.x, s4.y, s4.z, s4.w};
    uint4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (sidx[u] >= 0) a[u] = Raw<T>::load(A + (size_t)sidx[u] * D + c);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (sidx[u] >= 0) raw_add<T, V>(a[u], sa);
    for (int e = e0 + 4; e < e1; ++e) raw_add<T, V>(Raw<T>::load(A + (size_t)__ldg(in_src + e) * D + c), sa);
  }
  if (DO_B) {
    const int4 e4 = __ldg(in_eid4 + r);
    const int eidx[4] = {e4.x, e4.y, e4.z, e4.w};
    const float cst = round_to(act_fwd_t<T>(0.f, act), (const T*)nullptr);
    auto add_const = [&](int eid) {
      float bc[V];
#pragma unroll
      for (int i = 0; i < V; ++i) bc[i] = cst;
      if (DROP) apply_dropout<T, V>(bc, dr_bond, (uint64_t)eid * D + c);
#pragma unroll
      for (int i = 0; i < V; ++i) sb[i] += round_to(bc[i], (const T*)nullptr);
    };
    uint4 b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (eidx[u] >= 0 && eidx[u] < M) b[u] = Raw<T>::load(Bnew + (size_t)eidx[u] * D + c);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (eidx[u] >= 0) {
        if (eidx[u] < M) raw_add<T, V>(b[u], sb);
        else add_const(eidx[u]);
      }
    }
    for (int e = e0 + 4; e < e1; ++e) {
      const int eid = __ldg(in_eid + e);
      if (eid < M) raw_add<T, V>(Raw<T>::load(Bnew + (size_t)eid * D + c), sb);
      else add_const(eid);
    }
  }
  if ((DROP && aggr == GCB_AGGR_MEAN) && e1 > e0) {
    const float inv = 1.f / (float)(e1 - e0);
#pragma unroll
    for (int i = 0; i < V; ++i) { sa[i] *= inv; sb[i] *= inv; }
  }
  if (DO_A) store_vec(S + (size_t)r * (2 * D) + c, sa);
  if (DO_B) store_vec(S + (size_t)r * (2 * D) + D + c, sb);
}

// dZA_prev[j] = ( dZA[j] + sum_{e: src e = j} w(dst e) dS[dst e, 0:D] ) * dropfactor * act'(A_prev[j])
template <typename T, int CH, bool DROP>
__global__ void __launch_bounds__(kThreads, 5)
atom_bwd_gather_kernel(const T* __restrict__ dZA, const T* __restrict__ dS, const int32_t* __restrict__ out_ptr,
                       const int32_t* __restrict__ out_dst, const int4* __restrict__ out_dst4,
                       const int32_t* __restrict__ in_ptr, const T* __restrict__ A_prev, T* __restrict__ dZA_prev, int N,
                       int act, int aggr, Drop dr_prev, int rev) {
  GCB_V2_PROLOGUE(N)
  const int4 t4 = __ldg(out_dst4 + r);
  const int e0 = __ldg(out_ptr + r), e1 = __ldg(out_ptr + r + 1);
  const uint4 draw = Raw<T>::load(dZA + (size_t)r * D + c);
  const uint4 yraw = Raw<T>::load(A_prev + (size_t)r * D + c);
  const int tidx[4] = {t4.x, t4.y, t4.z, t4.w};
  uint4 g[4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (tidx[u] >= 0) g[u] = Raw<T>::load(dS + (size_t)tidx[u] * (2 * D) + c);
  float d[V], y[V];
  Raw<T>::to_float(draw, d);
  Raw<T>::to_float(yraw, y);
  auto weight = [&](int t) -> float {
    return (DROP && aggr == GCB_AGGR_MEAN) ? 1.f / (float)max(__ldg(in_ptr + t + 1) - __ldg(in_ptr + t), 1) : 1.f;
  };
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (tidx[u] >= 0) {
      if ((DROP && aggr == GCB_AGGR_MEAN)) raw_fma<T, V>(weight(tidx[u]), g[u], d);
      else raw_add<T, V>(g[u], d);
    }
  }
  for (int e = e0 + 4; e < e1; ++e) {
    const int t = __ldg(out_dst + e);
    raw_fma<T, V>(weight(t), Raw<T>::load(dS + (size_t)t * (2 * D) + c), d);
  }
  if (DROP) {
    mul_act_grad<V>(d, y, act, dr_prev, (uint64_t)r * D + c);
  } else {
    act_grad_mul_vec<T, V>(d, y, act);
  }
  store_vec(dZA_prev + (size_t)r * D + c, d);
}

// Destination side of the bond backward.  The tie bits and tie counts were recorded by the forward
// bond update, so this is a pure row-wise kernel:
//   dzb[r]    = ( w(dst r) dS[dst r, D:2D] + dBnext[r] + d_out_bond[r] ) * dropfactor * act'(B_l[r])   (0 if deg_in(r) = 0)
//   dPQ[r,:D] = dzb[r]  (= dP),   gsplit[r] = dzb[r] / count[r]
template <typename T, int CH, bool DROP>
__global__ void __launch_bounds__(kThreads, 6)
bond_bwd_tie_kernel(const T* __restrict__ dS, const T* __restrict__ dBnext, const T* __restrict__ d_out_bond,
                    const T* __restrict__ B_l, const int32_t* __restrict__ dst, const int32_t* __restrict__ in_ptr,
                    const uint8_t* __restrict__ tiecnt, T* __restrict__ dPQ, T* __restrict__ gsplit, int M, int act,
                    int aggr, Drop dr, int rev) {
  GCB_V2_PROLOGUE(M)
  const int e0 = __ldg(in_ptr + r), e1 = __ldg(in_ptr + r + 1);
  float d[V], g[V];
#pragma unroll
  for (int i = 0; i < V; ++i) d[i] = g[i] = 0.f;
  if (e1 > e0) {
    const int t = __ldg(dst + r);
    const uint4 sraw = Raw<T>::load(dS + (size_t)t * (2 * D) + D + c);
    const uint4 yraw = Raw<T>::load(B_l + (size_t)r * D + c);
    uint32_t cw[2] = {0u, 0u};
    if constexpr (V == 8) {
      const uint2 pk = __ldg(reinterpret_cast<const uint2*>(tiecnt + (size_t)r * D + c));
      cw[0] = pk.x; cw[1] = pk.y;
    } else {
      cw[0] = __ldg(reinterpret_cast<const uint32_t*>(tiecnt + (size_t)r * D + c));
    }
    Raw<T>::to_float(sraw, d);
    if ((DROP && aggr == GCB_AGGR_MEAN)) {
      const float w = 1.f / (float)max(__ldg(in_ptr + t + 1) - __ldg(in_ptr + t), 1);
#pragma unroll
      for (int i = 0; i < V; ++i) d[i] *= w;
    }
    if (dBnext) raw_add<T, V>(Raw<T>::load(dBnext + (size_t)r * D + c), d);
    if (d_out_bond) raw_add<T, V>(Raw<T>::load(d_out_bond + (size_t)r * D + c), d);
    float y[V];
    Raw<T>::to_float(yraw, y);
    if (DROP) {
      mul_act_grad<V>(d, y, act, dr, (uint64_t)r * D + c);
    } else {
      act_grad_mul_vec<T, V>(d, y, act);
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const uint32_t cnt = (cw[i >> 2] >> ((i & 3) * 8)) & 0xFFu;
      g[i] = cnt <= 1u ? d[i] : __fdividef(d[i], (float)cnt);
    }
  }
  store_vec(dPQ + (size_t)r * (2 * D) + c, d);
  store_vec(gsplit + (size_t)r * D + c, g);
}

// Source side: dPQ[j, D:2D] = sum_{e: src e = j} tie(e) * gsplit[dst e]   (tie bits from tiemask[dst-CSR position of e]).
template <typename T, int CH, bool DROP>
__global__ void __launch_bounds__(kThreads, 6)
bond_bwd_scatter_kernel(const T* __restrict__ gsplit, const uint8_t* __restrict__ tiemask,
                        const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_dst,
                        const int32_t* __restrict__ out_eid, const int4* __restrict__ out_dst4,
                        const int4* __restrict__ out_eid4, T* __restrict__ dPQ, int M, int rev) {
  GCB_V2_PROLOGUE(M)
  const int4 t4 = __ldg(out_dst4 + r), o4 = __ldg(out_eid4 + r);
  const int e0 = __ldg(out_ptr + r), e1 = __ldg(out_ptr + r + 1);
  const int tidx[4] = {t4.x, t4.y, t4.z, t4.w}, oidx[4] = {o4.x, o4.y, o4.z, o4.w};
  uint4 g[4];
  uint32_t bits[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (tidx[u] >= 0) {
      g[u] = Raw<T>::load(gsplit + (size_t)tidx[u] * D + c);
      bits[u] = __ldg(tiemask + (size_t)oidx[u] * CH + lig);
    }
  }
  float acc[V];
  vzero<T, V>(acc);
  auto consume = [&](const uint4& gr, uint32_t b) {
    float gf[V];
    Raw<T>::to_float(gr, gf);
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] += ((b >> i) & 1u) ? gf[i] : 0.f;
  };
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (tidx[u] >= 0) consume(g[u], bits[u]);
  for (int e = e0 + 4; e < e1; ++e) {
    const int t = __ldg(out_dst + e);
    consume(Raw<T>::load(gsplit + (size_t)t * D + c), __ldg(tiemask + (size_t)__ldg(out_eid + e) * CH + lig));
  }
  store_vec(dPQ + (size_t)r * (2 * D) + D + c, acc);
}

// dZA_L[n] = (dPool[graph(n)] + d_out_atom[n]) * dropfactor * act'(A_L[n]); one group per NODE
// (node -> graph map from the packed batch).
template <typename T, int CH, bool DROP>
__global__ void __launch_bounds__(kThreads)
pool_bwd_kernel(const float* __restrict__ dPool, const T* __restrict__ d_out_atom, const T* __restrict__ A_last,
                const int32_t* __restrict__ node_graph, T* __restrict__ dZA, int N, int act, Drop dr, int rev,
                const int32_t* __restrict__ mean_ptr) {
  // mean_ptr: graph_ptr when the pooling is a mean (every node gets dPool / node count of its graph), else null
  GCB_V2_PROLOGUE(N)
  float d[V], y[V];
  vzero<T, V>(d);
  if (dPool) {
    const int g = __ldg(node_graph + r);
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(dPool + (size_t)g * D + c + i));
      d[i] = t.x; d[i + 1] = t.y; d[i + 2] = t.z; d[i + 3] = t.w;
    }
    if (mean_ptr) {
      const float w = 1.f / (float)max(__ldg(mean_ptr + g + 1) - __ldg(mean_ptr + g), 1);
#pragma unroll
      for (int i = 0; i < V; ++i) d[i] *= w;
    }
  }
  if (d_out_atom) {
    float u[V];
    load_vec(d_out_atom + (size_t)r * D + c, u);
#pragma unroll
    for (int i = 0; i < V; ++i) d[i] += u[i];
  }
  load_vec(A_last + (size_t)r * D + c, y);
  if (DROP) {
    mul_act_grad<V>(d, y, act, dr, (uint64_t)r * D + c);
  } else {
    act_grad_mul_vec<T, V>(d, y, act);
  }
  store_vec(dZA + (size_t)r * D + c, d);
}

// Embedding gradient: one CTA per (token, split); 256 threads = (256/CH) row lanes x CH chunks.  Row
// lane k sums rows k, k+RL, k+2RL, ... of its slice, then the row lanes are combined in lane order.
template <typename T, int CH, bool HAS_Y>
__global__ void __launch_bounds__(kThreads, 4)
embed_bwd_partial_kernel(const T* __restrict__ dZ, const T* __restrict__ Y, const int32_t* __restrict__ tok_ptr,
                         const int32_t* __restrict__ tok_row, float* __restrict__ partial, int vocab, int splits, int act) {
  constexpr int V = Vec<T>::N;
  constexpr int D = CH * V;
  constexpr int RL = kThreads / CH;  // row lanes
  constexpr int UN = HAS_Y ? 4 : 8;  // rows in flight per thread
  __shared__ float red[RL][D + 4];
  const int t = blockIdx.x / splits, s = blockIdx.x % splits;
  const int lig = threadIdx.x % CH, rl = threadIdx.x / CH, c = lig * V;
  const int b0 = __ldg(tok_ptr + t), b1 = __ldg(tok_ptr + t + 1);
  const int per = (b1 - b0 + splits - 1) / splits;
  const int lo = b0 + s * per, hi = min(b1, lo + per);
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  // UN rows in flight per thread (index loads first, then the row loads, kept as raw 16-byte words until they are
  // added): the walk is a dependent index -> row chain and is latency bound; the summation order is unchanged
  int k = lo + rl;
  for (; k + (UN - 1) * RL < hi; k += UN * RL) {
    int row[UN];
#pragma unroll
    for (int j = 0; j < UN; ++j) row[j] = __ldg(tok_row + k + j * RL);
    uint4 d[UN], y[HAS_Y ? UN : 1];
#pragma unroll
    for (int j = 0; j < UN; ++j) {
      d[j] = Raw<T>::load(dZ + (size_t)row[j] * D + c);
      if constexpr (HAS_Y) y[j] = Raw<T>::load(Y + (size_t)row[j] * D + c);
    }
#pragma unroll
    for (int j = 0; j < UN; ++j) {
      float df[V];
      Raw<T>::to_float(d[j], df);
      if constexpr (HAS_Y) {
        float yf[V];
        Raw<T>::to_float(y[j], yf);
        act_grad_mul_vec<T, V>(df, yf, act);
      }
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += df[i];
    }
  }
  for (; k < hi; k += RL) {
    const int row = __ldg(tok_row + k);
    float d[V];
    load_vec(dZ + (size_t)row * D + c, d);
    if constexpr (HAS_Y) {
      float y[V];
      load_vec(Y + (size_t)row * D + c, y);
      act_grad_mul_vec<T, V>(d, y, act);
    }
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] += d[i];
  }
#pragma unroll
  for (int i = 0; i < V; ++i) red[rl][c + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < D) {
    float sum = 0.f;
#pragma unroll 4
    for (int k = 0; k < RL; ++k) sum += red[k][threadIdx.x];
    partial[((size_t)s * vocab + t) * D + threadIdx.x] = sum;
  }
}

}  // namespace v2
}  // namespace gcb
