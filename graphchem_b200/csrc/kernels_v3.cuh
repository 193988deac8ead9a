// Third-generation gather kernels: software-pipelined with cp.async (LDGSTS).
//
// The row kernels are latency bound (ncu: long-scoreboard stalls, ~25 resident warps, each with one
// row in flight).  Here every thread keeps THREE rows in flight without spending registers: the
// 16-byte chunks a row needs (its own chunk plus up to four neighbour chunks from the ELL-4 view)
// are copied global -> shared with cp.async into a per-thread private ring of three stages; row
// i+2 is issued and the indices of row i+3 are fetched before row i is consumed.  Slots are private
// to the issuing thread, so cp.async.wait_group is the only synchronisation.  Arithmetic and
// summation order are those of kernels_v2.cuh.
#pragma once
#include "kernels_v2.cuh"

namespace gcb {
namespace v3 {

using v2::Raw;
using v2::kThreads;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kStages = 3;
template <int NL> constexpr int smem_bytes() { return kStages * NL * kThreads * 16; }

#define GCB_V3_PROLOGUE(ROWS)                                                                        \
  constexpr int V = Vec<T>::N;                                                                       \
  constexpr int D = CH * V;                                                                          \
  constexpr int RPW = 32 / CH;                                                                       \
  extern __shared__ uint4 v3_smem[];                                                                 \
  const int lane = threadIdx.x & 31;                                                                 \
  const int lig = lane & (CH - 1);                                                                   \
  const int gig = lane / CH;                                                                         \
  const int row_raw = ((int)blockIdx.x * (kThreads / 32) + ((int)threadIdx.x >> 5)) * RPW + gig;     \
  const int row_stride = (int)gridDim.x * (kThreads / 32) * RPW;                                     \
  const int c = lig * V;                                                                             \
  auto slot = [&](int st, int l) -> uint4* { return v3_smem + ((st * NL + l) * kThreads + (int)threadIdx.x); };

// Bond update: Bout[r] = dropout(act(P[r] + max_{e->r} Q[src e])), act(0) for rows without in-edges.
template <typename T, int CH>
__global__ void __launch_bounds__(kThreads, 3)
bond_update_kernel(const T* __restrict__ PQ, const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_src,
                   const int4* __restrict__ in_src4, T* __restrict__ Bout, int M, int act, Drop dr) {
  constexpr int NL = 5;  // P chunk + 4 neighbour Q chunks
  GCB_V3_PROLOGUE(M)
  struct Idx { int4 s4; int e0, e1, r; };
  auto load_idx = [&](int row) {
    Idx ix;
    ix.r = row;
    if (row < M) { ix.s4 = __ldg(in_src4 + row); ix.e0 = __ldg(in_ptr + row); ix.e1 = __ldg(in_ptr + row + 1); }
    else { ix.s4 = make_int4(-1, -1, -1, -1); ix.e0 = ix.e1 = 0; }
    return ix;
  };
  auto issue = [&](const Idx& ix, int st) {
    if (ix.r < M) {
      cp_async16(slot(st, 0), PQ + (size_t)ix.r * (2 * D) + c);
      const int s[4] = {ix.s4.x, ix.s4.y, ix.s4.z, ix.s4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (s[u] >= 0) cp_async16(slot(st, 1 + u), PQ + (size_t)s[u] * (2 * D) + D + c);
    }
    cp_async_commit();
  };
  Idx i0 = load_idx(row_raw), i1 = load_idx(row_raw + row_stride), i2 = load_idx(row_raw + 2 * row_stride);
  issue(i0, 0);
  issue(i1, 1);
  for (int it = 0; i0.r < M; ++it) {
    const int st = it % kStages;
    issue(i2, (it + 2) % kStages);
    const Idx i3 = load_idx(i2.r + row_stride);
    cp_async_wait<2>();
    const int r = i0.r;
    const int s[4] = {i0.s4.x, i0.s4.y, i0.s4.z, i0.s4.w};
    uint4 mraw = Raw<T>::neg_inf();
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (s[u] >= 0) mraw = Raw<T>::vmax(mraw, *slot(st, 1 + u));
    for (int e = i0.e0 + 4; e < i0.e1; ++e)
      mraw = Raw<T>::vmax(mraw, Raw<T>::load(PQ + (size_t)__ldg(in_src + e) * (2 * D) + D + c));
    float o[V];
    if (i0.e0 == i0.e1) {
      const float z = act_fwd_t<T>(0.f, act);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = z;
    } else {
      float p[V], mx[V];
      Raw<T>::to_float(*slot(st, 0), p);
      Raw<T>::to_float(mraw, mx);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = act_fwd_t<T>(p[i] + mx[i], act);
    }
    apply_dropout<T, V>(o, dr, (uint64_t)r * D + c);
    store_vec(Bout + (size_t)r * D + c, o);
    i0 = i1; i1 = i2; i2 = i3;
  }
  cp_async_wait<0>();
}

}  // namespace v3
}  // namespace gcb
