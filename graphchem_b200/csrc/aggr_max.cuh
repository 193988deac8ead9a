// GeneralConv with aggr="max" (graphchem/nn/gcn.py:84-86 passes `aggr` straight to PyG's GeneralConv;
// PyG 2.6.1 nn/conv/general_conv.py + nn/aggr/basic.py MaxAggregation -> scatter(reduce='max'), i.e.
// zeros.scatter_reduce_(amax, include_self=False): rows without in-edges stay 0, ties split evenly in backward).
//
// The maximum does not commute with the linear maps, so this path is transform-then-aggregate, one row per EDGE
// in dst-CSR order p (position in in_ptr / in_src / in_eid; a node's in-edges are a contiguous row range):
//
//   forward   Z[p]    = [ A[in_src p] | Bnew[in_eid p] ]                           edge_rows_kernel        [E, 2D]
//             Msg     = Z [W_m | W_e]^T + (b_m + b_e)                              the NT GEMM             [E, D]
//             A'[i]   = act( max_{p in seg i} Msg[p]  (0 when empty)  + A[i] )     atom_max_update_kernel
//   backward  dMsg[p] = dZ[i] [Msg[p] == max_i] / #ties_i                          atom_max_bwd_kernel     [E, D]
//             dW, db  = dMsg^T Z, column sums of dMsg                              the TN GEMM
//             dZrow   = dMsg [W_m | W_e]                                           the NT GEMM             [E, 2D]
//   and the add path's source-side kernels finish the job with dZrow in the place of dS: atom_bwd_gather through
//   out_inpos (in-CSR position of every out-edge), bond_bwd_tie through `einpos` (in-CSR position of edge r).
// One thread owns one 16-byte channel chunk of one row (the mapping of kernels.cuh); any width that is a multiple of
// the vector length, both dtypes.  Correctness path: the tcgen05 / tile kernels serve add aggregation only.
#pragma once
#include "kernels.cuh"

namespace gcb {

// einpos[in_eid[p]] = p: where edge e sits in the dst-sorted CSR
__global__ void edge_inpos_kernel(const int32_t* __restrict__ in_eid, int32_t* __restrict__ einpos, int E) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < E) einpos[__ldg(in_eid + p)] = p;
}

template <typename T>
__global__ void edge_rows_kernel(const T* __restrict__ A, const T* __restrict__ Bnew, const int32_t* __restrict__ in_src,
                                 const int32_t* __restrict__ in_eid, T* __restrict__ Z, int E, int M, int D, int act,
                                 Drop dr_bond) {
  GCB_ROW_CHUNK_PROLOGUE(E)
  const int s = __ldg(in_src + r), eid = __ldg(in_eid + r);
  float a[V], b[V];
  load_vec(A + (size_t)s * D + c, a);
  if (eid < M) {
    load_vec(Bnew + (size_t)eid * D + c, b);
  } else {  // bond rows >= min(N, E) never change: act(0), dropped out like the stored rows
    const float cst = round_to(act_fwd_t<T>(0.f, act), (const T*)nullptr);
#pragma unroll
    for (int i = 0; i < V; ++i) b[i] = cst;
    apply_dropout<T, V>(b, dr_bond, (uint64_t)eid * D + c);
  }
  store_vec(Z + (size_t)r * 2 * D + c, a);
  store_vec(Z + (size_t)r * 2 * D + D + c, b);
}

template <typename T>
__global__ void atom_max_update_kernel(const T* __restrict__ Msg, const int32_t* __restrict__ in_ptr,
                                       const T* __restrict__ A_prev, T* __restrict__ A_out, int N, int D, int act) {
  GCB_ROW_CHUNK_PROLOGUE(N)
  const int e0 = in_ptr[r], e1 = in_ptr[r + 1];
  float mx[V], o[V];
#pragma unroll
  for (int i = 0; i < V; ++i) mx[i] = e1 > e0 ? -CUDART_INF_F : 0.f;
  for (int p = e0; p < e1; ++p) {
    float m[V];
    load_vec(Msg + (size_t)p * D + c, m);
#pragma unroll
    for (int i = 0; i < V; ++i) mx[i] = fmaxf(mx[i], m[i]);
  }
  load_vec(A_prev + (size_t)r * D + c, o);
#pragma unroll
  for (int i = 0; i < V; ++i) o[i] = act_fwd_t<T>(mx[i] + o[i], act);
  store_vec(A_out + (size_t)r * D + c, o);
}

template <typename T>
__global__ void atom_max_bwd_kernel(const T* __restrict__ Msg, const int32_t* __restrict__ in_ptr,
                                    const T* __restrict__ dZA, T* __restrict__ dMsg, int N, int D) {
  GCB_ROW_CHUNK_PROLOGUE(N)
  const int e0 = in_ptr[r], e1 = in_ptr[r + 1];
  if (e1 == e0) return;
  float mx[V], cnt[V], d[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { mx[i] = -CUDART_INF_F; cnt[i] = 0.f; }
  for (int p = e0; p < e1; ++p) {
    float m[V];
    load_vec(Msg + (size_t)p * D + c, m);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      if (m[i] > mx[i]) { mx[i] = m[i]; cnt[i] = 1.f; }
      else if (m[i] == mx[i]) cnt[i] += 1.f;
    }
  }
  load_vec(dZA + (size_t)r * D + c, d);
#pragma unroll
  for (int i = 0; i < V; ++i) d[i] = d[i] / cnt[i];
  for (int p = e0; p < e1; ++p) {
    float m[V], g[V];
    load_vec(Msg + (size_t)p * D + c, m);
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = m[i] == mx[i] ? d[i] : 0.f;
    store_vec(dMsg + (size_t)p * D + c, g);
  }
}

template <typename T>
int launch_edge_inpos(const int32_t* in_eid, int32_t* einpos, int E, cudaStream_t s) {
  if (E <= 0) return GCB_OK;
  GCB_KERNEL("edge_inpos_kernel", s, edge_inpos_kernel<<<(unsigned)ceil_div(E, 256), 256, 0, s>>>(in_eid, einpos, E));
  return GCB_OK;
}
template <typename T>
int launch_edge_rows(const T* A, const T* Bnew, const int32_t* in_src, const int32_t* in_eid, T* Z, int E, int M, int D,
                     int act, Drop dr_bond, cudaStream_t s) {
  if (E <= 0) return GCB_OK;
  GCB_KERNEL("edge_rows_kernel", s, edge_rows_kernel<T><<<row_chunk_blocks(E, D, Vec<T>::N), 256, 0, s>>>(A, Bnew, in_src, in_eid, Z, E, M, D, act, dr_bond));
  return GCB_OK;
}
template <typename T>
int launch_atom_max_update(const T* Msg, const int32_t* in_ptr, const T* A_prev, T* A_out, int N, int D, int act, cudaStream_t s) {
  if (N <= 0) return GCB_OK;
  GCB_KERNEL("atom_max_update_kernel", s, atom_max_update_kernel<T><<<row_chunk_blocks(N, D, Vec<T>::N), 256, 0, s>>>(Msg, in_ptr, A_prev, A_out, N, D, act));
  return GCB_OK;
}
template <typename T>
int launch_atom_max_bwd(const T* Msg, const int32_t* in_ptr, const T* dZA, T* dMsg, int N, int D, cudaStream_t s) {
  if (N <= 0) return GCB_OK;
  GCB_KERNEL("atom_max_bwd_kernel", s, atom_max_bwd_kernel<T><<<row_chunk_blocks(N, D, Vec<T>::N), 256, 0, s>>>(Msg, in_ptr, dZA, dMsg, N, D));
  return GCB_OK;
}

}  // namespace gcb
