// Launch wrappers for the row-wise kernels: pick the group-cooperative kernels of kernels_v2.cuh when
// the row width is 8, 16 or 32 sixteen-byte chunks (D = 64/128/256 in bf16, 32/64/128 in fp32) and the
// generic kernels of kernels.cuh otherwise.
#pragma once
#include <cstdlib>
#include "kernels_v2.cuh"

namespace gcb {

#define GCB_ROW_CASE(NAME, CHVAL, DROPFLAG, V2CALL)                                                 \
  case CHVAL: {                                                                                     \
    constexpr int CH = CHVAL;                                                                       \
    if (DROPFLAG) { constexpr bool DROP = true; GCB_KERNEL(NAME, s, V2CALL); }                      \
    else { constexpr bool DROP = false; GCB_KERNEL(NAME, s, V2CALL); }                              \
    break;                                                                                          \
  }
#define GCB_ROW_DISPATCH(NAME, ROWS, DVAL, DROPFLAG, V2CALL, V1CALL)                                \
  do {                                                                                              \
    if ((ROWS) <= 0) return GCB_OK;                                                                 \
    switch ((DVAL) / Vec<T>::N) {                                                                   \
      GCB_ROW_CASE(NAME, 8, DROPFLAG, V2CALL)                                                       \
      GCB_ROW_CASE(NAME, 16, DROPFLAG, V2CALL)                                                      \
      GCB_ROW_CASE(NAME, 32, DROPFLAG, V2CALL)                                                      \
      default: { GCB_KERNEL(NAME, s, V1CALL); break; }                                              \
    }                                                                                               \
    return GCB_OK;                                                                                  \
  } while (0)

#define GCB_V2_GRID(ROWS) v2::blocks_for<CH>(ROWS), v2::kThreads, 0, s
#define GCB_V2_GRID_FLAT(ROWS) v2::blocks_for<CH>(ROWS), v2::kThreads, 0, s
#define GCB_V1_GRID(ROWS, DVAL) row_chunk_blocks(ROWS, DVAL, Vec<T>::N), 256, 0, s

template <typename T>
int launch_bond_update(const T* PQ, const int32_t* in_ptr, const int32_t* in_src, const int32_t* in_eid,
                       const int32_t* in_src4, const int32_t* in_eid4, T* Bout, uint8_t* tiemask, uint8_t* tiecnt, int M,
                       int D, int act, Drop dr, int rev, cudaStream_t s) {
  if (tiemask != nullptr) {
    GCB_ROW_DISPATCH("bond_update_kernel", M, D, dr.p > 0.f,
                     (v2::bond_update_kernel<T, CH, DROP, true><<<GCB_V2_GRID(M)>>>(PQ, in_ptr, in_src, in_eid, (const int4*)in_src4, (const int4*)in_eid4, Bout, tiemask, tiecnt, M, act, dr, rev)),
                     (bond_update_kernel<T><<<GCB_V1_GRID(M, D)>>>(PQ, in_ptr, in_src, Bout, M, D, act, dr)));
  }
  GCB_ROW_DISPATCH("bond_update_kernel", M, D, dr.p > 0.f,
                   (v2::bond_update_kernel<T, CH, DROP, false><<<GCB_V2_GRID(M)>>>(PQ, in_ptr, in_src, in_eid, (const int4*)in_src4, (const int4*)in_eid4, Bout, tiemask, tiecnt, M, act, dr, rev)),
                   (bond_update_kernel<T><<<GCB_V1_GRID(M, D)>>>(PQ, in_ptr, in_src, Bout, M, D, act, dr)));
}

template <typename T>
int launch_atom_gather(const T* A, const T* Bnew, const int32_t* in_ptr, const int32_t* in_src, const int32_t* in_eid,
                       const int32_t* in_src4, const int32_t* in_eid4, T* S, int N, int M, int D, int act, int aggr, Drop dr_bond, int rev, cudaStream_t s) {
  const bool generic = dr_bond.p > 0.f || aggr == GCB_AGGR_MEAN;
  const int ch = D / Vec<T>::N;
  // lean path: one single-stream pass per half -- for batches large enough to be bandwidth-bound; small batches are
  // launch-latency-bound and take the one-launch kernel below (same arithmetic and order)
  if (!generic && N >= 2048 && (ch == 8 || ch == 16 || ch == 32)) {
    switch (ch) {
      case 8:
        GCB_KERNEL("atom_gather_kernel", s, (v2::atom_gather_kernel<T, 8, false, 0><<<v2::blocks_for<8>(N), v2::kThreads, 0, s>>>(A, Bnew, in_ptr, in_src, in_eid, (const int4*)in_src4, (const int4*)in_eid4, S, N, M, act, aggr, dr_bond, 0)));
        GCB_KERNEL("atom_gather_kernel", s, (v2::atom_gather_kernel<T, 8, false, 1><<<v2::blocks_for<8>(N), v2::kThreads, 0, s>>>(A, Bnew, in_ptr, in_src, in_eid, (const int4*)in_src4, (const int4*)in_eid4, S, N, M, act, aggr, dr_bond, 0)));
        break;
      case 16:
        GCB_KERNEL("atom_gather_kernel", s, (v2::atom_gather_kernel<T, 16, false, 0><<<v2::blocks_for<16>(N), v2::kThreads, 0, s>>>(A, Bnew, in_ptr, in_src, in_eid, (const int4*)in_src4, (const int4*)in_eid4, S, N, M, act, aggr, dr_bond, 0)));
        GCB_KERNEL("atom_gather_kernel", s, (v2::atom_gather_kernel<T, 16, false, 1><<<v2::blocks_for<16>(N), v2::kThreads, 0, s>>>(A, Bnew, in_ptr, in_src, in_eid, (const int4*)in_src4, (const int4*)in_eid4, S, N, M, act, aggr, dr_bond, 0)));
        break;
      default:
        GCB_KERNEL("atom_gather_kernel", s, (v2::atom_gather_kernel<T, 32, false, 0><<<v2::blocks_for<32>(N), v2::kThreads, 0, s>>>(A, Bnew, in_ptr, in_src, in_eid, (const int4*)in_src4, (const int4*)in_eid4, S, N, M, act, aggr, dr_bond, 0)));
        GCB_KERNEL("atom_gather_kernel", s, (v2::atom_gather_kernel<T, 32, false, 1><<<v2::blocks_for<32>(N), v2::kThreads, 0, s>>>(A, Bnew, in_ptr, in_src, in_eid, (const int4*)in_src4, (const int4*)in_eid4, S, N, M, act, aggr, dr_bond, 0)));
    }
    return GCB_OK;
  }
  GCB_ROW_DISPATCH("atom_gather_kernel", N, D, generic,
                   (v2::atom_gather_kernel<T, CH, DROP, 2><<<GCB_V2_GRID(N)>>>(A, Bnew, in_ptr, in_src, in_eid, (const int4*)in_src4, (const int4*)in_eid4, S, N, M, act, aggr, dr_bond, rev)),
                   (atom_gather_kernel<T><<<GCB_V1_GRID(N, D)>>>(A, Bnew, in_ptr, in_src, in_eid, S, N, M, D, act, aggr, dr_bond)));
}

template <typename T>
int launch_atom_bwd_gather(const T* dZA, const T* dS, const int32_t* out_ptr, const int32_t* out_dst,
                           const int32_t* out_dst4, const int32_t* in_ptr, const T* A_prev, T* dZA_prev, int N, int D, int act, int aggr,
                           Drop dr_prev, int rev, cudaStream_t s) {
  GCB_ROW_DISPATCH("atom_bwd_gather_kernel", N, D, dr_prev.p > 0.f || aggr == GCB_AGGR_MEAN,
                   (v2::atom_bwd_gather_kernel<T, CH, DROP><<<GCB_V2_GRID(N)>>>(dZA, dS, out_ptr, out_dst, (const int4*)out_dst4, in_ptr, A_prev, dZA_prev, N, act, aggr, dr_prev, rev)),
                   (atom_bwd_gather_kernel<T><<<GCB_V1_GRID(N, D)>>>(dZA, dS, out_ptr, out_dst, in_ptr, A_prev, dZA_prev, N, D, act, aggr, dr_prev)));
}

template <typename T>
int launch_bond_bwd_tie(const T* dS, const T* dBnext, const T* d_out_bond, const T* B_l, const T* PQ, const int32_t* dst,
                        const int32_t* in_ptr, const int32_t* in_src, const int32_t* in_eid, const int32_t* in_src4,
                        const int32_t* in_eid4, T* dPQ, T* mx, T* gsplit, uint8_t* tiemask, const uint8_t* tiecnt, int M, int D, int act,
                        int aggr, Drop dr, int rev, cudaStream_t s) {
  GCB_ROW_DISPATCH("bond_bwd_tie_kernel", M, D, dr.p > 0.f || aggr == GCB_AGGR_MEAN,
                   (v2::bond_bwd_tie_kernel<T, CH, DROP><<<GCB_V2_GRID(M)>>>(dS, dBnext, d_out_bond, B_l, dst, in_ptr, tiecnt, dPQ, gsplit, M, act, aggr, dr, rev)),
                   (bond_bwd_tie_kernel<T><<<GCB_V1_GRID(M, D)>>>(dS, dBnext, d_out_bond, B_l, PQ, dst, in_ptr, in_src, dPQ, mx, gsplit, M, D, act, aggr, dr)));
}

template <typename T>
int launch_bond_bwd_scatter(const T* PQ, const T* mx, const T* gsplit, const int32_t* out_ptr, const int32_t* out_dst,
                            const int32_t* out_eid, const int32_t* out_dst4, const int32_t* out_eid4,
                            const uint8_t* tiemask, T* dPQ, int M, int D, int rev, cudaStream_t s) {
  GCB_ROW_DISPATCH("bond_bwd_scatter_kernel", M, D, false,
                   (v2::bond_bwd_scatter_kernel<T, CH, DROP><<<GCB_V2_GRID(M)>>>(gsplit, tiemask, out_ptr, out_dst, out_eid, (const int4*)out_dst4, (const int4*)out_eid4, dPQ, M, rev)),
                   (bond_bwd_scatter_kernel<T><<<GCB_V1_GRID(M, D)>>>(PQ, mx, gsplit, out_ptr, out_dst, dPQ, M, D)));
}

// Pool backward: node-parallel (v2, via node_graph) or graph-parallel (generic).
template <typename T>
int launch_pool_bwd(const float* dPool, const T* d_out_atom, const T* A_last, const int32_t* graph_ptr,
                    const int32_t* graph_node, const int32_t* node_graph, T* dZA, int N, int G, int D, int act, Drop dr,
                    int mean, cudaStream_t s) {
  GCB_ROW_DISPATCH("pool_bwd_kernel", N, D, dr.p > 0.f,
                   (v2::pool_bwd_kernel<T, CH, DROP><<<GCB_V2_GRID_FLAT(N)>>>(dPool, d_out_atom, A_last, node_graph, dZA, N, act, dr, 0, mean ? graph_ptr : nullptr)),
                   (pool_bwd_kernel<T><<<GCB_V1_GRID(G, D)>>>(dPool, d_out_atom, A_last, graph_ptr, graph_node, dZA, G, D, act, dr, mean)));
}

constexpr int kEmbedSplits = 32;  // capacity of the partial-sum scratch
// splits actually used (A/B switch GCB_EMBED_SPLITS, 1..kEmbedSplits)
inline int embed_splits() {
  static const int v = [] {
    const char* e = std::getenv("GCB_EMBED_SPLITS");
    const int n = e ? std::atoi(e) : 16;  // 16 measured best on B200 (gpurun_out/r2_ab_q.jsonl)
    return n < 1 ? 1 : (n > kEmbedSplits ? kEmbedSplits : n);
  }();
  return v;
}

// partial[split][token][D]; the caller reduces over the kEmbedSplits splits in order.
template <typename T>
int launch_embed_bwd_partial(const T* dZ, const T* Y, const int32_t* tok_ptr, const int32_t* tok_row, float* partial,
                             int vocab, int D, int act, cudaStream_t s) {
  if (vocab <= 0) return GCB_OK;
  constexpr int V = Vec<T>::N;
  switch (D / V) {
    case 8:
      if (Y) GCB_KERNEL("embed_bwd_partial_kernel", s, (v2::embed_bwd_partial_kernel<T, 8, true><<<vocab * embed_splits(), v2::kThreads, 0, s>>>(dZ, Y, tok_ptr, tok_row, partial, vocab, embed_splits(), act)));
      else GCB_KERNEL("embed_bwd_partial_kernel", s, (v2::embed_bwd_partial_kernel<T, 8, false><<<vocab * embed_splits(), v2::kThreads, 0, s>>>(dZ, Y, tok_ptr, tok_row, partial, vocab, embed_splits(), act)));
      break;
    case 16:
      if (Y) GCB_KERNEL("embed_bwd_partial_kernel", s, (v2::embed_bwd_partial_kernel<T, 16, true><<<vocab * embed_splits(), v2::kThreads, 0, s>>>(dZ, Y, tok_ptr, tok_row, partial, vocab, embed_splits(), act)));
      else GCB_KERNEL("embed_bwd_partial_kernel", s, (v2::embed_bwd_partial_kernel<T, 16, false><<<vocab * embed_splits(), v2::kThreads, 0, s>>>(dZ, Y, tok_ptr, tok_row, partial, vocab, embed_splits(), act)));
      break;
    case 32:
      if (Y) GCB_KERNEL("embed_bwd_partial_kernel", s, (v2::embed_bwd_partial_kernel<T, 32, true><<<vocab * embed_splits(), v2::kThreads, 0, s>>>(dZ, Y, tok_ptr, tok_row, partial, vocab, embed_splits(), act)));
      else GCB_KERNEL("embed_bwd_partial_kernel", s, (v2::embed_bwd_partial_kernel<T, 32, false><<<vocab * embed_splits(), v2::kThreads, 0, s>>>(dZ, Y, tok_ptr, tok_row, partial, vocab, embed_splits(), act)));
      break;
    default:
      GCB_KERNEL("embed_bwd_partial_kernel", s, (embed_bwd_partial_kernel<T><<<(unsigned)ceil_div((int64_t)vocab * embed_splits() * (D / V), 256), 256, 0, s>>>(dZ, Y, tok_ptr, tok_row, partial, vocab, embed_splits(), D, act)));
  }
  return GCB_OK;
}

}  // namespace gcb
