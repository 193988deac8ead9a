// Row-wise kernels of the MoleculeGCN hot path: embedding, destination-segmented gather/reduce
// (forward), source-segmented gather/reduce (backward), pooling, loss, optimiser.
//
// Mapping used throughout: one thread owns one 16-byte channel chunk (4 fp32 / 8 bf16) of one
// row and walks that row's CSR segment sequentially.  Consecutive threads own consecutive chunks
// of the same row, so every neighbour row is fetched with coalesced 128-bit loads, every
// reduction runs in CSR order (= the reference's edge-id scatter order) inside one thread, and no
// atomics or cross-thread reductions are needed: results are deterministic.
//
// The arithmetic restated here is SURVEY.md 3.3 (PyG 2.6.1 EdgeConv / GeneralConv /
// global_add_pool as called from graphchem/nn/gcn.py:152-199).
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace gcb {

struct Drop {  // dropout descriptor for one tensor (p == 0 -> inactive)
  float p = 0.f;
  uint64_t seed = 0;
  uint32_t tag = 0;
};

#define GCB_ROW_CHUNK_PROLOGUE(ROWS)                                                  \
  constexpr int V = Vec<T>::N;                                                        \
  const int CH = D / V;                                                               \
  const int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;                \
  if (gidx >= (int64_t)(ROWS) * CH) return;                                           \
  const int r = (int)(gidx / CH);                                                     \
  const int c = (int)(gidx % CH) * V;

inline unsigned row_chunk_blocks(int64_t rows, int D, int vec) { return (unsigned)ceil_div(rows * (D / vec), 256); }

// out[r,:] = tab[tok[r],:]: the embedding lookup + activation of gcn.py:152-157 as a row copy out of the
// pre-activated table (prepare_weights_kernel); one 16-byte chunk per thread, 32-bit index math.
template <typename T>
__global__ void embed_rows_kernel(const T* __restrict__ tab, const int32_t* __restrict__ tok, T* __restrict__ out,
                                  int rows, int CH) {
  const unsigned gidx = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned r = gidx / (unsigned)CH, ch = gidx - r * (unsigned)CH;
  if (r >= (unsigned)rows) return;
  const uint4* src = reinterpret_cast<const uint4*>(tab) + (size_t)__ldg(tok + r) * CH + ch;
  reinterpret_cast<uint4*>(out)[(size_t)r * CH + ch] = __ldg(src);
}

template <typename T, int V>
__device__ __forceinline__ void apply_dropout(float (&v)[V], const Drop& dr, uint64_t base_idx) {
  if (dr.p > 0.f) {
    const float sc = 1.f / (1.f - dr.p);
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = dropout_keep(dr.seed, dr.tag, base_idx + i, dr.p) ? v[i] * sc : 0.f;
  }
}

// d * dropout-factor * act'(y_pre) where y is the stored (post-dropout) activation.
template <int V>
__device__ __forceinline__ void mul_act_grad(float (&d)[V], const float (&y)[V], int act, const Drop& dr,
                                             uint64_t base_idx) {
  if (dr.p > 0.f) {
    const float sc = 1.f / (1.f - dr.p), un = 1.f - dr.p;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const bool keep = dropout_keep(dr.seed, dr.tag, base_idx + i, dr.p);
      d[i] = keep ? d[i] * sc * act_grad_from_out(y[i] * un, act) : 0.f;
    }
  } else {
    act_grad_mul_vec_impl<false, V>(d, y, act);
  }
}

// Bond update (EdgeConv on the bond matrix, gcn.py:163-169):
//   Bout[r,:] = dropout(act( P[r,:] + max_{e: dst e = r} Q[src e,:] ))   rows with no in-edge: act(0)
// PQ is [M, 2D] = [P | Q] with P = B (W1-W2)^T + b, Q = B W2^T.
template <typename T>
__global__ void bond_update_kernel(const T* __restrict__ PQ, const int32_t* __restrict__ in_ptr,
                                   const int32_t* __restrict__ in_src, T* __restrict__ Bout, int M, int D, int act,
                                   Drop dr) {
  GCB_ROW_CHUNK_PROLOGUE(M)
  const int e0 = in_ptr[r], e1 = in_ptr[r + 1];
  float o[V];
  if (e0 == e1) {
    const float z = act_fwd_t<T>(0.f, act);
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] = z;
  } else {
    float p[V], mx[V];
    load_vec(PQ + (size_t)r * 2 * D + c, p);
#pragma unroll
    for (int i = 0; i < V; ++i) mx[i] = -CUDART_INF_F;
    for (int e = e0; e < e1; ++e) {
      const int s = __ldg(in_src + e);
      float q[V];
      load_vec(PQ + (size_t)s * 2 * D + D + c, q);
#pragma unroll
      for (int i = 0; i < V; ++i) mx[i] = fmaxf(mx[i], q[i]);
    }
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] = act_fwd_t<T>(p[i] + mx[i], act);
  }
  apply_dropout<T, V>(o, dr, (uint64_t)r * D + c);
  store_vec(Bout + (size_t)r * D + c, o);
}

// Atom aggregation (the gather half of GeneralConv, gcn.py:172):
//   S[i, 0:D]  = sum_{e: dst e = i} A[src e,:]
//   S[i, D:2D] = sum_{e: dst e = i} Bnew[e,:]      Bnew[e] = act(0) (dropped out) for e >= M
// (aggr == MEAN divides both by max(deg,1)).
template <typename T>
__global__ void atom_gather_kernel(const T* __restrict__ A, const T* __restrict__ Bnew,
                                   const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_src,
                                   const int32_t* __restrict__ in_eid, T* __restrict__ S, int N, int M, int D,
                                   int act, int aggr, Drop dr_bond) {
  GCB_ROW_CHUNK_PROLOGUE(N)
  const int e0 = in_ptr[r], e1 = in_ptr[r + 1];
  float sa[V], sb[V];
#pragma unroll
  for (int i = 0; i < V; ++i) sa[i] = sb[i] = 0.f;
  const float cst = round_to(act_fwd_t<T>(0.f, act), (const T*)nullptr);
  for (int e = e0; e < e1; ++e) {
    const int s = __ldg(in_src + e), eid = __ldg(in_eid + e);
    float a[V];
    load_vec(A + (size_t)s * D + c, a);
    if (eid < M) {
      float b[V];
      load_vec(Bnew + (size_t)eid * D + c, b);
#pragma unroll
      for (int i = 0; i < V; ++i) { sa[i] += a[i]; sb[i] += b[i]; }
    } else {
      float b[V];
#pragma unroll
      for (int i = 0; i < V; ++i) b[i] = cst;
      apply_dropout<T, V>(b, dr_bond, (uint64_t)eid * D + c);
#pragma unroll
      for (int i = 0; i < V; ++i) { sa[i] += a[i]; sb[i] += round_to(b[i], (const T*)nullptr); }
    }
  }
  if (aggr == GCB_AGGR_MEAN && e1 > e0) {
    const float inv = 1.f / (float)(e1 - e0);
#pragma unroll
    for (int i = 0; i < V; ++i) { sa[i] *= inv; sb[i] *= inv; }
  }
  store_vec(S + (size_t)r * 2 * D + c, sa);
  store_vec(S + (size_t)r * 2 * D + D + c, sb);
}

// x = dropout(x) in place.
template <typename T>
__global__ void dropout_inplace_kernel(T* __restrict__ x, int rows, int D, Drop dr) {
  GCB_ROW_CHUNK_PROLOGUE(rows)
  float v[V];
  load_vec(x + (size_t)r * D + c, v);
  apply_dropout<T, V>(v, dr, (uint64_t)r * D + c);
  store_vec(x + (size_t)r * D + c, v);
}

// fp32 tensors of arbitrary width (readout activations): x[i] = dropout(x[i]), i = flat index.
__global__ void dropout_scalar_kernel(float* __restrict__ x, int64_t n, Drop dr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  x[i] = dropout_keep(dr.seed, dr.tag, (uint64_t)i, dr.p) ? x[i] / (1.f - dr.p) : 0.f;
}

// Constant rows of the returned out_bond: rows [row0, row0+rows) = dropout(act(0)).
template <typename T>
__global__ void fill_const_rows_kernel(T* __restrict__ out, int row0, int rows, int D, int act, Drop dr) {
  GCB_ROW_CHUNK_PROLOGUE(rows)
  float v[V];
  const float z = act_fwd_t<T>(0.f, act);
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = z;
  apply_dropout<T, V>(v, dr, (uint64_t)(row0 + r) * D + c);
  store_vec(out + (size_t)(row0 + r) * D + c, v);
}

// out_bond of a GCB_PACK_EACH batch: out[e,:] = B[pos[e],:], or act(0) where pos[e] < 0.
template <typename T>
__global__ void gather_bond_rows_kernel(const T* __restrict__ B, const int32_t* __restrict__ pos, T* __restrict__ out,
                                        int rows, int D, int act) {
  GCB_ROW_CHUNK_PROLOGUE(rows)
  const int p = __ldg(pos + r);
  float v[V];
  if (p >= 0) {
    load_vec(B + (size_t)p * D + c, v);
  } else {
    const float z = act_fwd_t<T>(0.f, act);
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = z;
  }
  store_vec(out + (size_t)r * D + c, v);
}

// global_add_pool (gcn.py:181): out[g,:] = sum_{i in graph g} A[i,:], in node order.
template <typename T>
__global__ void pool_kernel(const T* __restrict__ A, const int32_t* __restrict__ graph_ptr,
                            const int32_t* __restrict__ graph_node, float* __restrict__ out, int G, int D, int mean) {
  GCB_ROW_CHUNK_PROLOGUE(G)
  const int k0 = graph_ptr[r], k1 = graph_ptr[r + 1];
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  int k = k0;
  for (; k + 4 <= k1; k += 4) {  // four node ids, then four rows in flight; same order of addition
    int n[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) n[u] = __ldg(graph_node + k + u);
    float a[4][V];
#pragma unroll
    for (int u = 0; u < 4; ++u) load_vec(A + (size_t)n[u] * D + c, a[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += a[u][i];
  }
  for (; k < k1; ++k) {
    const int n = __ldg(graph_node + k);
    float a[V];
    load_vec(A + (size_t)n * D + c, a);
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] += a[i];
  }
  if (mean && k1 > k0) {  // global_mean_pool: sum / max(count, 1)
    const float inv = 1.f / (float)(k1 - k0);
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] *= inv;
  }
  float* o = out + (size_t)r * D + c;
#pragma unroll
  for (int i = 0; i < V; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
}

// ---- backward ----------------------------------------------------------------------------------

// dZA_L[n,:] = (dPool[graph(n),:] + d_out_atom[n,:]) * dropfactor * act'(A_L[n,:])
// (with_act == 0: no activation factor -- used when n_messages == 0 ... never: A_0 = act(emb)).
template <typename T>
__global__ void pool_bwd_kernel(const float* __restrict__ dPool, const T* __restrict__ d_out_atom,
                                const T* __restrict__ A_last, const int32_t* __restrict__ graph_ptr,
                                const int32_t* __restrict__ graph_node, T* __restrict__ dZA, int G, int D, int act,
                                Drop dr, int mean) {
  GCB_ROW_CHUNK_PROLOGUE(G)
  const int k0 = graph_ptr[r], k1 = graph_ptr[r + 1];
  float dp[V];
#pragma unroll
  for (int i = 0; i < V; ++i) dp[i] = 0.f;
  if (dPool) {
    const float w = mean && k1 > k0 ? 1.f / (float)(k1 - k0) : 1.f;
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      float4 t = __ldg(reinterpret_cast<const float4*>(dPool + (size_t)r * D + c + i));
      dp[i] = w * t.x; dp[i + 1] = w * t.y; dp[i + 2] = w * t.z; dp[i + 3] = w * t.w;
    }
  }
  for (int k = k0; k < k1; ++k) {
    const int n = __ldg(graph_node + k);
    float d[V], y[V];
#pragma unroll
    for (int i = 0; i < V; ++i) d[i] = dp[i];
    if (d_out_atom) {
      float u[V];
      load_vec(d_out_atom + (size_t)n * D + c, u);
#pragma unroll
      for (int i = 0; i < V; ++i) d[i] += u[i];
    }
    load_vec(A_last + (size_t)n * D + c, y);
    mul_act_grad<V>(d, y, act, dr, (uint64_t)n * D + c);
    store_vec(dZA + (size_t)n * D + c, d);
  }
}

// dZA_prev[j,:] = ( dZA[j,:] + sum_{e: src e = j} w(dst e) * dS[dst e, 0:D] ) * dropfactor * act'(A_prev[j,:])
// w = 1 (add) or 1/max(deg_in,1) (mean).
template <typename T>
__global__ void atom_bwd_gather_kernel(const T* __restrict__ dZA, const T* __restrict__ dS,
                                       const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_dst,
                                       const int32_t* __restrict__ in_ptr, const T* __restrict__ A_prev,
                                       T* __restrict__ dZA_prev, int N, int D, int act, int aggr, Drop dr_prev) {
  GCB_ROW_CHUNK_PROLOGUE(N)
  float d[V];
  load_vec(dZA + (size_t)r * D + c, d);
  const int e0 = out_ptr[r], e1 = out_ptr[r + 1];
  for (int e = e0; e < e1; ++e) {
    const int t = __ldg(out_dst + e);
    float g[V];
    load_vec(dS + (size_t)t * 2 * D + c, g);
    float w = 1.f;
    if (aggr == GCB_AGGR_MEAN) w = 1.f / (float)max(__ldg(in_ptr + t + 1) - __ldg(in_ptr + t), 1);
#pragma unroll
    for (int i = 0; i < V; ++i) d[i] = fmaf(w, g[i], d[i]);
  }
  float y[V];
  load_vec(A_prev + (size_t)r * D + c, y);
  mul_act_grad<V>(d, y, act, dr_prev, (uint64_t)r * D + c);
  store_vec(dZA_prev + (size_t)r * D + c, d);
}

// Backward through the bond activation and the amax aggregation, destination side.
//   dB[r]   = w(dst[r]) * dS[dst[r], D:2D] + dBnext[r] + d_out_bond[r]
//   dzb[r]  = dB[r] * dropfactor * act'(B_l[r])        (0 when row r has no in-edge)
//   mx[r]   = max_{e->r} Q[src e],  cnt[r] = #{e->r : Q[src e] == mx[r]}   (per channel)
//   dPQ[r, 0:D] = dzb[r]   (dP);   gsplit[r] = dzb[r] / cnt[r]   (torch splits ties evenly)
template <typename T>
__global__ void bond_bwd_tie_kernel(const T* __restrict__ dS, const T* __restrict__ dBnext,
                                    const T* __restrict__ d_out_bond, const T* __restrict__ B_l,
                                    const T* __restrict__ PQ, const int32_t* __restrict__ dst,
                                    const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_src,
                                    T* __restrict__ dPQ, T* __restrict__ mx_out, T* __restrict__ gsplit, int M, int D,
                                    int act, int aggr, Drop dr) {
  GCB_ROW_CHUNK_PROLOGUE(M)
  const int e0 = in_ptr[r], e1 = in_ptr[r + 1];
  float d[V], mx[V], cnt[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { d[i] = 0.f; mx[i] = 0.f; cnt[i] = 1.f; }
  if (e1 > e0) {
    const int t = __ldg(dst + r);
    load_vec(dS + (size_t)t * 2 * D + D + c, d);
    if (aggr == GCB_AGGR_MEAN) {
      const float w = 1.f / (float)max(__ldg(in_ptr + t + 1) - __ldg(in_ptr + t), 1);
#pragma unroll
      for (int i = 0; i < V; ++i) d[i] *= w;
    }
    if (dBnext) {
      float u[V];
      load_vec(dBnext + (size_t)r * D + c, u);
#pragma unroll
      for (int i = 0; i < V; ++i) d[i] += u[i];
    }
    if (d_out_bond) {
      float u[V];
      load_vec(d_out_bond + (size_t)r * D + c, u);
#pragma unroll
      for (int i = 0; i < V; ++i) d[i] += u[i];
    }
    float y[V];
    load_vec(B_l + (size_t)r * D + c, y);
    mul_act_grad<V>(d, y, act, dr, (uint64_t)r * D + c);
#pragma unroll
    for (int i = 0; i < V; ++i) { mx[i] = -CUDART_INF_F; cnt[i] = 0.f; }
    for (int e = e0; e < e1; ++e) {
      const int s = __ldg(in_src + e);
      float q[V];
      load_vec(PQ + (size_t)s * 2 * D + D + c, q);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        if (q[i] > mx[i]) { mx[i] = q[i]; cnt[i] = 1.f; }
        else if (q[i] == mx[i]) cnt[i] += 1.f;
      }
    }
  }
  float g[V];
#pragma unroll
  for (int i = 0; i < V; ++i) g[i] = d[i] / cnt[i];
  store_vec(dPQ + (size_t)r * 2 * D + c, d);
  store_vec(mx_out + (size_t)r * D + c, mx);
  store_vec(gsplit + (size_t)r * D + c, g);
}

// Source side of the amax backward:  dPQ[j, D:2D] = dQ[j] = sum_{e: src e = j} [Q[j] == mx[dst e]] * gsplit[dst e]
template <typename T>
__global__ void bond_bwd_scatter_kernel(const T* __restrict__ PQ, const T* __restrict__ mx, const T* __restrict__ gsplit,
                                        const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_dst,
                                        T* __restrict__ dPQ, int M, int D) {
  GCB_ROW_CHUNK_PROLOGUE(M)
  float q[V], acc[V];
  load_vec(PQ + (size_t)r * 2 * D + D + c, q);
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  const int e0 = out_ptr[r], e1 = out_ptr[r + 1];
  for (int e = e0; e < e1; ++e) {
    const int t = __ldg(out_dst + e);
    float m[V], g[V];
    load_vec(mx + (size_t)t * D + c, m);
    load_vec(gsplit + (size_t)t * D + c, g);
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] += (q[i] == m[i]) ? g[i] : 0.f;
  }
  store_vec(dPQ + (size_t)r * 2 * D + D + c, acc);
}

// Embedding gradient (embedding_dense_backward): rows bucketed by token (stable), each bucket cut
// into `splits` contiguous slices; partial[s][t][:] = sum of slice s of token t.  When Y != null
// the incoming gradient is first multiplied by act'(Y[row]) (layer-0 activations, no dropout).
template <typename T>
__global__ void embed_bwd_partial_kernel(const T* __restrict__ dZ, const T* __restrict__ Y,
                                         const int32_t* __restrict__ tok_ptr, const int32_t* __restrict__ tok_row,
                                         float* __restrict__ partial, int vocab, int splits, int D, int act) {
  constexpr int V = Vec<T>::N;
  const int CH = D / V;
  const int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= (int64_t)vocab * splits * CH) return;
  const int c = (int)(gidx % CH) * V;
  const int s = (int)((gidx / CH) % splits);
  const int t = (int)(gidx / ((int64_t)CH * splits));
  const int b0 = tok_ptr[t], b1 = tok_ptr[t + 1];
  const int per = (b1 - b0 + splits - 1) / splits;
  const int lo = b0 + s * per, hi = min(b1, lo + per);
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  for (int k = lo; k < hi; ++k) {
    const int row = __ldg(tok_row + k);
    float d[V];
    load_vec(dZ + (size_t)row * D + c, d);
    if (Y) {
      float y[V];
      load_vec(Y + (size_t)row * D + c, y);
#pragma unroll
      for (int i = 0; i < V; ++i) d[i] *= act_grad_from_out(y[i], act);
    }
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] += d[i];
  }
  float* o = partial + ((size_t)s * vocab + t) * D + c;
#pragma unroll
  for (int i = 0; i < V; ++i) o[i] = acc[i];
}

// ---- loss / optimiser / weight preparation ------------------------------------------------------

// F.mse_loss(pred, y) (mean) and its gradient.  Single CTA, fixed-order tree: deterministic.
__global__ void __launch_bounds__(1024) mse_kernel(const float* __restrict__ pred, const float* __restrict__ y, int64_t n,
                                                   float* __restrict__ loss, float* __restrict__ d_pred,
                                                   float grad_scale) {
  __shared__ float red[1024];
  float acc = 0.f;
  const float inv_n = 1.f / (float)n;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    const float d = pred[i] - y[i];
    acc = fmaf(d, d, acc);
    if (d_pred) d_pred[i] = 2.f * d * inv_n * grad_scale;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0 && loss) loss[0] = red[0] * inv_n;
}

// torch.optim.Adam (no amsgrad, no weight decay): one fused pass over the flat buffers.
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, const float* __restrict__ hyper,
                            const int32_t* __restrict__ step_count, float b1, float b2, float eps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lr = hyper[0], gs = hyper[1];
  const float t = (float)(step_count[0] + 1);
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float grad = g[i] * gs;
  const float mi = b1 * m[i] + (1.f - b1) * grad;
  const float vi = b2 * v[i] + (1.f - b2) * grad * grad;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
  p[i] = p[i] - (lr / bc1) * (mi / denom);
}
__global__ void increment_kernel(int32_t* c) { c[0] += 1; }

}  // namespace gcb
