// Second-generation row-wise kernels: same arithmetic and the same summation order as kernels.cuh
// (so results are bit-identical to the generic path), restructured for memory-level parallelism:
//
//   * a row is owned by a GROUP of CH = D / (16 bytes) lanes (8, 16 or 32 -> 4, 2 or 1 rows per warp);
//   * the group loads the row's CSR segment cooperatively (one coalesced index load, then
//     __shfl_sync broadcasts) instead of one dependent scalar load per edge per lane;
//   * neighbour-row loads are issued four at a time before they are consumed;
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

#define GCB_V2_PROLOGUE(ROWS)                                                                        \
  constexpr int V = Vec<T>::N;                                                                       \
  constexpr int D = CH * V;                                                                          \
  constexpr int RPW = 32 / CH;                                                                       \
  const int lane = threadIdx.x & 31;                                                                 \
  const int lig = lane & (CH - 1);                                                                   \
  const int gig = lane / CH;                                                                         \
  const unsigned gmask = CH == 32 ? 0xffffffffu : (((1u << (CH & 31)) - 1u) << (gig * CH));          \
  const int row_raw = ((int)blockIdx.x * (kThreads / 32) + ((int)threadIdx.x >> 5)) * RPW + gig;     \
  const bool row_ok = row_raw < (ROWS);                                                              \
  const int r = row_ok ? row_raw : (ROWS) - 1;                                                       \
  const int c = lig * V;                                                                             \
  (void)D; (void)gmask; (void)c;

template <typename T, int V>
__device__ __forceinline__ void vzero(float (&a)[V]) {
#pragma unroll
  for (int i = 0; i < V; ++i) a[i] = 0.f;
}

// Bond update: Bout[r] = dropout(act(P[r] + max_{e->r} Q[src e])), act(0) for rows without in-edges.
template <typename T, int CH>
__global__ void __launch_bounds__(kThreads)
bond_update_kernel(const T* __restrict__ PQ, const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_src,
                   T* __restrict__ Bout, int M, int act, Drop dr) {
  GCB_V2_PROLOGUE(M)
  const int e0 = __ldg(in_ptr + r), e1 = __ldg(in_ptr + r + 1);
  float o[V];
  if (e0 == e1) {
    const float z = act_fwd_t<T>(0.f, act);
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] = z;
  } else {
    float p[V], mx[V];
    load_vec(PQ + (size_t)r * (2 * D) + c, p);
#pragma unroll
    for (int i = 0; i < V; ++i) mx[i] = -CUDART_INF_F;
    for (int base = e0; base < e1; base += CH) {
      const int cnt = min(CH, e1 - base);
      const int my_src = lig < cnt ? __ldg(in_src + base + lig) : 0;
      for (int j = 0; j < cnt; j += 4) {
        float q[4][V];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (j + u < cnt) {
            const int s = __shfl_sync(gmask, my_src, j + u, CH);
            load_vec(PQ + (size_t)s * (2 * D) + D + c, q[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (j + u < cnt) {
#pragma unroll
            for (int i = 0; i < V; ++i) mx[i] = fmaxf(mx[i], q[u][i]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] = act_fwd_t<T>(p[i] + mx[i], act);
  }
  apply_dropout<T, V>(o, dr, (uint64_t)r * D + c);
  if (row_ok) store_vec(Bout + (size_t)r * D + c, o);
}

// Atom aggregation: S[i] = [ sum_{e->i} A[src e] | sum_{e->i} Bnew[eid e] ]  (Bnew[e >= M] = act(0)).
template <typename T, int CH>
__global__ void __launch_bounds__(kThreads)
atom_gather_kernel(const T* __restrict__ A, const T* __restrict__ Bnew, const int32_t* __restrict__ in_ptr,
                   const int32_t* __restrict__ in_src, const int32_t* __restrict__ in_eid, T* __restrict__ S, int N, int M,
                   int act, int aggr, Drop dr_bond) {
  GCB_V2_PROLOGUE(N)
  const int e0 = __ldg(in_ptr + r), e1 = __ldg(in_ptr + r + 1);
  float sa[V], sb[V];
  vzero<T, V>(sa);
  vzero<T, V>(sb);
  const float cst = round_to(act_fwd_t<T>(0.f, act), (const T*)nullptr);
  for (int base = e0; base < e1; base += CH) {
    const int cnt = min(CH, e1 - base);
    const int my_src = lig < cnt ? __ldg(in_src + base + lig) : 0;
    const int my_eid = lig < cnt ? __ldg(in_eid + base + lig) : 0;
    for (int j = 0; j < cnt; j += 4) {
      float a[4][V], b[4][V];
      int eid[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j + u < cnt) {
          const int s = __shfl_sync(gmask, my_src, j + u, CH);
          eid[u] = __shfl_sync(gmask, my_eid, j + u, CH);
          load_vec(A + (size_t)s * D + c, a[u]);
          if (eid[u] < M) load_vec(Bnew + (size_t)eid[u] * D + c, b[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j + u < cnt) {
          if (eid[u] >= M) {
#pragma unroll
            for (int i = 0; i < V; ++i) b[u][i] = cst;
            apply_dropout<T, V>(b[u], dr_bond, (uint64_t)eid[u] * D + c);
#pragma unroll
            for (int i = 0; i < V; ++i) b[u][i] = round_to(b[u][i], (const T*)nullptr);
          }
#pragma unroll
          for (int i = 0; i < V; ++i) { sa[i] += a[u][i]; sb[i] += b[u][i]; }
        }
      }
    }
  }
  if (aggr == GCB_AGGR_MEAN && e1 > e0) {
    const float inv = 1.f / (float)(e1 - e0);
#pragma unroll
    for (int i = 0; i < V; ++i) { sa[i] *= inv; sb[i] *= inv; }
  }
  if (row_ok) {
    store_vec(S + (size_t)r * (2 * D) + c, sa);
    store_vec(S + (size_t)r * (2 * D) + D + c, sb);
  }
}

// dZA_prev[j] = ( dZA[j] + sum_{e: src e = j} w(dst e) dS[dst e, 0:D] ) * dropfactor * act'(A_prev[j])
template <typename T, int CH>
__global__ void __launch_bounds__(kThreads)
atom_bwd_gather_kernel(const T* __restrict__ dZA, const T* __restrict__ dS, const int32_t* __restrict__ out_ptr,
                       const int32_t* __restrict__ out_dst, const int32_t* __restrict__ in_ptr,
                       const T* __restrict__ A_prev, T* __restrict__ dZA_prev, int N, int act, int aggr, Drop dr_prev) {
  GCB_V2_PROLOGUE(N)
  const int e0 = __ldg(out_ptr + r), e1 = __ldg(out_ptr + r + 1);
  float d[V], y[V];
  load_vec(dZA + (size_t)r * D + c, d);
  load_vec(A_prev + (size_t)r * D + c, y);
  for (int base = e0; base < e1; base += CH) {
    const int cnt = min(CH, e1 - base);
    int my_t = 0;
    float my_w = 1.f;
    if (lig < cnt) {
      my_t = __ldg(out_dst + base + lig);
      if (aggr == GCB_AGGR_MEAN) my_w = 1.f / (float)max(__ldg(in_ptr + my_t + 1) - __ldg(in_ptr + my_t), 1);
    }
    for (int j = 0; j < cnt; j += 4) {
      float g[4][V], w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j + u < cnt) {
          const int t = __shfl_sync(gmask, my_t, j + u, CH);
          w[u] = __shfl_sync(gmask, my_w, j + u, CH);
          load_vec(dS + (size_t)t * (2 * D) + c, g[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j + u < cnt) {
#pragma unroll
          for (int i = 0; i < V; ++i) d[i] = fmaf(w[u], g[u][i], d[i]);
        }
      }
    }
  }
  if (dr_prev.p > 0.f) {
    mul_act_grad<V>(d, y, act, dr_prev, (uint64_t)r * D + c);
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) d[i] *= act_grad_t<T>(y[i], act);
  }
  if (row_ok) store_vec(dZA_prev + (size_t)r * D + c, d);
}

// Destination side of the bond backward (see kernels.cuh::bond_bwd_tie_kernel for the math).
template <typename T, int CH>
__global__ void __launch_bounds__(kThreads)
bond_bwd_tie_kernel(const T* __restrict__ dS, const T* __restrict__ dBnext, const T* __restrict__ d_out_bond,
                    const T* __restrict__ B_l, const T* __restrict__ PQ, const int32_t* __restrict__ dst,
                    const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_src, T* __restrict__ dPQ,
                    T* __restrict__ mx_out, T* __restrict__ gsplit, int M, int act, int aggr, Drop dr) {
  GCB_V2_PROLOGUE(M)
  const int e0 = __ldg(in_ptr + r), e1 = __ldg(in_ptr + r + 1);
  float d[V], mx[V], cnt_t[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { d[i] = 0.f; mx[i] = 0.f; cnt_t[i] = 1.f; }
  if (e1 > e0) {
    const int t = __ldg(dst + r);
    float y[V];
    load_vec(dS + (size_t)t * (2 * D) + D + c, d);
    load_vec(B_l + (size_t)r * D + c, y);
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
    if (dr.p > 0.f) {
      mul_act_grad<V>(d, y, act, dr, (uint64_t)r * D + c);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) d[i] *= act_grad_t<T>(y[i], act);
    }
#pragma unroll
    for (int i = 0; i < V; ++i) { mx[i] = -CUDART_INF_F; cnt_t[i] = 0.f; }
    for (int base = e0; base < e1; base += CH) {
      const int cnt = min(CH, e1 - base);
      const int my_src = lig < cnt ? __ldg(in_src + base + lig) : 0;
      for (int j = 0; j < cnt; j += 4) {
        float q[4][V];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (j + u < cnt) {
            const int s = __shfl_sync(gmask, my_src, j + u, CH);
            load_vec(PQ + (size_t)s * (2 * D) + D + c, q[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (j + u < cnt) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
              if (q[u][i] > mx[i]) { mx[i] = q[u][i]; cnt_t[i] = 1.f; }
              else if (q[u][i] == mx[i]) cnt_t[i] += 1.f;
            }
          }
        }
      }
    }
  }
  float g[V];
#pragma unroll
  for (int i = 0; i < V; ++i) g[i] = cnt_t[i] == 1.f ? d[i] : d[i] / cnt_t[i];
  if (row_ok) {
    store_vec(dPQ + (size_t)r * (2 * D) + c, d);
    store_vec(mx_out + (size_t)r * D + c, mx);
    store_vec(gsplit + (size_t)r * D + c, g);
  }
}

// Source side: dPQ[j, D:2D] = sum_{e: src e = j} [Q[j] == mx[dst e]] * gsplit[dst e]
template <typename T, int CH>
__global__ void __launch_bounds__(kThreads)
bond_bwd_scatter_kernel(const T* __restrict__ PQ, const T* __restrict__ mx, const T* __restrict__ gsplit,
                        const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_dst, T* __restrict__ dPQ,
                        int M) {
  GCB_V2_PROLOGUE(M)
  const int e0 = __ldg(out_ptr + r), e1 = __ldg(out_ptr + r + 1);
  float q[V], acc[V];
  load_vec(PQ + (size_t)r * (2 * D) + D + c, q);
  vzero<T, V>(acc);
  for (int base = e0; base < e1; base += CH) {
    const int cnt = min(CH, e1 - base);
    const int my_t = lig < cnt ? __ldg(out_dst + base + lig) : 0;
    for (int j = 0; j < cnt; j += 4) {
      float m[4][V], g[4][V];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j + u < cnt) {
          const int t = __shfl_sync(gmask, my_t, j + u, CH);
          load_vec(mx + (size_t)t * D + c, m[u]);
          load_vec(gsplit + (size_t)t * D + c, g[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (j + u < cnt) {
#pragma unroll
          for (int i = 0; i < V; ++i) acc[i] += (q[i] == m[u][i]) ? g[u][i] : 0.f;
        }
      }
    }
  }
  if (row_ok) store_vec(dPQ + (size_t)r * (2 * D) + D + c, acc);
}

// dZA_L[n] = (dPool[graph(n)] + d_out_atom[n]) * dropfactor * act'(A_L[n]); one group per NODE
// (node -> graph map from the packed batch).
template <typename T, int CH>
__global__ void __launch_bounds__(kThreads)
pool_bwd_kernel(const float* __restrict__ dPool, const T* __restrict__ d_out_atom, const T* __restrict__ A_last,
                const int32_t* __restrict__ node_graph, T* __restrict__ dZA, int N, int act, Drop dr) {
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
  }
  if (d_out_atom) {
    float u[V];
    load_vec(d_out_atom + (size_t)r * D + c, u);
#pragma unroll
    for (int i = 0; i < V; ++i) d[i] += u[i];
  }
  load_vec(A_last + (size_t)r * D + c, y);
  if (dr.p > 0.f) {
    mul_act_grad<V>(d, y, act, dr, (uint64_t)r * D + c);
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) d[i] *= act_grad_t<T>(y[i], act);
  }
  if (row_ok) store_vec(dZA + (size_t)r * D + c, d);
}

// Embedding gradient: one CTA per (token, split); 256 threads = (256/CH) row lanes x CH chunks.  Row
// lane k sums rows k, k+RL, k+2RL, ... of its slice, then the row lanes are combined in lane order.
template <typename T, int CH>
__global__ void __launch_bounds__(kThreads)
embed_bwd_partial_kernel(const T* __restrict__ dZ, const T* __restrict__ Y, const int32_t* __restrict__ tok_ptr,
                         const int32_t* __restrict__ tok_row, float* __restrict__ partial, int vocab, int splits, int act) {
  constexpr int V = Vec<T>::N;
  constexpr int D = CH * V;
  constexpr int RL = kThreads / CH;  // row lanes
  __shared__ float red[RL][D + 4];
  const int t = blockIdx.x / splits, s = blockIdx.x % splits;
  const int lig = threadIdx.x % CH, rl = threadIdx.x / CH, c = lig * V;
  const int b0 = __ldg(tok_ptr + t), b1 = __ldg(tok_ptr + t + 1);
  const int per = (b1 - b0 + splits - 1) / splits;
  const int lo = b0 + s * per, hi = min(b1, lo + per);
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  for (int k = lo + rl; k < hi; k += RL) {
    const int row = __ldg(tok_row + k);
    float d[V];
    load_vec(dZ + (size_t)row * D + c, d);
    if (Y) {
      float y[V];
      load_vec(Y + (size_t)row * D + c, y);
#pragma unroll
      for (int i = 0; i < V; ++i) d[i] *= act_grad_t<T>(y[i], act);
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
