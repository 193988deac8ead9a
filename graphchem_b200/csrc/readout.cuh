// Fused readout MLP (reference gcn.py:184-199) in fp32: one kernel for the forward chain
// Linear -> act -> dropout ... -> Linear over the pooled graph vectors, one kernel for its backward
// (all weight/bias gradient partials, dPool) plus one ordered reduction.  Replaces three GEMM launches
// forward and ten launches backward; the arithmetic is tiny (<= 0.1 GF at 8192 graphs), so the point is
// launch count and latency, not FLOP/s.  Each CTA owns GB consecutive graphs.
#pragma once
#include <cstdlib>
#include "kernels.cuh"

namespace gcb {

constexpr int kReadoutGB = 32;       // graphs per CTA, at most
// Graphs per CTA for a batch of G graphs: the chain is bound by the warp instructions ONE SM can issue, so a small batch is
// spread over several SMs (every CTA stages the layer's weights again, in parallel): 4 graphs per CTA up to ~600 graphs
// (148 SMs), then more.  A/B: GCB_READOUT_GB_SMALL=32 restores one CTA per 32 graphs.
inline int readout_gb(int G) {
  const char* e = std::getenv("GCB_READOUT_GB_SMALL");
  if (e && std::atoi(e) >= 1 && std::atoi(e) <= kReadoutGB) return std::atoi(e);
  int gb = (G + 147) / 148;
  gb = gb < 4 ? 4 : gb;
  return gb > kReadoutGB ? kReadoutGB : gb;
}
// 1024 threads: a 16-graph batch (the notebooks') is ONE CTA on one SM, 200 k / 510 k warp instructions forward / backward at
// two warps per scheduler (IPC 1.2-1.9, 85 / 142 us under ncu); 32 warps hide the shared-memory and index-math latencies
constexpr int kReadoutThreads = 1024;
constexpr int kReadoutMaxLin = 32;   // Linear layers (n_readout + 1)

struct ReadoutPtrs { float* h[kReadoutMaxLin + 1]; };  // H[0] = pooled, H[k] = input of Linear k

struct ReadoutDesc {
  int n_lin;
  int in_dim[kReadoutMaxLin], out_dim[kReadoutMaxLin];
  int64_t w_off[kReadoutMaxLin], b_off[kReadoutMaxLin];  // offsets into the flat parameter / gradient vector
  int64_t p_off[kReadoutMaxLin];                          // offset of layer k's [W | b] block inside one partial
  int64_t partial_stride;                                 // floats per CTA partial
  int max_dim;                                            // max over all in/out dims
  int max_w;                                              // max over layers of (in + 1) * (out + 1)
};

// H[0] = pooled [G, D]; H[k+1] = dropout(act(H[k] W_k^T + b_k)) for k < n_lin-1; out = H[n_lin-1] W^T + b.
__global__ void __launch_bounds__(kReadoutThreads)
readout_fwd_kernel(ReadoutDesc rd, const float* __restrict__ params, ReadoutPtrs hp, float* __restrict__ out,
                   int G, int act, Drop dr_base, int gb) {
  extern __shared__ float rs[];
  const int ld = rd.max_dim + 1;
  float* cur = rs;
  float* nxt = rs + gb * ld;
  float* Wt = rs + 2 * gb * ld;  // current layer's weights, transposed: Wt[j * (on + 1) + o]
  const int g0 = blockIdx.x * gb;
  const int ng = min(gb, G - g0);
  for (int i = threadIdx.x; i < ng * rd.in_dim[0]; i += kReadoutThreads) {
    const int g = i / rd.in_dim[0], j = i % rd.in_dim[0];
    cur[g * ld + j] = hp.h[0][(size_t)(g0 + g) * rd.in_dim[0] + j];
  }
  __syncthreads();
  for (int k = 0; k < rd.n_lin; ++k) {
    const int in = rd.in_dim[k], on = rd.out_dim[k];
    const bool last = k == rd.n_lin - 1;
    const float* W = params + rd.w_off[k];
    const float* b = params + rd.b_off[k];
    float* dst = last ? out : hp.h[k + 1];
    const int onp = on + 1;
    for (int i = threadIdx.x; i < on * in; i += kReadoutThreads) Wt[(i % in) * onp + (i / in)] = __ldg(W + i);
    __syncthreads();
    for (int i = threadIdx.x; i < ng * on; i += kReadoutThreads) {
      const int g = i / on, o = i % on;
      const float* h = cur + g * ld;
      float a0 = __ldg(b + o), a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int j = 0;
      for (; j + 3 < in; j += 4) {
        a0 = fmaf(h[j], Wt[j * onp + o], a0);
        a1 = fmaf(h[j + 1], Wt[(j + 1) * onp + o], a1);
        a2 = fmaf(h[j + 2], Wt[(j + 2) * onp + o], a2);
        a3 = fmaf(h[j + 3], Wt[(j + 3) * onp + o], a3);
      }
      for (; j < in; ++j) a0 = fmaf(h[j], Wt[j * onp + o], a0);
      float acc = (a0 + a1) + (a2 + a3);
      if (!last) {
        acc = act_fwd(acc, act);
        if (dr_base.p > 0.f) {
          Drop dr = dr_base;
          dr.tag += (uint32_t)(k + 1);
          acc = dropout_keep(dr.seed, dr.tag, (uint64_t)(g0 + g) * on + o, dr.p) ? acc / (1.f - dr.p) : 0.f;
        }
        nxt[g * ld + o] = acc;
      }
      if (dst) dst[(size_t)(g0 + g) * on + o] = acc;
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
}

// Backward: d_out [G, out] -> per-CTA partials of every dW_k / db_k, and dPool [G, D].
__global__ void __launch_bounds__(kReadoutThreads)
readout_bwd_kernel(ReadoutDesc rd, const float* __restrict__ params, ReadoutPtrs hp,
                   const float* __restrict__ d_out, float* __restrict__ partial, float* __restrict__ dPool, int G,
                   int act, Drop dr_base, int gb) {
  extern __shared__ float rs[];
  const int ld = rd.max_dim + 1;
  float* dz = rs;                           // [GB][ld] gradient w.r.t. the current layer's output
  float* hin = rs + gb * ld;        // [GB][ld] the current layer's input H[k]
  float* dh = rs + 2 * gb * ld;     // [GB][ld] gradient w.r.t. the current layer's input
  float* Ws = rs + 3 * gb * ld;     // current layer's weights: Ws[o * (in + 1) + j]
  const int g0 = blockIdx.x * gb;
  const int ng = min(gb, G - g0);
  float* my_partial = partial + (size_t)blockIdx.x * rd.partial_stride;
  const int o_last = rd.out_dim[rd.n_lin - 1];
  for (int i = threadIdx.x; i < ng * o_last; i += kReadoutThreads) {
    const int g = i / o_last, o = i % o_last;
    dz[g * ld + o] = d_out[(size_t)(g0 + g) * o_last + o];
  }
  for (int k = rd.n_lin - 1; k >= 0; --k) {
    const int in = rd.in_dim[k], on = rd.out_dim[k];
    const float* W = params + rd.w_off[k];
    for (int i = threadIdx.x; i < ng * in; i += kReadoutThreads) {
      const int g = i / in, j = i % in;
      hin[g * ld + j] = hp.h[k][(size_t)(g0 + g) * in + j];
    }
    const int inp = in + 1;
    for (int i = threadIdx.x; i < on * in; i += kReadoutThreads) Ws[(i / in) * inp + (i % in)] = __ldg(W + i);
    __syncthreads();
    // weight / bias gradient partials of this CTA's graphs (graph order: deterministic)
    float* pw = my_partial + rd.p_off[k];
    for (int i = threadIdx.x; i < on * in; i += kReadoutThreads) {
      const int o = i / in, j = i % in;
      float acc = 0.f;
      for (int g = 0; g < ng; ++g) acc = fmaf(dz[g * ld + o], hin[g * ld + j], acc);
      pw[i] = acc;
    }
    for (int o = threadIdx.x; o < on; o += kReadoutThreads) {
      float acc = 0.f;
      for (int g = 0; g < ng; ++g) acc += dz[g * ld + o];
      pw[(size_t)on * in + o] = acc;
    }
    // gradient w.r.t. the layer input; through activation + dropout of H[k] when k > 0
    for (int i = threadIdx.x; i < ng * in; i += kReadoutThreads) {
      const int g = i / in, j = i % in;
      float a0 = 0.f, a1 = 0.f;
      int o = 0;
      for (; o + 1 < on; o += 2) {
        a0 = fmaf(dz[g * ld + o], Ws[o * inp + j], a0);
        a1 = fmaf(dz[g * ld + o + 1], Ws[(o + 1) * inp + j], a1);
      }
      for (; o < on; ++o) a0 = fmaf(dz[g * ld + o], Ws[o * inp + j], a0);
      float acc = a0 + a1;
      if (k > 0) {
        float y = hin[g * ld + j], f = 1.f;
        if (dr_base.p > 0.f) {
          Drop dr = dr_base;
          dr.tag += (uint32_t)k;
          f = dropout_keep(dr.seed, dr.tag, (uint64_t)(g0 + g) * in + j, dr.p) ? 1.f / (1.f - dr.p) : 0.f;
          y *= (1.f - dr.p);
        }
        acc *= f * act_grad_from_out(y, act);
        dh[g * ld + j] = acc;
      } else {
        dPool[(size_t)(g0 + g) * in + j] = acc;
      }
    }
    __syncthreads();
    float* t = dz; dz = dh; dh = t;
  }
}


// ------------------------------------------------------------------------------------------------
// Register-tiled variants for large batches (thousands of graphs): 64 graphs per CTA, every thread
// owns a 4 x 4 (forward) or 4 x 8 (backward) output tile and walks the reduction with 16-byte
// shared-memory loads, so the inner loops are FFMA-bound instead of LDS-bound (the kernels above
// issue two shared loads per FFMA).  Weights are kept [out][in + 4] in shared memory: with in % 32 == 0
// the row stride is 4 banks and the 16 output lanes of a warp hit disjoint banks.  Needs every
// layer's input width to be a multiple of 4.
constexpr int kReadoutGB2 = 64;  // graphs per CTA (GB below): 64, or 32 when that puts two CTAs on every SM

__host__ __device__ inline int readout2_ld(const ReadoutDesc& rd) { return rd.max_dim + 4; }
inline size_t readout2_fwd_smem(const ReadoutDesc& rd, int gb = kReadoutGB2) {
  return (size_t)(2 * gb * readout2_ld(rd) + rd.max_dim * readout2_ld(rd)) * sizeof(float);
}
inline size_t readout2_bwd_smem(const ReadoutDesc& rd, int gb = kReadoutGB2) {
  return (size_t)(3 * gb * readout2_ld(rd) + rd.max_dim * readout2_ld(rd)) * sizeof(float);
}

constexpr int kReadout2Threads = 512;  // 16 output lanes x 32 graph lanes

template <int GB>
__global__ void __launch_bounds__(kReadout2Threads)
readout2_fwd_kernel(ReadoutDesc rd, const float* __restrict__ params, ReadoutPtrs hp, float* __restrict__ out, int G,
                    int act, Drop dr_base) {
  extern __shared__ __align__(16) float rs2[];
  const int ld = readout2_ld(rd);
  float* cur = rs2;
  float* nxt = rs2 + GB * ld;
  float* Ws = rs2 + 2 * GB * ld;  // [on][ld]
  const int g0 = blockIdx.x * GB;
  const int ng = min(GB, G - g0);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  {
    const int in4 = rd.in_dim[0] >> 2;
    for (int i = threadIdx.x; i < GB * in4; i += kReadout2Threads) {
      const int g = i / in4, j4 = i % in4;
      const float4 v = g < ng ? __ldg(reinterpret_cast<const float4*>(hp.h[0] + (size_t)(g0 + g) * rd.in_dim[0]) + j4)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(cur + g * ld + 4 * j4) = v;
    }
  }
  for (int k = 0; k < rd.n_lin; ++k) {
    const int in = rd.in_dim[k], on = rd.out_dim[k], in4 = in >> 2;
    const bool last = k == rd.n_lin - 1;
    const float* W = params + rd.w_off[k];
    const float* b = params + rd.b_off[k];
    float* dst = last ? out : hp.h[k + 1];
    for (int i = threadIdx.x; i < on * in4; i += kReadout2Threads) {
      const int o = i / in4, j4 = i % in4;
      *reinterpret_cast<float4*>(Ws + o * ld + 4 * j4) = __ldg(reinterpret_cast<const float4*>(W + (size_t)o * in) + j4);
    }
    __syncthreads();
    for (int ob = 0; ob < on; ob += 64) {
      constexpr int GA = GB / 32;  // graphs per thread: g = ty + 32 a
      float acc[GA][4];
#pragma unroll
      for (int a = 0; a < GA; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
      const float* hrow[GA];
      const float* wrow[4];
#pragma unroll
      for (int a = 0; a < GA; ++a) hrow[a] = cur + (ty + 32 * a) * ld;
#pragma unroll
      for (int c = 0; c < 4; ++c) wrow[c] = Ws + min(ob + tx + 16 * c, on - 1) * ld;
      for (int j4 = 0; j4 < in4; ++j4) {
        float4 h[GA], w[4];
#pragma unroll
        for (int a = 0; a < GA; ++a) h[a] = *reinterpret_cast<const float4*>(hrow[a] + 4 * j4);
#pragma unroll
        for (int c = 0; c < 4; ++c) w[c] = *reinterpret_cast<const float4*>(wrow[c] + 4 * j4);
#pragma unroll
        for (int a = 0; a < GA; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            acc[a][c] = fmaf(h[a].x, w[c].x, acc[a][c]);
            acc[a][c] = fmaf(h[a].y, w[c].y, acc[a][c]);
            acc[a][c] = fmaf(h[a].z, w[c].z, acc[a][c]);
            acc[a][c] = fmaf(h[a].w, w[c].w, acc[a][c]);
          }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int o = ob + tx + 16 * c;
        if (o >= on) continue;
        const float bo = __ldg(b + o);
#pragma unroll
        for (int a = 0; a < GA; ++a) {
          const int g = ty + 32 * a;
          float v = acc[a][c] + bo;
          if (!last) {
            v = act_fwd(v, act);
            if (dr_base.p > 0.f) {
              Drop dr = dr_base;
              dr.tag += (uint32_t)(k + 1);
              v = dropout_keep(dr.seed, dr.tag, (uint64_t)(g0 + g) * on + o, dr.p) ? v / (1.f - dr.p) : 0.f;
            }
            nxt[g * ld + o] = v;
          }
          if (dst && g < ng) dst[(size_t)(g0 + g) * on + o] = v;
        }
      }
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
}

template <int GB>
__global__ void __launch_bounds__(kReadout2Threads)
readout2_bwd_kernel(ReadoutDesc rd, const float* __restrict__ params, ReadoutPtrs hp, const float* __restrict__ d_out,
                    float* __restrict__ partial, float* __restrict__ dPool, int G, int act, Drop dr_base) {
  extern __shared__ __align__(16) float rs2[];
  const int ld = readout2_ld(rd);
  float* dz = rs2;                            // [GB][ld] gradient w.r.t. the current layer's output (0 for padded graphs)
  float* hin = rs2 + GB * ld;        // [GB][ld] the current layer's input H[k]
  float* dh = rs2 + 2 * GB * ld;     // [GB][ld] gradient w.r.t. the current layer's input
  float* Ws = rs2 + 3 * GB * ld;     // [on][ld]
  const int g0 = blockIdx.x * GB;
  const int ng = min(GB, G - g0);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float* my_partial = partial + (size_t)blockIdx.x * rd.partial_stride;
  const int o_last = rd.out_dim[rd.n_lin - 1];
  for (int i = threadIdx.x; i < GB * o_last; i += kReadout2Threads) {
    const int g = i / o_last, o = i % o_last;
    dz[g * ld + o] = g < ng ? d_out[(size_t)(g0 + g) * o_last + o] : 0.f;
  }
  for (int k = rd.n_lin - 1; k >= 0; --k) {
    const int in = rd.in_dim[k], on = rd.out_dim[k], in4 = in >> 2;
    const float* W = params + rd.w_off[k];
    for (int i = threadIdx.x; i < GB * in4; i += kReadout2Threads) {
      const int g = i / in4, j4 = i % in4;
      const float4 v = g < ng ? __ldg(reinterpret_cast<const float4*>(hp.h[k] + (size_t)(g0 + g) * in) + j4)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(hin + g * ld + 4 * j4) = v;
    }
    for (int i = threadIdx.x; i < on * in4; i += kReadout2Threads) {
      const int o = i / in4, j4 = i % in4;
      *reinterpret_cast<float4*>(Ws + o * ld + 4 * j4) = __ldg(reinterpret_cast<const float4*>(W + (size_t)o * in) + j4);
    }
    __syncthreads();
    float* pw = my_partial + rd.p_off[k];
    // ---- weight-gradient partial of this CTA's graphs: outputs o = ob + tx + 16 (2 (ty >> 4) + c), c < 2,
    //      inputs j = jb + 8 (ty & 15) .. + 7
    for (int ob = 0; ob < on; ob += 64) {
      for (int jb = 0; jb < in; jb += 128) {
        constexpr int CW = 2;
        const int j = jb + 8 * (ty & 15), oh = ob + 32 * (ty >> 4);
        if (j >= in) continue;  // in % 8 may be 4: the last lane group handles the 4-wide remainder below
        const bool wide = j + 8 <= in;
        float acc[CW][8];
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[c][e] = 0.f;
        for (int g = 0; g < GB; ++g) {
          const float4 h0 = *reinterpret_cast<const float4*>(hin + g * ld + j);
          const float4 h1 = wide ? *reinterpret_cast<const float4*>(hin + g * ld + j + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
          for (int c = 0; c < CW; ++c) {
            const float d = dz[g * ld + min(oh + tx + 16 * c, on - 1)];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[c][e] = fmaf(d, hv[e], acc[c][e]);
          }
        }
#pragma unroll
        for (int c = 0; c < CW; ++c) {
          const int o = oh + tx + 16 * c;
          if (o >= on) continue;
          float* q = pw + (size_t)o * in + j;
          *reinterpret_cast<float4*>(q) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
          if (wide) *reinterpret_cast<float4*>(q + 4) = make_float4(acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
        }
      }
    }
    for (int o = threadIdx.x; o < on; o += kReadout2Threads) {
      float acc = 0.f;
      for (int g = 0; g < GB; ++g) acc += dz[g * ld + o];
      pw[(size_t)on * in + o] = acc;
    }
    // ---- gradient w.r.t. the layer input: graphs g = ty + 32 a (a < 2), inputs j = jb + 8 tx .. + 7
    for (int jb = 0; jb < in; jb += 128) {
      constexpr int GA = GB / 32;
      const int j = jb + 8 * tx;
      if (j >= in) continue;
      const bool wide = j + 8 <= in;
      float acc[GA][8];
#pragma unroll
      for (int a = 0; a < GA; ++a)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[a][e] = 0.f;
      for (int o = 0; o < on; ++o) {
        const float4 w0 = *reinterpret_cast<const float4*>(Ws + o * ld + j);
        const float4 w1 = wide ? *reinterpret_cast<const float4*>(Ws + o * ld + j + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int a = 0; a < GA; ++a) {
          const float d = dz[(ty + 32 * a) * ld + o];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[a][e] = fmaf(d, wv[e], acc[a][e]);
        }
      }
#pragma unroll
      for (int a = 0; a < GA; ++a) {
        const int g = ty + 32 * a;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (!wide && e >= 4) continue;
          float v = acc[a][e];
          if (k > 0) {
            float y = hin[g * ld + j + e], f = 1.f;
            if (dr_base.p > 0.f) {
              Drop dr = dr_base;
              dr.tag += (uint32_t)k;
              f = dropout_keep(dr.seed, dr.tag, (uint64_t)(g0 + g) * in + j + e, dr.p) ? 1.f / (1.f - dr.p) : 0.f;
              y *= (1.f - dr.p);
            }
            dh[g * ld + j + e] = g < ng ? v * f * act_grad_from_out(y, act) : 0.f;
          } else if (g < ng) {
            dPool[(size_t)(g0 + g) * in + j + e] = v;
          }
        }
      }
    }
    __syncthreads();
    float* t = dz; dz = dh; dh = t;
  }
}

}  // namespace gcb
