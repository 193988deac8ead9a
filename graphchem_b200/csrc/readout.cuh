// Fused readout MLP (reference gcn.py:184-199) in fp32: one kernel for the forward chain
// Linear -> act -> dropout ... -> Linear over the pooled graph vectors, one kernel for its backward
// (all weight/bias gradient partials, dPool) plus one ordered reduction.  Replaces three GEMM launches
// forward and ten launches backward; the arithmetic is tiny (<= 0.1 GF at 8192 graphs), so the point is
// launch count and latency, not FLOP/s.  Each CTA owns GB consecutive graphs.
#pragma once
#include "kernels.cuh"

namespace gcb {

constexpr int kReadoutGB = 32;       // graphs per CTA
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
__global__ void __launch_bounds__(256)
readout_fwd_kernel(ReadoutDesc rd, const float* __restrict__ params, ReadoutPtrs hp, float* __restrict__ out,
                   int G, int act, Drop dr_base) {
  extern __shared__ float rs[];
  const int ld = rd.max_dim + 1;
  float* cur = rs;
  float* nxt = rs + kReadoutGB * ld;
  float* Wt = rs + 2 * kReadoutGB * ld;  // current layer's weights, transposed: Wt[j * (on + 1) + o]
  const int g0 = blockIdx.x * kReadoutGB;
  const int ng = min(kReadoutGB, G - g0);
  for (int i = threadIdx.x; i < ng * rd.in_dim[0]; i += 256) {
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
    for (int i = threadIdx.x; i < on * in; i += 256) Wt[(i % in) * onp + (i / in)] = __ldg(W + i);
    __syncthreads();
    for (int i = threadIdx.x; i < ng * on; i += 256) {
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
__global__ void __launch_bounds__(256)
readout_bwd_kernel(ReadoutDesc rd, const float* __restrict__ params, ReadoutPtrs hp,
                   const float* __restrict__ d_out, float* __restrict__ partial, float* __restrict__ dPool, int G,
                   int act, Drop dr_base) {
  extern __shared__ float rs[];
  const int ld = rd.max_dim + 1;
  float* dz = rs;                           // [GB][ld] gradient w.r.t. the current layer's output
  float* hin = rs + kReadoutGB * ld;        // [GB][ld] the current layer's input H[k]
  float* dh = rs + 2 * kReadoutGB * ld;     // [GB][ld] gradient w.r.t. the current layer's input
  float* Ws = rs + 3 * kReadoutGB * ld;     // current layer's weights: Ws[o * (in + 1) + j]
  const int g0 = blockIdx.x * kReadoutGB;
  const int ng = min(kReadoutGB, G - g0);
  float* my_partial = partial + (size_t)blockIdx.x * rd.partial_stride;
  const int o_last = rd.out_dim[rd.n_lin - 1];
  for (int i = threadIdx.x; i < ng * o_last; i += 256) {
    const int g = i / o_last, o = i % o_last;
    dz[g * ld + o] = d_out[(size_t)(g0 + g) * o_last + o];
  }
  for (int k = rd.n_lin - 1; k >= 0; --k) {
    const int in = rd.in_dim[k], on = rd.out_dim[k];
    const float* W = params + rd.w_off[k];
    for (int i = threadIdx.x; i < ng * in; i += 256) {
      const int g = i / in, j = i % in;
      hin[g * ld + j] = hp.h[k][(size_t)(g0 + g) * in + j];
    }
    const int inp = in + 1;
    for (int i = threadIdx.x; i < on * in; i += 256) Ws[(i / in) * inp + (i % in)] = __ldg(W + i);
    __syncthreads();
    // weight / bias gradient partials of this CTA's graphs (graph order: deterministic)
    float* pw = my_partial + rd.p_off[k];
    for (int i = threadIdx.x; i < on * in; i += 256) {
      const int o = i / in, j = i % in;
      float acc = 0.f;
      for (int g = 0; g < ng; ++g) acc = fmaf(dz[g * ld + o], hin[g * ld + j], acc);
      pw[i] = acc;
    }
    for (int o = threadIdx.x; o < on; o += 256) {
      float acc = 0.f;
      for (int g = 0; g < ng; ++g) acc += dz[g * ld + o];
      pw[(size_t)on * in + o] = acc;
    }
    // gradient w.r.t. the layer input; through activation + dropout of H[k] when k > 0
    for (int i = threadIdx.x; i < ng * in; i += 256) {
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

}  // namespace gcb
