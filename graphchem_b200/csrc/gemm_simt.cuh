// Generic CUDA-core GEMMs (fp32 FFMA accumulation) with fused epilogues.
//
// These serve (a) the fp32 parity mode, where 1e-5 agreement with the reference rules out single
// pass TF32/BF16 tensor-core math, (b) the tiny readout MLP (gcn.py:184-199) and
// (c) every shape the tcgen05 kernels in gemm_tc.cuh do not cover.  All reductions have a fixed
// order: results are bit-reproducible run to run.
#pragma once
#include <cstdlib>
#include "common.cuh"

namespace gcb {

// Epilogue of the row-major "NT" GEMM:  C[r,n] = f( acc[r,n] + w(r) * bias[n] + res[r,n] )
//   w(r) = 1                         when rowptr == nullptr
//        = deg(r)                    when rowptr != nullptr, aggr == ADD   (PyG adds the bias per edge)
//        = deg(r) > 0                when rowptr != nullptr, aggr == MEAN
//   f    = act (act >= 0), identity (act < 0); or, when actgrad_y != nullptr,
//          C = acc * act'(y[r,n]) * dropout-factor   (backward through an activation)
struct Epilogue {
  const float* bias = nullptr;
  const int32_t* rowptr = nullptr;
  int aggr = GCB_AGGR_ADD;
  const void* res = nullptr;  // same dtype as C, leading dimension ld_res
  int ld_res = 0;
  const int32_t* res_idx = nullptr;  // when set: row r adds res[res_idx[r]] (a lookup table: the pre-activated embedding rows)
  int act = -1;
  const void* actgrad_y = nullptr;  // same dtype as C, leading dimension ld_y
  int ld_y = 0;
  // dropout applied to the tensor y was produced with (training only, p > 0)
  float p_drop = 0.f;
  uint64_t seed = 0;
  uint32_t drop_tag = 0;
  int rev = 0;  // tcgen05 path only: walk the row tiles last-to-first
};

template <typename T, int CNT>
__device__ __forceinline__ void load_seg(const T* __restrict__ row, int k0, int K, bool vec_ok, float (&out)[CNT]) {
  // CNT contiguous elements row[k0 .. k0+CNT), zero-filled past K.
  if (CNT >= 4 && vec_ok && k0 + CNT <= K) {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
      for (int i = 0; i < CNT; i += 4) {
        float4 t = __ldg(reinterpret_cast<const float4*>(row + k0 + i));
        out[i] = t.x; out[i + 1] = t.y; out[i + 2] = t.z; out[i + 3] = t.w;
      }
    } else {
      if constexpr (CNT == 8) {
        uint4 t = __ldg(reinterpret_cast<const uint4*>(row + k0));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); out[2 * i] = f.x; out[2 * i + 1] = f.y; }
      } else {  // CNT == 4
        uint2 t = __ldg(reinterpret_cast<const uint2*>(row + k0));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 2; ++i) { float2 f = __bfloat1622float2(h[i]); out[2 * i] = f.x; out[2 * i + 1] = f.y; }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < CNT; ++i) out[i] = (k0 + i < K) ? to_float(row[k0 + i]) : 0.f;
  }
}

// C[M,N] = epilogue( A[M,K] * W ), W given as [N,K] row-major (W_KN == false, y = x W^T like
// nn.Linear) or as [K,N] row-major (W_KN == true).
// TM = rows per thread: 8 -> 128-row tiles (large problems), 2 -> 32-row tiles (readout MLP, small batches).
template <typename T, bool W_KN, int TM>
__global__ void __launch_bounds__(256)
gemm_nt_kernel(const T* __restrict__ A, int lda, const T* __restrict__ W, int ldw, T* __restrict__ C, int ldc,
               int M, int N, int K, Epilogue ep) {
  constexpr int BM = 16 * TM, BN = 64, BK = 16;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Ws[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  // A loader: row a_r (0..BM-1), TM contiguous k starting at a_k
  const int a_r = tid % BM, a_k = (tid / BM) * TM;
  const bool a_vec = (lda % 8 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool a_row_ok = (m0 + a_r) < M;
  const T* a_row = A + (size_t)(a_row_ok ? m0 + a_r : 0) * lda;
  // W loader
  int w_n, w_k;
  if (W_KN) { w_k = tid >> 4; w_n = (tid & 15) * 4; } else { w_n = tid & 63; w_k = (tid >> 6) * 4; }
  const bool w_vec = (ldw % 4 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);

  float a_reg[TM], w_reg[4];
  auto fetch = [&](int k0) {
    if (a_row_ok) load_seg<T, TM>(a_row, k0 + a_k, K, a_vec, a_reg);
    else {
#pragma unroll
      for (int i = 0; i < TM; ++i) a_reg[i] = 0.f;
    }
    if (W_KN) {
      const int k = k0 + w_k;
      if (k < K) load_seg<T, 4>(W + (size_t)k * ldw, n0 + w_n, N, w_vec, w_reg);
      else { w_reg[0] = w_reg[1] = w_reg[2] = w_reg[3] = 0.f; }
    } else {
      const int n = n0 + w_n;
      if (n < N) load_seg<T, 4>(W + (size_t)n * ldw, k0 + w_k, K, w_vec, w_reg);
      else { w_reg[0] = w_reg[1] = w_reg[2] = w_reg[3] = 0.f; }
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < TM; ++i) As[buf][a_k + i][a_r] = a_reg[i];
    if (W_KN) {
      *reinterpret_cast<float4*>(&Ws[buf][w_k][w_n]) = make_float4(w_reg[0], w_reg[1], w_reg[2], w_reg[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) Ws[buf][w_k + i][w_n] = w_reg[i];
    }
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) fetch((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[buf][k][ty * TM + i];
      float4 w = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float b[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) stash(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue ----
  const T* res = reinterpret_cast<const T*>(ep.res);
  const T* ygrad = reinterpret_cast<const T*>(ep.actgrad_y);
  const float keep_scale = ep.p_drop > 0.f ? 1.f / (1.f - ep.p_drop) : 1.f;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = m0 + ty * TM + i;
    if (r >= M) continue;
    float wrow = 1.f;
    if (ep.rowptr) {
      const int deg = ep.rowptr[r + 1] - ep.rowptr[r];
      wrow = ep.aggr == GCB_AGGR_MEAN ? (deg > 0 ? 1.f : 0.f) : (float)deg;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (ygrad) {
        float y = to_float(ygrad[(size_t)r * ep.ld_y + n]);
        float f = 1.f;
        if (ep.p_drop > 0.f) {
          const bool keep = dropout_keep(ep.seed, ep.drop_tag, (uint64_t)r * (uint64_t)N + n, ep.p_drop);
          f = keep ? keep_scale : 0.f;
          y = y * (1.f - ep.p_drop);  // undo the 1/(1-p) scaling to recover act(x)
        }
        v = v * f * act_grad_t<T>(y, ep.act);
      } else {
        if (ep.bias) v = fmaf(wrow, ep.bias[n], v);
        if (res) v += to_float(res[(size_t)(ep.res_idx ? ep.res_idx[r] : r) * ep.ld_res + n]);
        if (ep.act >= 0) v = act_fwd_t<T>(v, ep.act);
      }
      C[(size_t)r * ldc + n] = from_float<T>(v);
    }
  }
}

// Small-problem form (the notebooks' 16-molecule batches: a few hundred rows, K <= 256): the step is a chain of launch
// latencies there, and the tiled kernel above spends K/16 global-load latencies + barriers per launch (19 us for
// [170 x 256] x [256 x 128]).  Here a CTA owns a [32 x 32] output tile and fetches the WHOLE K extent of its A rows and W
// rows in one go (all loads in flight at once, one barrier), then every thread accumulates a 2 x 2 block with 16-byte
// shared-memory loads.  Same accumulation order (k ascending, one fmaf per term) as the tiled kernel: bit-identical results.
template <typename T, bool W_KN>
__global__ void __launch_bounds__(256)
gemm_nt_small_kernel(const T* __restrict__ A, int lda, const T* __restrict__ W, int ldw, T* __restrict__ C, int ldc,
                     int M, int N, int K, Epilogue ep) {
  constexpr int BM = 32, BN = 32;
  extern __shared__ __align__(16) float sm_small[];
  const int ld = K + 4;                 // row stride: 16-byte aligned, 4 banks off per row
  float* As = sm_small;                 // [BM][ld]
  float* Ws = sm_small + BM * ld;       // [BN][ld]
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int k4n = K >> 2;
  const bool a_vec = (lda % 8 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool w_vec = (ldw % 8 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  for (int i = tid; i < BM * k4n; i += 256) {
    const int r = i / k4n, k = (i - r * k4n) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (m0 + r < M) load_seg<T, 4>(A + (size_t)(m0 + r) * lda, k, K, a_vec, v);
    *reinterpret_cast<float4*>(As + r * ld + k) = make_float4(v[0], v[1], v[2], v[3]);
  }
  if (W_KN) {  // W[k][n]: a thread reads 4 consecutive n of one k and scatters them over 4 staged rows
    for (int i = tid; i < K * (BN / 4); i += 256) {
      const int k = i / (BN / 4), n = (i - k * (BN / 4)) * 4;
      float v[4];
      load_seg<T, 4>(W + (size_t)k * ldw, n0 + n, N, w_vec, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) Ws[(n + j) * ld + k] = v[j];
    }
  } else {
    for (int i = tid; i < BN * k4n; i += 256) {
      const int n = i / k4n, k = (i - n * k4n) * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (n0 + n < N) load_seg<T, 4>(W + (size_t)(n0 + n) * ldw, k, K, w_vec, v);
      *reinterpret_cast<float4*>(Ws + n * ld + k) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;  // rows ty*2 + i, columns tx + 16 j
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  const float* a0 = As + (ty * 2) * ld;
  const float* a1 = a0 + ld;
  const float* w0 = Ws + tx * ld;
  const float* w1 = Ws + (tx + 16) * ld;
  for (int k = 0; k < K; k += 4) {
    const float4 x0 = *reinterpret_cast<const float4*>(a0 + k), x1 = *reinterpret_cast<const float4*>(a1 + k);
    const float4 y0 = *reinterpret_cast<const float4*>(w0 + k), y1 = *reinterpret_cast<const float4*>(w1 + k);
    const float xa[2][4] = {{x0.x, x0.y, x0.z, x0.w}, {x1.x, x1.y, x1.z, x1.w}};
    const float yb[2][4] = {{y0.x, y0.y, y0.z, y0.w}, {y1.x, y1.y, y1.z, y1.w}};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j] = fmaf(xa[i][q], yb[j][q], acc[i][j]);
  }
  const T* res = reinterpret_cast<const T*>(ep.res);
  const T* ygrad = reinterpret_cast<const T*>(ep.actgrad_y);
  const float keep_scale = ep.p_drop > 0.f ? 1.f / (1.f - ep.p_drop) : 1.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = m0 + ty * 2 + i;
    if (r >= M) continue;
    float wrow = 1.f;
    if (ep.rowptr) {
      const int deg = ep.rowptr[r + 1] - ep.rowptr[r];
      wrow = ep.aggr == GCB_AGGR_MEAN ? (deg > 0 ? 1.f : 0.f) : (float)deg;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (ygrad) {
        float y = to_float(ygrad[(size_t)r * ep.ld_y + n]);
        float f = 1.f;
        if (ep.p_drop > 0.f) {
          const bool keep = dropout_keep(ep.seed, ep.drop_tag, (uint64_t)r * (uint64_t)N + n, ep.p_drop);
          f = keep ? keep_scale : 0.f;
          y = y * (1.f - ep.p_drop);
        }
        v = v * f * act_grad_t<T>(y, ep.act);
      } else {
        if (ep.bias) v = fmaf(wrow, ep.bias[n], v);
        if (res) v += to_float(res[(size_t)(ep.res_idx ? ep.res_idx[r] : r) * ep.ld_res + n]);
        if (ep.act >= 0) v = act_fwd_t<T>(v, ep.act);
      }
      C[(size_t)r * ldc + n] = from_float<T>(v);
    }
  }
}

constexpr int kSmallGemmRows = 2048, kSmallGemmK = 256;

template <typename T>
inline int launch_gemm_nt(const T* A, int lda, const T* W, int ldw, bool w_kn, T* C, int ldc, int M, int N, int K,
                          const Epilogue& ep, cudaStream_t s) {
  if (M <= 0 || N <= 0) return GCB_OK;
  static const bool small_off = std::getenv("GCB_NO_SMALL_GEMM") && std::getenv("GCB_NO_SMALL_GEMM")[0] == '1';
  if (!small_off && M <= kSmallGemmRows && K <= kSmallGemmK && K % 4 == 0 && K > 0) {
    const int smem = 64 * (K + 4) * (int)sizeof(float);  // <= 66.6 KB
    static bool attr = false;
    if (!attr) {
      GCB_CUDA(cudaFuncSetAttribute(gemm_nt_small_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * (kSmallGemmK + 4) * 4));
      GCB_CUDA(cudaFuncSetAttribute(gemm_nt_small_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * (kSmallGemmK + 4) * 4));
      attr = true;
    }
    dim3 grid((unsigned)ceil_div(M, 32), (unsigned)ceil_div(N, 32));
    if (w_kn) {
      GCB_KERNEL("gemm_nt_kernel", s, gemm_nt_small_kernel<T, true><<<grid, 256, smem, s>>>(A, lda, W, ldw, C, ldc, M, N, K, ep));
    } else {
      GCB_KERNEL("gemm_nt_kernel", s, gemm_nt_small_kernel<T, false><<<grid, 256, smem, s>>>(A, lda, W, ldw, C, ldc, M, N, K, ep));
    }
    return GCB_OK;
  }
  if (M > 16384) {
    dim3 grid((unsigned)ceil_div(M, 128), (unsigned)ceil_div(N, 64));
    if (w_kn) {
      GCB_KERNEL("gemm_nt_kernel", s, gemm_nt_kernel<T, true, 8><<<grid, 256, 0, s>>>(A, lda, W, ldw, C, ldc, M, N, K, ep));
    } else {
      GCB_KERNEL("gemm_nt_kernel", s, gemm_nt_kernel<T, false, 8><<<grid, 256, 0, s>>>(A, lda, W, ldw, C, ldc, M, N, K, ep));
    }
  } else {
    dim3 grid((unsigned)ceil_div(M, 32), (unsigned)ceil_div(N, 64));
    if (w_kn) {
      GCB_KERNEL("gemm_nt_kernel", s, gemm_nt_kernel<T, true, 2><<<grid, 256, 0, s>>>(A, lda, W, ldw, C, ldc, M, N, K, ep));
    } else {
      GCB_KERNEL("gemm_nt_kernel", s, gemm_nt_kernel<T, false, 2><<<grid, 256, 0, s>>>(A, lda, W, ldw, C, ldc, M, N, K, ep));
    }
  }
  return GCB_OK;
}

// ---- "TN" GEMM: weight gradients ----------------------------------------------------------------
// partial[z][n][k] = sum_{r in split z} A[r,n] * B[r,k]      (A: [rows,Nn], B: [rows,Kk])
// bias_partial[z][n] = sum_{r in split z} w(r) * A[r,n]       (optional, from the k-tile-0 CTAs)
// Splits are contiguous row ranges; a second kernel adds the partials in split order.
template <typename T>
__global__ void __launch_bounds__(256)
gemm_tn_kernel(const T* __restrict__ A, int lda, const T* __restrict__ B, int ldb, float* __restrict__ partial,
               float* __restrict__ bias_partial, const int32_t* __restrict__ rowptr, int aggr, int rows, int Nn,
               int Kk, int rows_per_split, int direct_accumulate = -1) {
  // direct_accumulate >= 0 (single split): `partial` / `bias_partial` ARE dW / db, written (0) or added to (1) here --
  // the same values the one-split reduction would produce, without its two launches
  constexpr int BT = 64, BR = 16;
  __shared__ __align__(16) float As[2][BR][BT + 4];
  __shared__ __align__(16) float Bs[2][BR][BT + 4];
  __shared__ float wrow_s[2][BR];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * BT, n0 = blockIdx.y * BT, z = blockIdx.z;
  const int r_begin = z * rows_per_split;
  const int r_end = min(rows, r_begin + rows_per_split);
  const int l_r = tid >> 4, l_c = (tid & 15) * 4;
  const bool a_vec = (lda % 8 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool b_vec = (ldb % 8 == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  const bool want_bias = bias_partial != nullptr && blockIdx.x == 0;

  float a_reg[4], b_reg[4], w_reg = 0.f;
  auto fetch = [&](int r0) {
    const int r = r0 + l_r;
    if (r < r_end) {
      load_seg<T, 4>(A + (size_t)r * lda, n0 + l_c, Nn, a_vec, a_reg);
      load_seg<T, 4>(B + (size_t)r * ldb, k0 + l_c, Kk, b_vec, b_reg);
      if (want_bias && l_c == 0) {
        if (rowptr) {
          const int deg = rowptr[r + 1] - rowptr[r];
          w_reg = aggr == GCB_AGGR_MEAN ? (deg > 0 ? 1.f : 0.f) : (float)deg;
        } else w_reg = 1.f;
      }
    } else {
      a_reg[0] = a_reg[1] = a_reg[2] = a_reg[3] = 0.f;
      b_reg[0] = b_reg[1] = b_reg[2] = b_reg[3] = 0.f;
      w_reg = 0.f;
    }
  };
  auto stash = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][l_r][l_c]) = make_float4(a_reg[0], a_reg[1], a_reg[2], a_reg[3]);
    *reinterpret_cast<float4*>(&Bs[buf][l_r][l_c]) = make_float4(b_reg[0], b_reg[1], b_reg[2], b_reg[3]);
    if (want_bias && l_c == 0) wrow_s[buf][l_r] = w_reg;
  };

  const int ty = tid >> 4, tx = tid & 15;  // n = n0 + ty*4 + i, k = k0 + tx*4 + j
  float acc[4][4];
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nt = (r_end - r_begin + BR - 1) / BR;
  if (nt > 0) {
    fetch(r_begin);
    stash(0);
  }
  __syncthreads();
  for (int t = 0; t < nt; ++t) {
    const int buf = t & 1;
    if (t + 1 < nt) fetch(r_begin + (t + 1) * BR);
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[buf][r][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][r][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (want_bias && tx == 0) {
        const float w = wrow_s[buf][r];
#pragma unroll
        for (int i = 0; i < 4; ++i) bacc[i] = fmaf(w, a[i], bacc[i]);
      }
    }
    if (t + 1 < nt) stash(buf ^ 1);
    __syncthreads();
  }
  float* out = partial + (size_t)z * Nn * Kk;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= Nn) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < Kk) out[(size_t)n * Kk + k] = direct_accumulate == 1 ? out[(size_t)n * Kk + k] + acc[i][j] : acc[i][j];
    }
    if (want_bias && tx == 0) bias_partial[(size_t)z * Nn + n] = direct_accumulate == 1 ? bias_partial[(size_t)z * Nn + n] + bacc[i] : bacc[i];
  }
}

// out[i] (+)= sum_{z < splits} partial[z*count + i]   (fixed order)
// 256 threads = 64 elements x 4 z-groups; each group adds its contiguous quarter of the splits
// in order, the four group sums are combined as (g0 + g1) + (g2 + g3): a fixed summation tree.
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partial, int splits, int64_t stride, int64_t count,
                       float* __restrict__ out, int accumulate) {
  __shared__ float red[4][64];
  const int e = threadIdx.x & 63, zg = threadIdx.x >> 6;
  const int64_t i = (int64_t)blockIdx.x * 64 + e;
  const int per = (splits + 3) / 4;
  const int z0 = zg * per, z1 = min(splits, z0 + per);
  float s = 0.f;
  if (i < count) {
#pragma unroll 4
    for (int z = z0; z < z1; ++z) s += __ldg(partial + (size_t)z * stride + i);
  }
  red[zg][e] = s;
  __syncthreads();
  if (zg == 0 && i < count) {
    const float t = (red[0][e] + red[1][e]) + (red[2][e] + red[3][e]);
    out[i] = accumulate ? out[i] + t : t;
  }
}

inline int launch_reduce_partials(const float* partial, int splits, int64_t count, float* out, int accumulate,
                                  cudaStream_t s, int64_t stride = -1) {
  if (count <= 0) return GCB_OK;
  if (stride < 0) stride = count;
  GCB_KERNEL("reduce_partials_kernel", s, reduce_partials_kernel<<<(unsigned)ceil_div(count, 64), 256, 0, s>>>(partial, splits, stride, count, out, accumulate));
  return GCB_OK;
}

// Small-row form of the weight-gradient GEMM (rows <= kTnDirectRows: the notebooks' 16-molecule batches): a CTA owns a
// [32 x 32] block of dW and stages 128 rows of its A / B column windows per barrier (the tiled kernel above: 16 rows per
// barrier, one global-load latency each -- 11 round trips for 170 rows, ~13 us per launch of a step that is a chain of
// launch latencies).  Writes dW / db directly.  Same accumulation order (rows ascending, one fmaf per term): bit-identical.
template <typename T>
__global__ void __launch_bounds__(256)
gemm_tn_small_kernel(const T* __restrict__ A, int lda, const T* __restrict__ B, int ldb, float* __restrict__ dW,
                     float* __restrict__ db, const int32_t* __restrict__ rowptr, int aggr, int rows, int Nn, int Kk,
                     int accumulate) {
  constexpr int BT = 32, RC = 128;
  __shared__ __align__(16) float As[RC][BT + 4];
  __shared__ __align__(16) float Bs[RC][BT + 4];
  __shared__ float ws[RC];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * BT, n0 = blockIdx.y * BT;
  const bool a_vec = (lda % 8 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool b_vec = (ldb % 8 == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  const bool want_bias = db != nullptr && blockIdx.x == 0;
  const int ty = tid >> 4, tx = tid & 15;  // n = n0 + ty*2 + i, k = k0 + tx + 16 j
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  float bacc[2] = {0.f, 0.f};
  for (int r0 = 0; r0 < rows; r0 += RC) {
    for (int i = tid; i < RC * (BT / 4); i += 256) {
      const int r = i >> 3, c4 = (i & 7) * 4, row = r0 + r;
      float va[4] = {0.f, 0.f, 0.f, 0.f}, vb[4] = {0.f, 0.f, 0.f, 0.f};
      if (row < rows) {
        load_seg<T, 4>(A + (size_t)row * lda, n0 + c4, Nn, a_vec, va);
        load_seg<T, 4>(B + (size_t)row * ldb, k0 + c4, Kk, b_vec, vb);
      }
      *reinterpret_cast<float4*>(&As[r][c4]) = make_float4(va[0], va[1], va[2], va[3]);
      *reinterpret_cast<float4*>(&Bs[r][c4]) = make_float4(vb[0], vb[1], vb[2], vb[3]);
    }
    if (want_bias && tid < RC) {
      const int row = r0 + tid;
      float w = 0.f;
      if (row < rows) {
        if (rowptr) {
          const int deg = rowptr[row + 1] - rowptr[row];
          w = aggr == GCB_AGGR_MEAN ? (deg > 0 ? 1.f : 0.f) : (float)deg;
        } else w = 1.f;
      }
      ws[tid] = w;
    }
    __syncthreads();
    const int nr = min(RC, rows - r0);
    for (int r = 0; r < nr; ++r) {
      const float2 a2 = *reinterpret_cast<const float2*>(&As[r][ty * 2]);
      const float b0 = Bs[r][tx], b1 = Bs[r][tx + 16];
      acc[0][0] = fmaf(a2.x, b0, acc[0][0]); acc[0][1] = fmaf(a2.x, b1, acc[0][1]);
      acc[1][0] = fmaf(a2.y, b0, acc[1][0]); acc[1][1] = fmaf(a2.y, b1, acc[1][1]);
      if (want_bias && tx == 0) {
        const float w = ws[r];
        bacc[0] = fmaf(w, a2.x, bacc[0]);
        bacc[1] = fmaf(w, a2.y, bacc[1]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int n = n0 + ty * 2 + i;
    if (n >= Nn) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = k0 + tx + 16 * j;
      if (k < Kk) {
        float* o = dW + (size_t)n * Kk + k;
        *o = accumulate ? *o + acc[i][j] : acc[i][j];
      }
    }
    if (want_bias && tx == 0) db[n] = accumulate ? db[n] + bacc[i] : bacc[i];
  }
}

// One split up to kTnDirectRows rows: in the small-batch regime (the notebooks' 16-molecule batches, ~200 rows) a step is a
// chain of ~60 launch latencies, and the two ordered reductions behind every weight-gradient GEMM are a fifth of them.
constexpr int kTnDirectRows = 512;
inline int tn_splits(int rows) {
  if (rows <= 0) return 0;
  if (rows <= kTnDirectRows) return 1;
  int s = (int)ceil_div(rows, 128);
  if (s > 64) s = 64;
  return s;
}

// dW[Nn,Kk] (+)= A^T B ; db[Nn] (+)= sum_r w(r) A[r,:]  (db may be null).  partial needs
// tn_splits(rows) * Nn * (Kk + 1) floats.
template <typename T>
inline int launch_gemm_tn(const T* A, int lda, const T* B, int ldb, int rows, int Nn, int Kk, float* partial,
                          float* dW, float* db, const int32_t* rowptr, int aggr, int accumulate, cudaStream_t s) {
  if (Nn <= 0 || Kk <= 0) return GCB_OK;
  const int splits = tn_splits(rows);
  if (splits == 1) {  // direct: no partials, no reductions
    static const bool small_off = std::getenv("GCB_NO_SMALL_GEMM") && std::getenv("GCB_NO_SMALL_GEMM")[0] == '1';
    if (!small_off) {
      dim3 grid((unsigned)ceil_div(Kk, 32), (unsigned)ceil_div(Nn, 32), 1);
      GCB_KERNEL("gemm_tn_kernel", s, gemm_tn_small_kernel<T><<<grid, 256, 0, s>>>(A, lda, B, ldb, dW, db, rowptr, aggr, rows, Nn, Kk, accumulate ? 1 : 0));
      return GCB_OK;
    }
    dim3 grid((unsigned)ceil_div(Kk, 64), (unsigned)ceil_div(Nn, 64), 1);
    GCB_KERNEL("gemm_tn_kernel", s, gemm_tn_kernel<T><<<grid, 256, 0, s>>>(A, lda, B, ldb, dW, db, rowptr, aggr, rows, Nn, Kk,
                                                                            (int)align_up(rows, 16), accumulate ? 1 : 0));
    return GCB_OK;
  }
  float* bias_partial = db ? partial + (size_t)splits * Nn * Kk : nullptr;
  if (splits > 0) {
    const int rps = (int)align_up(ceil_div(rows, splits), 16);
    dim3 grid((unsigned)ceil_div(Kk, 64), (unsigned)ceil_div(Nn, 64), (unsigned)splits);
    GCB_KERNEL("gemm_tn_kernel", s, gemm_tn_kernel<T><<<grid, 256, 0, s>>>(A, lda, B, ldb, partial, bias_partial, rowptr, aggr, rows, Nn, Kk, rps));
  }
  GCB_TRY(launch_reduce_partials(partial, splits, (int64_t)Nn * Kk, dW, accumulate, s));
  if (db) GCB_TRY(launch_reduce_partials(bias_partial, splits, Nn, db, accumulate, s));
  return GCB_OK;
}

}  // namespace gcb
