// Shared device helpers for the graphchem_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/graphchem_b200.h"

#ifndef __CUDA_ARCH_FEAT_SM100_ALL
#if defined(__CUDA_ARCH__)
#error "graphchem_b200 kernels are written for sm_100a only (compile with -gencode arch=compute_100a,code=sm_100a)"
#endif
#endif

namespace gcb {

typedef __nv_bfloat16 bf16;

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- error plumbing -------------------------------------------------------------------------
void set_cuda_error(cudaError_t e, const char* file, int line);
#define GCB_CUDA(expr)                                  \
  do {                                                  \
    cudaError_t _e = (expr);                            \
    if (_e != cudaSuccess) {                            \
      gcb::set_cuda_error(_e, __FILE__, __LINE__);      \
      return GCB_ERR_CUDA;                              \
    }                                                   \
  } while (0)
#define GCB_LAUNCH_CHECK() GCB_CUDA(cudaGetLastError())
#define GCB_TRY(expr)          \
  do {                         \
    int _rc = (expr);          \
    if (_rc != GCB_OK) return _rc; \
  } while (0)

// ---- optional per-kernel timing (debug/bench facility; thread-local, off by default) ------------
// When enabled with gcb_timing_enable(1) every kernel launch is bracketed by two CUDA events on its
// stream (so it must not be used under stream capture).  Launches are always counted.
struct KernelTimer {
  KernelTimer(const char* name, cudaStream_t s);
  void stop();
  const char* name_;
  cudaStream_t s_;
  int slot_;
};
#define GCB_KERNEL(NAME, STREAM, ...)          \
  do {                                         \
    gcb::KernelTimer _kt(NAME, STREAM);        \
    __VA_ARGS__;                               \
    _kt.stop();                                \
    GCB_LAUNCH_CHECK();                        \
  } while (0)

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------
// Kernels launched through launch_pdl() may begin while the previous kernel of the stream is still draining:
// their set-up (barrier / TMEM initialisation, staging of the immutable batch indices) overlaps its tail, and
// pdl_wait() blocks until that kernel has completed and its writes are visible.  Everything that reads or
// writes activations must come after pdl_wait(); both calls are no-ops in a normally launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();  // GCB_NO_PDL=1 turns the launch attribute off
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// ---- 16-byte vector access: VEC elements of T ---------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<float> { static constexpr int N = 4; };
template <> struct Vec<bf16> { static constexpr int N = 8; };

__device__ __forceinline__ void load_vec(const float* __restrict__ p, float (&v)[4]) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load_vec(const bf16* __restrict__ p, float (&v)[8]) {
  uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store_vec(float* __restrict__ p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store_vec(bf16* __restrict__ p, const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}
// Round-trip through the storage type (what a later reader of the stored value will see).
__device__ __forceinline__ float round_to(float v, const float*) { return v; }
__device__ __forceinline__ float round_to(float v, const bf16*) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_float<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---- activations (gcn.py:42; tests/test_gcn.py:107-113) ----------------------------------------
// Forward follows torch's CPU kernels: softplus(beta=1, threshold=20), sigmoid, relu, tanh,
// leaky_relu(0.01).  The derivative is evaluated from the OUTPUT y = act(x) so that only the
// activations themselves need to be kept for the backward pass.
__device__ __forceinline__ float act_fwd(float x, int act) {
  switch (act) {
    case GCB_ACT_SOFTPLUS: return x > 20.f ? x : log1pf(expf(x));
    case GCB_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case GCB_ACT_RELU: return fmaxf(x, 0.f);
    case GCB_ACT_TANH: return tanhf(x);
    default: return x > 0.f ? x : 0.01f * x;
  }
}
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
  switch (act) {
    case GCB_ACT_SOFTPLUS: return -expm1f(-y);  // sigmoid(x) = 1 - exp(-softplus(x))
    case GCB_ACT_SIGMOID: return y * (1.f - y);
    case GCB_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case GCB_ACT_TANH: return 1.f - y * y;
    default: return y > 0.f ? 1.f : 0.01f;
  }
}

// Fast variants for the bf16 mode (results are rounded to 8 mantissa bits anyway): one MUFU per
// transcendental, flush-to-zero forms issued directly -- `__expf` / `__logf` without -ftz wrap every
// MUFU in denormal range-scaling code (~3x the instructions), and these kernels are issue-bound.
// ~1e-6 relative error.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float act_fwd_fast(float x, int act) {
  switch (act) {
    case GCB_ACT_SOFTPLUS: return x > 20.f ? x : kLn2 * lg2_ftz(1.f + ex2_ftz(kLog2e * x));
    case GCB_ACT_SIGMOID: return rcp_ftz(1.f + ex2_ftz(-kLog2e * x));
    case GCB_ACT_RELU: return fmaxf(x, 0.f);
    case GCB_ACT_TANH: { const float e = ex2_ftz(2.f * kLog2e * fminf(x, 15.f)); return (e - 1.f) * rcp_ftz(e + 1.f); }
    default: return x > 0.f ? x : 0.01f * x;
  }
}
__device__ __forceinline__ float act_grad_from_out_fast(float y, int act) {
  switch (act) {
    case GCB_ACT_SOFTPLUS: return 1.f - ex2_ftz(-kLog2e * y);
    case GCB_ACT_SIGMOID: return y * (1.f - y);
    case GCB_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case GCB_ACT_TANH: return 1.f - y * y;
    default: return y > 0.f ? 1.f : 0.01f;
  }
}
// Storage-type dispatch: fp32 = parity mode (torch-like precise math), bf16 = fast math.
template <typename T> __device__ __forceinline__ float act_fwd_t(float x, int act) {
  if constexpr (sizeof(T) == 2) return act_fwd_fast(x, act); else return act_fwd(x, act);
}
template <typename T> __device__ __forceinline__ float act_grad_t(float y, int act) {
  if constexpr (sizeof(T) == 2) return act_grad_from_out_fast(y, act); else return act_grad_from_out(y, act);
}

// Vector forms.  A switch on `act` per ELEMENT compiles to one indirect branch per element and
// serialises the MUFU latencies; one uniform switch per VECTOR leaves N independent evaluations of
// straight-line code inside each case.  Same arithmetic per element as the scalar forms.
template <bool FAST, int N>
__device__ __forceinline__ void act_fwd_vec_impl(float (&v)[N], int act) {
#define GCB_ACT_CASE(A)                                                                   \
  case A: {                                                                               \
    _Pragma("unroll") for (int i = 0; i < N; ++i) v[i] = FAST ? act_fwd_fast(v[i], A) : act_fwd(v[i], A); \
    break;                                                                                \
  }
  switch (act) {
    GCB_ACT_CASE(GCB_ACT_SOFTPLUS)
    GCB_ACT_CASE(GCB_ACT_SIGMOID)
    GCB_ACT_CASE(GCB_ACT_RELU)
    GCB_ACT_CASE(GCB_ACT_TANH)
    default: {
      _Pragma("unroll") for (int i = 0; i < N; ++i) v[i] = FAST ? act_fwd_fast(v[i], GCB_ACT_LEAKY_RELU) : act_fwd(v[i], GCB_ACT_LEAKY_RELU);
    }
  }
#undef GCB_ACT_CASE
}
// d[i] *= act'(y[i]) with y the stored activation output
template <bool FAST, int N>
__device__ __forceinline__ void act_grad_mul_vec_impl(float (&d)[N], const float (&y)[N], int act) {
#define GCB_ACT_CASE(A)                                                                   \
  case A: {                                                                               \
    _Pragma("unroll") for (int i = 0; i < N; ++i) d[i] *= FAST ? act_grad_from_out_fast(y[i], A) : act_grad_from_out(y[i], A); \
    break;                                                                                \
  }
  switch (act) {
    GCB_ACT_CASE(GCB_ACT_SOFTPLUS)
    GCB_ACT_CASE(GCB_ACT_SIGMOID)
    GCB_ACT_CASE(GCB_ACT_RELU)
    GCB_ACT_CASE(GCB_ACT_TANH)
    default: {
      _Pragma("unroll") for (int i = 0; i < N; ++i) d[i] *= FAST ? act_grad_from_out_fast(y[i], GCB_ACT_LEAKY_RELU) : act_grad_from_out(y[i], GCB_ACT_LEAKY_RELU);
    }
  }
#undef GCB_ACT_CASE
}
template <typename T, int N> __device__ __forceinline__ void act_fwd_vec(float (&v)[N], int act) {
  act_fwd_vec_impl<sizeof(T) == 2, N>(v, act);
}
template <typename T, int N> __device__ __forceinline__ void act_grad_mul_vec(float (&d)[N], const float (&y)[N], int act) {
  act_grad_mul_vec_impl<sizeof(T) == 2, N>(d, y, act);
}

// ---- counter-based RNG for dropout (Philox-4x32-10) --------------------------------------------
// Keyed by (seed, tensor tag) and indexed by the element's linear index, so that the backward pass
// regenerates the same keep-mask without storing it.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
// keep-mask for element `idx` of tensor `tag`: true with probability 1-p.
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint32_t tag, uint64_t idx, float p) {
  uint4 r = philox4x32(make_uint4((uint32_t)(idx >> 2), (uint32_t)(idx >> 34), tag, 0u),
                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t w = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  return (w >> 8) * (1.0f / 16777216.0f) >= p;
}

}  // namespace gcb
