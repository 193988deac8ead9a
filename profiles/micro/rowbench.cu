// Micro-benchmark: bond_update-shaped kernels with features added one at a time, to find what
// separates a 5 TB/s neighbour gather from the production row kernels.  N rows of PQ [N, 256] bf16
// (P | Q halves of 128), ELL-4 neighbour indices (deg ~ 2.17 +- , -1 padded), output B [N,128] bf16.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
  __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a), y = *reinterpret_cast<__nv_bfloat162*>(&b);
  __nv_bfloat162 m = __hmax2(x, y);
  return *reinterpret_cast<uint32_t*>(&m);
}
__device__ __forceinline__ float softplus_fast(float x) { return x > 20.f ? x : __logf(1.f + __expf(x)); }

// MODE 0: gather+max only (xor-free), write max.  MODE 1: + P chunk, add in fp32, no act.  MODE 2: + softplus.
template <int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB) k_flat(const uint4* __restrict__ PQ, const int4* __restrict__ ell,
                                                    uint4* __restrict__ out, int rows) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int r = t >> 4, c = t & 15;
  if (r >= rows) return;
  const int4 s4 = __ldg(ell + r);
  const int s[4] = {s4.x, s4.y, s4.z, s4.w};
  uint4 q[4];
  uint4 p = make_uint4(0, 0, 0, 0);
  if (MODE >= 1) p = __ldg(PQ + (size_t)r * 32 + c);
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (s[u] >= 0) q[u] = __ldg(PQ + (size_t)s[u] * 32 + 16 + c);
  uint4 m = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (s[u] >= 0) { m.x = max2(m.x, q[u].x); m.y = max2(m.y, q[u].y); m.z = max2(m.z, q[u].z); m.w = max2(m.w, q[u].w); }
  if (MODE >= 1) {
    const uint32_t pw[4] = {p.x, p.y, p.z, p.w}, mw[4] = {m.x, m.y, m.z, m.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = __uint_as_float(pw[i] << 16) + __uint_as_float(mw[i] << 16);
      float b = __uint_as_float(pw[i] & 0xFFFF0000u) + __uint_as_float(mw[i] & 0xFFFF0000u);
      if (MODE >= 2) { a = softplus_fast(a); b = softplus_fast(b); }
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      o[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    m = make_uint4(o[0], o[1], o[2], o[3]);
  }
  out[(size_t)r * 16 + c] = m;
}


__device__ __forceinline__ float act_rt(float x, int act) {
  switch (act) {
    case 0: return x > 20.f ? x : __logf(1.f + __expf(x));
    case 1: return __fdividef(1.f, 1.f + __expf(-x));
    case 2: return fmaxf(x, 0.f);
    case 3: { const float e = __expf(2.f * fminf(x, 15.f)); return __fdividef(e - 1.f, e + 1.f); }
    default: return x > 0.f ? x : 0.01f * x;
  }
}
// FEAT bit0: load in_ptr (e0,e1) and run the overflow loop; bit1: runtime activation switch
template <int FEAT>
__global__ void __launch_bounds__(256) k_feat(const uint4* __restrict__ PQ, const int4* __restrict__ ell,
                                              const int* __restrict__ ptr, const int* __restrict__ csr,
                                              uint4* __restrict__ out, int rows, int act) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int r = t >> 4, c = t & 15;
  if (r >= rows) return;
  const int4 s4 = __ldg(ell + r);
  int e0 = 0, e1 = 0;
  if (FEAT & 1) { e0 = __ldg(ptr + r); e1 = __ldg(ptr + r + 1); }
  const int s[4] = {s4.x, s4.y, s4.z, s4.w};
  uint4 q[4];
  const uint4 p = __ldg(PQ + (size_t)r * 32 + c);
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (s[u] >= 0) q[u] = __ldg(PQ + (size_t)s[u] * 32 + 16 + c);
  uint4 m = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (s[u] >= 0) { m.x = max2(m.x, q[u].x); m.y = max2(m.y, q[u].y); m.z = max2(m.z, q[u].z); m.w = max2(m.w, q[u].w); }
  if (FEAT & 1) {
    for (int e = e0 + 4; e < e1; ++e) {
      const uint4 qq = __ldg(PQ + (size_t)__ldg(csr + e) * 32 + 16 + c);
      m.x = max2(m.x, qq.x); m.y = max2(m.y, qq.y); m.z = max2(m.z, qq.z); m.w = max2(m.w, qq.w);
    }
  }
  const uint32_t pw[4] = {p.x, p.y, p.z, p.w}, mw[4] = {m.x, m.y, m.z, m.w};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float a = __uint_as_float(pw[i] << 16) + __uint_as_float(mw[i] << 16);
    float b = __uint_as_float(pw[i] & 0xFFFF0000u) + __uint_as_float(mw[i] & 0xFFFF0000u);
    if (FEAT & 2) { a = act_rt(a, act); b = act_rt(b, act); } else { a = softplus_fast(a); b = softplus_fast(b); }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    o[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  out[(size_t)r * 16 + c] = make_uint4(o[0], o[1], o[2], o[3]);
}

template <typename F>
float time_ms(F f, int iters = 20) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / iters;
}

int main(int argc, char** argv) {
  const int rows = argc > 1 ? atoi(argv[1]) : 188416;
  const int reps = 8;  // distinct buffers cycled so that inputs exceed L2 for small `rows`
  std::vector<uint4*> PQ(reps), out(reps);
  for (int i = 0; i < reps; ++i) {
    cudaMalloc(&PQ[i], (size_t)rows * 512); cudaMalloc(&out[i], (size_t)rows * 256);
    cudaMemset(PQ[i], 0x3c, (size_t)rows * 512);
  }
  std::vector<int> h((size_t)rows * 4);
  srand(1);
  for (int r = 0; r < rows; ++r) {
    const int deg = 1 + rand() % 3 + (rand() % 6 == 0);  // mean ~2.17
    for (int u = 0; u < 4; ++u) {
      int s = r - 11 + rand() % 23;
      s = s < 0 ? 0 : (s >= rows ? rows - 1 : s);
      h[(size_t)r * 4 + u] = u < deg ? s : -1;
    }
  }
  int4* ell; cudaMalloc(&ell, h.size() * 4);
  cudaMemcpy(ell, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int g = (rows * 16 + 255) / 256;
  const double bytes = (double)rows * (512 + 256 + 16);  // PQ once + out + ell
  int it = 0;
#define RUN(NAME, ...) { float t = time_ms([&] { __VA_ARGS__; ++it; }); printf("%-34s %7.1f us  %6.0f GB/s\n", NAME, t * 1e3, bytes / t / 1e6); }
  RUN("max only            (occ 8)", (k_flat<0, 8><<<g, 256>>>(PQ[it % reps], ell, out[it % reps], rows)));
  RUN("max + P add         (occ 8)", (k_flat<1, 8><<<g, 256>>>(PQ[it % reps], ell, out[it % reps], rows)));
  RUN("max + P + softplus  (occ 8)", (k_flat<2, 8><<<g, 256>>>(PQ[it % reps], ell, out[it % reps], rows)));
  RUN("max + P + softplus  (occ 4)", (k_flat<2, 4><<<g, 256>>>(PQ[it % reps], ell, out[it % reps], rows)));
  RUN("max + P + softplus  (occ 2)", (k_flat<2, 2><<<g, 256>>>(PQ[it % reps], ell, out[it % reps], rows)));
  std::vector<int> hp(rows + 1, 0);
  for (int r = 0; r < rows; ++r) { int d = 0; for (int u = 0; u < 4; ++u) d += h[(size_t)r * 4 + u] >= 0; hp[r + 1] = hp[r] + d; }
  int *ptr, *csr; cudaMalloc(&ptr, hp.size() * 4); cudaMalloc(&csr, (size_t)hp[rows] * 4 + 16);
  cudaMemcpy(ptr, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice); cudaMemset(csr, 0, (size_t)hp[rows] * 4 + 16);
  RUN("+ in_ptr loads/overflow check", (k_feat<1><<<g, 256>>>(PQ[it % reps], ell, ptr, csr, out[it % reps], rows, 0)));
  RUN("+ runtime act switch", (k_feat<2><<<g, 256>>>(PQ[it % reps], ell, ptr, csr, out[it % reps], rows, 0)));
  RUN("+ both", (k_feat<3><<<g, 256>>>(PQ[it % reps], ell, ptr, csr, out[it % reps], rows, 0)));
  return 0;
}
