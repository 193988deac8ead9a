// Micro-benchmark: what HBM bandwidth do 16-byte-per-thread access patterns reach on B200 as a
// function of loads in flight per thread and access pattern (stream / neighbour gather)?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

template <int U>
__global__ void __launch_bounds__(256) copy_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
  size_t i = ((size_t)blockIdx.x * 256 + threadIdx.x);
  const size_t stride = (size_t)gridDim.x * 256;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(in + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) out[i + u * stride] = v[u];
  }
}

// rows of 256 B (16 lanes x 16 B); each row sums K neighbour rows given by idx (row + small offset)
template <int K>
__global__ void __launch_bounds__(256) gather_kernel(const uint4* __restrict__ in, const int* __restrict__ idx,
                                                    uint4* __restrict__ out, int rows) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  const int r = t >> 4, c = t & 15;
  if (r >= rows) return;
  int s[K];
#pragma unroll
  for (int k = 0; k < K; ++k) s[k] = __ldg(idx + (size_t)r * K + k);
  uint4 v[K];
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = __ldg(in + (size_t)s[k] * 16 + c);
  uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int k = 0; k < K; ++k) { acc.x ^= v[k].x; acc.y ^= v[k].y; acc.z ^= v[k].z; acc.w ^= v[k].w; }
  out[(size_t)r * 16 + c] = acc;
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

int main() {
  const size_t bytes = 768ull << 20;  // 768 MB in, 768 MB out: far beyond L2
  const size_t n = bytes / 16;
  uint4 *in, *out;
  cudaMalloc(&in, bytes); cudaMalloc(&out, bytes);
  cudaMemset(in, 1, bytes); cudaMemset(out, 0, bytes);
  printf("copy (read+write bytes), grid = all threads or persistent\n");
  for (int grid : {148 * 8, 148 * 16, 148 * 32, 0}) {
    const int g = grid ? grid : (int)(n / 256);
    float t1 = time_ms([&] { copy_kernel<1><<<g, 256>>>(in, out, n); });
    float t2 = time_ms([&] { copy_kernel<2><<<grid ? g : g / 2, 256>>>(in, out, n); });
    float t4 = time_ms([&] { copy_kernel<4><<<grid ? g : g / 4, 256>>>(in, out, n); });
    float t8 = time_ms([&] { copy_kernel<8><<<grid ? g : g / 8, 256>>>(in, out, n); });
    printf("  grid %8d: U=1 %6.0f  U=2 %6.0f  U=4 %6.0f  U=8 %6.0f GB/s\n", g, 2 * bytes / t1 / 1e6, 2 * bytes / t2 / 1e6,
           2 * bytes / t4 / 1e6, 2 * bytes / t8 / 1e6);
  }
  // neighbour gather: rows = 3M (768 MB), neighbours within +-23 rows
  const int rows = (int)(n / 16);
  for (int K : {1, 2, 4}) {
    std::vector<int> h((size_t)rows * K);
    srand(1);
    for (int r = 0; r < rows; ++r)
      for (int k = 0; k < K; ++k) { int s = r + (rand() % 47) - 23; h[(size_t)r * K + k] = s < 0 ? 0 : (s >= rows ? rows - 1 : s); }
    int* idx; cudaMalloc(&idx, h.size() * 4);
    cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const int g = (rows * 16 + 255) / 256;
    float t = 0;
    if (K == 1) t = time_ms([&] { gather_kernel<1><<<g, 256>>>(in, idx, out, rows); });
    if (K == 2) t = time_ms([&] { gather_kernel<2><<<g, 256>>>(in, idx, out, rows); });
    if (K == 4) t = time_ms([&] { gather_kernel<4><<<g, 256>>>(in, idx, out, rows); });
    // DRAM-level bytes: each input row once (768 MB) + output (768 MB) + idx
    printf("  gather K=%d: %.3f ms  -> %.0f GB/s (unique in + out + idx)\n", K, t, (2.0 * bytes + (double)rows * K * 4) / t / 1e6);
    cudaFree(idx);
  }
  return 0;
}
