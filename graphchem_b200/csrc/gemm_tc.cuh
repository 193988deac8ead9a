// tcgen05 / TMEM / TMA GEMM for the per-layer linear transforms (bf16 in, fp32 accumulate in TMEM).
//
//   C[rows, N] = epilogue( A[rows, K] * W[N, K]^T )          rows ~ 1e5..1e6, K, N in {64,128,256}
//
// The problem is "skinny": the weight matrix (<= 64 KB) stays resident in shared memory for the
// whole kernel, A streams through a TMA-fed ring of 128-row tiles, one elected thread issues
// tcgen05.mma (M=128, N<=256, K=16 per instruction) into a double-buffered TMEM accumulator, and
// eight epilogue warps drain TMEM (tcgen05.ld), apply bias / degree-scaled bias / residual /
// activation, stage the bf16 tile in swizzled shared memory and hand it to a TMA store.  Persistent:
// one CTA per SM, static round-robin over row tiles.  The kernel is HBM-bound by design
// (arithmetic intensity 50-100 flop/B); the tensor pipe only has to stay out of the way.
#pragma once
#include <cstdlib>

#include "tc_common.cuh"

namespace gcb {
namespace tc {

struct TcEpilogue {
  const float* bias = nullptr;      // [N] or null
  const int32_t* rowptr = nullptr;  // when set: bias scaled by deg (add) or [deg > 0] (mean)
  int aggr = GCB_AGGR_ADD;
  const bf16* res = nullptr;        // residual [rows, ld_res] or null
  int ld_res = 0;
  const int32_t* res_idx = nullptr; // when set: row r adds res[res_idx[r]] (a lookup table, e.g. the pre-activated embedding rows)
  int act = -1;                     // enum gcb_act or -1
  int rev = 0;                      // walk the row tiles last-to-first (L2 reuse of a just-written operand)
};

template <int K, int N, bool RES = true>
struct NtShape {
  static constexpr int BM = 128;
  static constexpr int KS = K / 64;  // 64-element (128-byte) K slabs
  static constexpr int NS = N / 64;
  static constexpr int STAGES = (K <= 128) ? 3 : 2;
  static constexpr int W_BYTES = N * K * 2;
  static constexpr int A_STAGE_BYTES = BM * K * 2;
  static constexpr int OUT_BYTES = BM * N * 2;
  static constexpr int BAR_BYTES = 256 + N * 4;  // mbarriers + TMEM slot, then the bias vector
  static constexpr int SMEM_BYTES = 1024 + W_BYTES + STAGES * A_STAGE_BYTES + OUT_BYTES + BAR_BYTES;
  static constexpr int TMEM_COLS = 2 * N < 32 ? 32 : 2 * N;  // two accumulators; power of two for N in {64,128,256}
  // four warps per TMEM lane quarter when the tile is wide enough: the epilogue (tcgen05.ld -> bias /
  // residual / activation -> swizzled staging) is a dependent chain per warp, so it needs warps, not ILP
  // (the residual prefetch registers of a 256-wide tile do not fit next to 16 warps' worth of threads)
  static constexpr int EPI_WARPS = (N == 128 || (N == 256 && !RES)) ? 16 : 8;
  static constexpr int COL_PARTS = EPI_WARPS / 4;
  static constexpr int THREADS = 64 + EPI_WARPS * 32;
  // Store groups: the warps whose columns fall into the same 64-column output slab(s) stage, publish and TMA-store
  // those slabs on their own (named barrier per group, bulk group per leader thread), so that one group's store
  // overlaps the other groups' TMEM drains instead of all 16 warps marching through two CTA-wide barriers per tile.
  static constexpr int COLS_PER_PART = N / COL_PARTS;
  static constexpr int PARTS_PER_GROUP = COLS_PER_PART >= 64 ? 1 : 64 / COLS_PER_PART;
  static constexpr int SLABS_PER_GROUP = COLS_PER_PART >= 64 ? COLS_PER_PART / 64 : 1;
  static constexpr int GROUPS = COL_PARTS / PARTS_PER_GROUP;
  static constexpr int GROUP_THREADS = PARTS_PER_GROUP * 128;
  static_assert(GROUPS * SLABS_PER_GROUP == NS && GROUPS >= 1 && GROUPS <= 8, "store groups must tile the output slabs");
  static_assert(K % 64 == 0 && N % 64 == 0 && N <= 256, "unsupported GEMM shape");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

template <int K, int N, bool RES, bool COALQ = false>
__global__ void __launch_bounds__(NtShape<K, N, RES>::THREADS, 1)
gemm_nt_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmC, int rows, TcEpilogue ep) {
  using S = NtShape<K, N, RES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sA = sW + S::W_BYTES;
  uint8_t* sOut = sA + S::STAGES * S::A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + S::OUT_BYTES);
  uint64_t* full = bars;                  // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + S::STAGES;     // [STAGES]  MMA -> TMA
  uint64_t* tfull = bars + 2 * S::STAGES; // [2]       MMA -> epilogue
  uint64_t* tempty = tfull + 2;           // [2]       epilogue -> MMA
  uint64_t* wfull = tempty + 2;           // [1]       weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (rows + S::BM - 1) / S::BM;

  if (ep.bias) {
    for (int i = threadIdx.x; i < N; i += S::THREADS) sBias[i] = __ldg(ep.bias + i);
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmW); prefetch_tmap(&tmC);
    for (int i = 0; i < S::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], S::EPI_WARPS); }
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, S::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // set-up done (barriers, TMEM, bias).  The weights are immutable while a step runs, so their load may still
  // overlap the previous kernel's tail; everything that touches activations comes after pdl_wait().
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull, S::W_BYTES);
      for (int ks = 0; ks < S::KS; ++ks) tma_load_2d(sW + ks * (N * 128), &tmW, wfull, ks * 64, 0);
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], S::A_STAGE_BYTES);
        uint8_t* dst = sA + stage * S::A_STAGE_BYTES;
        const int ptile = ep.rev ? num_tiles - 1 - tile : tile;
        for (int ks = 0; ks < S::KS; ++ks) tma_load_2d(dst + ks * (S::BM * 128), &tmA, &full[stage], ks * 64, ptile * S::BM);
        if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_bf16(S::BM, N, 0, 0);
      mbar_wait(wfull, 0);
      const uint32_t w_base = smem_u32(sW);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + stage * S::A_STAGE_BYTES);
#pragma unroll
        for (int kk = 0; kk < K / 16; ++kk) {
          const uint32_t ks = kk >> 2, ko = (kk & 3) * 32;
          const uint64_t da = smem_desc_sw128(a_base + ks * (S::BM * 128) + ko, 16, 1024);
          const uint64_t dw = smem_desc_sw128(w_base + ks * (N * 128) + ko, 16, 1024);
          umma_bf16(tmem_base + acc * N, da, dw, idesc, kk > 0 ? 1u : 0u);
        }
        umma_commit(&empty[stage]);  // smem stage reusable once these MMAs have read it
        umma_commit(&tfull[acc]);    // accumulator complete
        if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (EPI_WARPS warps) =====================
    pdl_wait();
    const int ew = warp - 2;              // 0..EPI_WARPS-1
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int half = ew >> 2;             // column part handled by this warp
    const int row_in_tile = q * 32 + lane;
    const int ep_tid = threadIdx.x - 64;
    constexpr int COLS_PER_WARP = N / S::COL_PARTS;
    constexpr int NCHUNK = COLS_PER_WARP / 32;
    const int group = half / S::PARTS_PER_GROUP;               // store group of this warp
    const bool leader = ep_tid == group * S::GROUP_THREADS;    // first thread of the group's first warp
    // Software pipeline across tiles: the residual row chunk and the row degree of the NEXT tile are
    // loaded into registers while the current tile is being finished, so that no global-memory latency
    // sits between "accumulator ready" and "tile staged".
    uint4 res_raw[RES ? NCHUNK : 1][4];
    int rp_lo = 0, rp_hi = 0;  // row pointers of the NEXT tile's row: subtracted when used, a tile later
    const int res_col0 = half * COLS_PER_WARP;
    // COAL (opt-in, see the launcher; the forward atom transform, N = 128): the residual tile is fetched with COALESCED loads -- the group's
    // 256 threads walk its [128 rows x 128 B] slab 16 bytes per lane, a warp covering four full 128-byte row
    // segments -- one tile ahead into registers, parked in the group's (idle) staging slab in the swizzled layout,
    // and every thread then picks up its own row's 64 bytes from shared memory and overwrites them in place with the
    // result.  The thread = TMEM lane = row mapping made the direct loads a 256-byte-stride gather (32 lines per
    // instruction, half of every sector unused): 39 us instead of 25 us for the plain GEMM (profiles/r2_gemm_epilogue.jsonl).
    constexpr bool COAL = COALQ && RES && N == 128 && S::GROUP_THREADS == 256 && S::SLABS_PER_GROUP == 1;
    const int gtid = ep_tid - group * S::GROUP_THREADS;  // thread index inside the store group
    auto coal_prefetch = [&](int tile_row0) {
      if constexpr (COAL) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int qd = gtid + 256 * k, rr = qd >> 3, cc = qd & 7;  // 16-byte chunk cc of slab row rr
          const int gr = tile_row0 + rr;
          res_raw[0][k] = gr < rows ? __ldg(reinterpret_cast<const uint4*>(ep.res + (size_t)(ep.res_idx ? __ldg(ep.res_idx + gr) : gr) * ep.ld_res + group * 64 + cc * 8))
                                    : make_uint4(0u, 0u, 0u, 0u);
        }
      }
    };
    if ((int)blockIdx.x < num_tiles) {
      const int r0 = (ep.rev ? num_tiles - 1 - (int)blockIdx.x : (int)blockIdx.x) * S::BM + row_in_tile;
      if (r0 < rows) {
        if (ep.rowptr) { rp_hi = __ldg(ep.rowptr + r0 + 1); rp_lo = __ldg(ep.rowptr + r0); }
        if constexpr (RES && !COAL) {
          const uint4* rp = reinterpret_cast<const uint4*>(ep.res + (size_t)(ep.res_idx ? __ldg(ep.res_idx + r0) : r0) * ep.ld_res + res_col0);
#pragma unroll
          for (int ci = 0; ci < NCHUNK; ++ci)
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) res_raw[ci][j4] = __ldg(rp + ci * 4 + j4);
        }
      }
      if constexpr (COAL) coal_prefetch((ep.rev ? num_tiles - 1 - (int)blockIdx.x : (int)blockIdx.x) * S::BM);
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int ptile = ep.rev ? num_tiles - 1 - tile : tile;
      const int r = ptile * S::BM + row_in_tile;
      const bool valid = r < rows;
      const int r_next = r + (ep.rev ? -1 : 1) * (int)gridDim.x * S::BM;
      const bool next_valid = (tile + (int)gridDim.x < num_tiles) && r_next < rows;
      float wrow = 1.f;
      if (ep.rowptr) {
        const int deg = rp_hi - rp_lo;
        wrow = ep.aggr == GCB_AGGR_MEAN ? (deg > 0 ? 1.f : 0.f) : (float)deg;
        if (next_valid) { rp_hi = __ldg(ep.rowptr + r_next + 1); rp_lo = __ldg(ep.rowptr + r_next); }
      }
      const int rr_next = next_valid ? (RES && ep.res_idx ? __ldg(ep.res_idx + r_next) : r_next) : 0;
      const uint4* rp_next = RES ? reinterpret_cast<const uint4*>(ep.res + (size_t)rr_next * ep.ld_res + res_col0) : nullptr;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      // this group's previous TMA store must have finished reading the group's staging slabs
      if (leader) tma_store_wait_read();
      named_bar_sync(1 + group, S::GROUP_THREADS);
      // TMEM loads run one 32-column chunk ahead of the arithmetic (tcgen05.wait::ld covers all outstanding loads)
      uint32_t v[2][32];
      const uint32_t tsrc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * N + res_col0;
      tmem_ld_32x32(tsrc, v[0]);
      if constexpr (COAL) {
        uint8_t* gslab = sOut + group * (S::BM * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int qd = gtid + 256 * k, rr = qd >> 3, cc = qd & 7;
          *reinterpret_cast<uint4*>(gslab + rr * 128 + ((cc ^ (rr & 7)) << 4)) = res_raw[0][k];
        }
        named_bar_sync(1 + group, S::GROUP_THREADS);  // the group's residual slab is complete
        if (tile + (int)gridDim.x < num_tiles) coal_prefetch((ep.rev ? num_tiles - 1 - (tile + (int)gridDim.x) : tile + (int)gridDim.x) * S::BM);
      }
#pragma unroll
      for (int ci = 0; ci < NCHUNK; ++ci) {
        const int c0 = res_col0 + ci * 32;
        tmem_ld_wait();
        if (ci + 1 < NCHUNK) tmem_ld_32x32(tsrc + (ci + 1) * 32, v[(ci + 1) & 1]);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[ci & 1][j]);
        if (ep.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + j);
            f[j] = fmaf(wrow, b4.x, f[j]); f[j + 1] = fmaf(wrow, b4.y, f[j + 1]);
            f[j + 2] = fmaf(wrow, b4.z, f[j + 2]); f[j + 3] = fmaf(wrow, b4.w, f[j + 3]);
          }
        }
        if constexpr (COAL) {
          const uint8_t* rs = sOut + (c0 >> 6) * (S::BM * 128) + row_in_tile * 128;
          const int ch0 = (c0 & 63) >> 3;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const uint4 rw = *reinterpret_cast<const uint4*>(rs + (((ch0 + j4) ^ (row_in_tile & 7)) << 4));
            const uint32_t w[4] = {rw.x, rw.y, rw.z, rw.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              f[j4 * 8 + 2 * i] += __uint_as_float(w[i] << 16);
              f[j4 * 8 + 2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
            }
          }
        } else if constexpr (RES) {
          if (valid) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint32_t w[4] = {res_raw[ci][j4].x, res_raw[ci][j4].y, res_raw[ci][j4].z, res_raw[ci][j4].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                f[j4 * 8 + 2 * i] += __uint_as_float(w[i] << 16);
                f[j4 * 8 + 2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
              }
            }
          }
          if (next_valid) {  // refill this chunk's registers for the next tile
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) res_raw[ci][j4] = __ldg(rp_next + ci * 4 + j4);
          }
        }
        if (ep.act >= 0) act_fwd_vec_impl<true, 32>(f, ep.act);
        // stage as bf16 in the 128B-swizzled layout the TMA store expects
        uint8_t* slab = sOut + (c0 >> 6) * (S::BM * 128) + row_in_tile * 128;
        const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          uint4 t;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[j4 * 8 + 2 * i], f[j4 * 8 + 2 * i + 1]);
          *reinterpret_cast<uint4*>(slab + (((chunk0 + j4) ^ (row_in_tile & 7)) << 4)) = t;
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      // publish the group's staged slabs to the async proxy, then the group leader issues their store
      fence_proxy_async_smem();
      named_bar_sync(1 + group, S::GROUP_THREADS);
      if (leader) {
#pragma unroll
        for (int k = 0; k < S::SLABS_PER_GROUP; ++k) {
          const int ns = group * S::SLABS_PER_GROUP + k;
          tma_store_2d(&tmC, sOut + ns * (S::BM * 128), ns * 64, ptile * S::BM);
        }
        tma_store_commit();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (leader) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

template <int K, int N, bool RES>
inline int launch_gemm_nt_tc_res(const bf16* A, int lda, const bf16* W, bf16* C, int ldc, int rows,
                                 const TcEpilogue& ep, cudaStream_t s) {
  using S = NtShape<K, N, RES>;
  static bool attr_set = false;
  if (!attr_set) {
    GCB_CUDA(cudaFuncSetAttribute(gemm_nt_tc_kernel<K, N, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap tmA, tmW, tmC;
  if (make_tmap_bf16(&tmA, A, rows, K, lda, S::BM) || make_tmap_bf16(&tmW, W, N, K, K, N) ||
      make_tmap_bf16(&tmC, C, rows, N, ldc, S::BM))
    return GCB_ERR_CUDA;
  const int num_tiles = (rows + S::BM - 1) / S::BM;
  const int grid = num_tiles < kNumSMs ? num_tiles : kNumSMs;
  // one timing name per template instance (bench.py's roofline names ONE kernel, with that shape's bytes)
  static constexpr const char* kName = (K == 128 && N == 256) ? (RES ? "gemm_nt_tc_k128n256_res" : "gemm_nt_tc_k128n256")
                                       : (K == 256 && N == 128) ? (RES ? "gemm_nt_tc_k256n128_res" : "gemm_nt_tc_k256n128")
                                       : (K == 64 && N == 128) ? (RES ? "gemm_nt_tc_k64n128_res" : "gemm_nt_tc_k64n128")
                                                                : (RES ? "gemm_nt_tc_k128n64_res" : "gemm_nt_tc_k128n64");
  if constexpr (RES && N == 128 && K == 256) {
    // A/B switch: GCB_RES_COAL=1 fetches the residual tile with coalesced loads through the staging slab.  Measured
    // SLOWER on B200 (forward atom GEMM 48.7 vs 42.5 us, train step 1.689 vs 1.658 ms, profiles/r2_fused_ab.jsonl: the
    // extra barrier and shared-memory round trip cost more than the 256-byte-stride loads), so it stays opt-in.
    const char* e = std::getenv("GCB_RES_COAL");
    if (e && e[0] == '1') {
      static bool attr2 = false;
      if (!attr2) {
        GCB_CUDA(cudaFuncSetAttribute(gemm_nt_tc_kernel<K, N, RES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
        attr2 = true;
      }
      GCB_KERNEL(kName, s, launch_pdl(gemm_nt_tc_kernel<K, N, RES, true>, grid, S::THREADS, S::SMEM_BYTES, s, tmA, tmW, tmC, rows, ep));
      return GCB_OK;
    }
  }
  GCB_KERNEL(kName, s, launch_pdl(gemm_nt_tc_kernel<K, N, RES>, grid, S::THREADS, S::SMEM_BYTES, s, tmA, tmW, tmC, rows, ep));
  return GCB_OK;
}

// The residual is a compile-time variant: without it the epilogue keeps no prefetch registers.
template <int K, int N>
inline int launch_gemm_nt_tc_impl(const bf16* A, int lda, const bf16* W, bf16* C, int ldc, int rows,
                                  const TcEpilogue& ep, cudaStream_t s) {
  return ep.res ? launch_gemm_nt_tc_res<K, N, true>(A, lda, W, C, ldc, rows, ep, s)
                : launch_gemm_nt_tc_res<K, N, false>(A, lda, W, C, ldc, rows, ep, s);
}

// -------------------------------------------------------------------------------------------------
// Streamed-weight variant for wide models (embedding_dim 256: [2D, D] weights are 256 KB and cannot stay
// in shared memory).  Classic tiling: a CTA owns a [128 x 256] output tile at a time, A [128 x 64] and
// W [256 x 64] K-slabs stream through a 3-stage TMA ring (weights come from L2, where all of W lives),
// the accumulator is double-buffered in TMEM, the epilogue is the same as above (bias / degree-scaled
// bias / residual / activation, swizzled staging, TMA store).  Persistent, tiles enumerated row-tile
// major so that the column tiles of one row tile reuse its A slabs out of L2.
struct NtStream {
  static constexpr int BM = 128, BN = 256, BK = 64;
  static constexpr int STAGES = 3;
  static constexpr int A_BYTES = BM * BK * 2, W_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + W_BYTES;  // 16 + 32 KB
  static constexpr int OUT_BYTES = BM * BN * 2;
  static constexpr int MAX_N = 1024;                               // bias vector kept in shared memory
  static constexpr int BAR_BYTES = 256 + MAX_N * 4;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + OUT_BYTES + BAR_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int EPI_WARPS = 8, COL_PARTS = 2;
  static constexpr int THREADS = 64 + EPI_WARPS * 32;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

__global__ void __launch_bounds__(NtStream::THREADS, 1)
gemm_nt_tcs_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                   const __grid_constant__ CUtensorMap tmC, int rows, int N, int K, TcEpilogue ep) {
  using S = NtStream;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sStage = smem;
  uint8_t* sOut = sStage + S::STAGES * S::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + S::OUT_BYTES);
  uint64_t* full = bars;                  // [STAGES]
  uint64_t* empty = bars + S::STAGES;     // [STAGES]
  uint64_t* tfull = bars + 2 * S::STAGES; // [2]
  uint64_t* tempty = tfull + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (rows + S::BM - 1) / S::BM, num_n = N / S::BN, num_tiles = num_m * num_n, ksteps = K / S::BK;

  if (ep.bias) {
    for (int i = threadIdx.x; i < N; i += S::THREADS) sBias[i] = __ldg(ep.bias + i);
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmW); prefetch_tmap(&tmC);
    for (int i = 0; i < S::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], S::EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, S::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // set-up done (barriers, TMEM, bias): from here on the kernel touches activations of the previous kernel
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * S::BM, n0 = (tile % num_n) * S::BN;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], S::STAGE_BYTES);
          uint8_t* a = sStage + stage * S::STAGE_BYTES;
          tma_load_2d(a, &tmA, &full[stage], ks * S::BK, m0);
          tma_load_2d(a + S::A_BYTES, &tmW, &full[stage], ks * S::BK, n0);
          if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_bf16(S::BM, S::BN, 0, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sStage + stage * S::STAGE_BYTES);
          const uint32_t w_base = a_base + S::A_BYTES;
#pragma unroll
          for (int kk = 0; kk < S::BK / 16; ++kk) {
            const uint64_t da = smem_desc_sw128(a_base + kk * 32, 16, 1024);
            const uint64_t dw = smem_desc_sw128(w_base + kk * 32, 16, 1024);
            umma_bf16(tmem_base + acc * S::BN, da, dw, idesc, (ks > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int ew = warp - 2;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int row_in_tile = q * 32 + lane;
    const int ep_tid = threadIdx.x - 64;
    constexpr int COLS_PER_WARP = S::BN / S::COL_PARTS;  // 128
    constexpr int NCHUNK = COLS_PER_WARP / 32;           // 4
    // two store groups (the four warps of a column part own two 64-column slabs): see NtShape
    const int group = half;
    const bool leader = ep_tid == group * 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * S::BM, n0 = (tile % num_n) * S::BN;
      const int r = m0 + row_in_tile;
      const bool valid = r < rows;
      const int rres = (ep.res_idx && valid) ? __ldg(ep.res_idx + r) : r;  // residual row (identity or table lookup)
      float wrow = 1.f;
      if (ep.rowptr) {
        const int deg = valid ? __ldg(ep.rowptr + r + 1) - __ldg(ep.rowptr + r) : 0;
        wrow = ep.aggr == GCB_AGGR_MEAN ? (deg > 0 ? 1.f : 0.f) : (float)deg;
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (leader) tma_store_wait_read();  // the group's previous store has finished reading its staging slabs
      named_bar_sync(1 + group, 128);
#pragma unroll
      for (int ci = 0; ci < NCHUNK; ++ci) {
        const int c0 = half * COLS_PER_WARP + ci * 32;  // column inside the tile
        uint4 res_raw[4];
        if (ep.res && valid) {
          const uint4* rp = reinterpret_cast<const uint4*>(ep.res + (size_t)rres * ep.ld_res + n0 + c0);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) res_raw[j4] = __ldg(rp + j4);
        }
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * S::BN + c0, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (ep.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sBias + n0 + c0 + j);
            f[j] = fmaf(wrow, b4.x, f[j]); f[j + 1] = fmaf(wrow, b4.y, f[j + 1]);
            f[j + 2] = fmaf(wrow, b4.z, f[j + 2]); f[j + 3] = fmaf(wrow, b4.w, f[j + 3]);
          }
        }
        if (ep.res && valid) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const uint32_t w[4] = {res_raw[j4].x, res_raw[j4].y, res_raw[j4].z, res_raw[j4].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              f[j4 * 8 + 2 * i] += __uint_as_float(w[i] << 16);
              f[j4 * 8 + 2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
            }
          }
        }
        if (ep.act >= 0) act_fwd_vec_impl<true, 32>(f, ep.act);
        uint8_t* slab = sOut + (c0 >> 6) * (S::BM * 128) + row_in_tile * 128;
        const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          uint4 t;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[j4 * 8 + 2 * i], f[j4 * 8 + 2 * i + 1]);
          *reinterpret_cast<uint4*>(slab + (((chunk0 + j4) ^ (row_in_tile & 7)) << 4)) = t;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      fence_proxy_async_smem();
      named_bar_sync(1 + group, 128);
      if (leader) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int ns = group * 2 + k;
          tma_store_2d(&tmC, sOut + ns * (S::BM * 128), n0 + ns * 64, m0);
        }
        tma_store_commit();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (leader) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

inline int launch_gemm_nt_tcs(const bf16* A, int lda, const bf16* W, bf16* C, int ldc, int rows, int N, int K,
                              const TcEpilogue& ep, cudaStream_t s) {
  using S = NtStream;
  static bool attr_set = false;
  if (!attr_set) {
    GCB_CUDA(cudaFuncSetAttribute(gemm_nt_tcs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap tmA, tmW, tmC;
  if (make_tmap_bf16(&tmA, A, rows, K, lda, S::BM) || make_tmap_bf16(&tmW, W, N, K, K, S::BN) ||
      make_tmap_bf16(&tmC, C, rows, N, ldc, S::BM))
    return GCB_ERR_CUDA;
  const int num_tiles = ((rows + S::BM - 1) / S::BM) * (N / S::BN);
  const int grid = num_tiles < kNumSMs ? num_tiles : kNumSMs;
  GCB_KERNEL("gemm_nt_tcs_kernel", s, launch_pdl(gemm_nt_tcs_kernel, grid, S::THREADS, S::SMEM_BYTES, s, tmA, tmW, tmC, rows, N, K, ep));
  return GCB_OK;
}

// Returns GCB_ERR_UNSUPPORTED when the shape has no tcgen05 instantiation (caller falls back to
// the CUDA-core GEMM).
inline int launch_gemm_nt_tc(const bf16* A, int lda, const bf16* W, bf16* C, int ldc, int rows, int N, int K,
                             const TcEpilogue& ep, cudaStream_t s) {
  if (rows <= 0) return GCB_OK;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(C)) & 15) return GCB_ERR_UNSUPPORTED;
  if (lda % 8 || ldc % 8) return GCB_ERR_UNSUPPORTED;
  if (K == 128 && N == 256) return launch_gemm_nt_tc_impl<128, 256>(A, lda, W, C, ldc, rows, ep, s);
  if (K == 256 && N == 128) return launch_gemm_nt_tc_impl<256, 128>(A, lda, W, C, ldc, rows, ep, s);
  if (K == 64 && N == 128) return launch_gemm_nt_tc_impl<64, 128>(A, lda, W, C, ldc, rows, ep, s);
  if (K == 128 && N == 64) return launch_gemm_nt_tc_impl<128, 64>(A, lda, W, C, ldc, rows, ep, s);
  // wide models: weights streamed from L2 (any K that is a multiple of 64, any N that is a multiple of 256)
  if (K % NtStream::BK == 0 && N % NtStream::BN == 0 && N <= NtStream::MAX_N && K >= NtStream::BK)
    return launch_gemm_nt_tcs(A, lda, W, C, ldc, rows, N, K, ep, s);
  return GCB_ERR_UNSUPPORTED;
}

}  // namespace tc
}  // namespace gcb

// =================================================================================================
// Weight-gradient ("TN") GEMM on tcgen05:   dW[NN, KK] = sum_r A[r, NN]^T B[r, KK]   (+ column sums)
//
// The reduction runs over graph rows, so both operands are MN-major for the tensor core: a TMA tile
// of 128 rows x 64 columns (128-byte swizzled rows) is read as "K = row, M/N = column".  Each CTA
// owns a contiguous set of 128-row tiles and keeps its partial dW in TMEM for the whole kernel
// (NN/128 accumulators of KK fp32 columns); partials are written once per CTA and summed in CTA
// order by reduce_partials_kernel -> deterministic.  Warps 2-5 compute the weighted column sums
// db[n] = sum_r w(r) A[r, n] (bias gradients) from the same shared-memory tile while the MMAs run.
// =================================================================================================
namespace gcb {
namespace tc {

template <int NN, int KK>
struct TnShape {
  static constexpr int BR = 128;                 // rows (MMA K extent) per tile
  static constexpr int AS = NN / 64, BS = KK / 64;
  static constexpr int MB = NN / 128;            // 128-row output blocks
  static constexpr int STAGES = 2;
  static constexpr int A_BYTES = BR * NN * 2, B_BYTES = BR * KK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 128 + 512 + 4096;  // + barriers, row weights, bias reduction
  static constexpr int TMEM_COLS = MB * KK;      // 256 for both D=128 shapes
  static constexpr int THREADS = 192;
  static_assert(NN % 128 == 0 && KK % 64 == 0 && KK <= 256 && MB * KK <= 512, "unsupported TN shape");
  static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS >= 32, "TMEM columns must be a power of two");
};

// Column-blocked form (BLK; wide models, embedding_dim 256: dW is [256 x 512] / [512 x 256] and does not fit the 512
// TMEM columns): the grid is (row ranges) x (blocks_m x blocks_n) CTAs, CTA (ci, mb, nb) owns the [128 x 256] block
// (mb, nb) of its row range's partial dW -- the same kernel on column windows of A and B (TMA coordinates), written into
// a [NN_total x KK_total] partial with that row stride.  The CTAs of one row range walk the same row tiles, so the
// windows they re-read (A blocks_n times, B blocks_m times) meet in L2.  Bias sums: the nb == 0 blocks whose A window
// lies inside the first `bias_cols` columns.
struct TnBlocks {
  int blocks_m = 1, blocks_n = 1;  // dW blocks of 128 rows x 256 columns
  int nn_total = 0, kk_total = 0;  // full dW extent (row stride of the partial = kk_total)
  int bias_cols = 0;               // leading columns of A that carry a bias gradient
};

template <int NN, int KK, bool BLK = false>
__global__ void __launch_bounds__(192, 1)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int rows,
                  float* __restrict__ partial, int want_bias_i, const int32_t* __restrict__ rowptr,
                  int aggr, int rev, int acc_partial, TnBlocks blk) {
  using S = TnShape<NN, KK>;
  static_assert(!BLK || (NN == 128 && KK == 256), "the blocked form runs on [128 x 256] blocks");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sT = smem;  // [STAGES][A tile | B tile]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sT + S::STAGES * S::STAGE_BYTES);
  uint64_t* full = bars;               // [STAGES]
  uint64_t* empty = bars + S::STAGES;  // [STAGES]  (MMA commit + colsum warps)
  uint64_t* done = empty + S::STAGES;  // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  float* w_s = reinterpret_cast<float*>(bars + 16);  // [128] row weights of the current tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (rows + S::BR - 1) / S::BR;
  // contiguous tile range of this CTA (fixed partition -> fixed summation order)
  const int nblk = BLK ? blk.blocks_m * blk.blocks_n : 1;
  const int ci = BLK ? (int)blockIdx.x / nblk : (int)blockIdx.x;      // row range
  const int bi = BLK ? (int)blockIdx.x % nblk : 0;
  const int mb_blk = BLK ? bi / blk.blocks_n : 0, nb_blk = BLK ? bi % blk.blocks_n : 0;
  const int a_col0 = mb_blk * 128, b_col0 = nb_blk * 256;              // column windows of A and B
  const int n_ranges = BLK ? (int)gridDim.x / nblk : (int)gridDim.x;
  const int per = (num_tiles + n_ranges - 1) / n_ranges;
  const int t_begin = ci * per;
  const int t_end = min(num_tiles, t_begin + per);
  // blocked: every row range's partial carries NN_total bias slots; a block computes the 128 of its A window when it is an
  // nb == 0 block inside the bias columns, writes zeros when it is an nb == 0 block outside them, and nothing otherwise
  const bool bias_owner = !BLK || nb_blk == 0;
  const bool want_bias = want_bias_i != 0 && bias_owner && (!BLK || a_col0 < blk.bias_cols);
  const int out_nn = BLK ? blk.nn_total : NN, out_kk = BLK ? blk.kk_total : KK;
  // per-CTA (row range) partial block: [NN*KK weight-gradient partial | NN bias sums]
  float* const cta_partial = partial + (size_t)ci * ((size_t)out_nn * out_kk + (want_bias_i != 0 ? out_nn : 0));

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int i = 0; i < S::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], want_bias ? 5 : 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, S::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // set-up done (barriers, TMEM, bias): from here on the kernel touches activations of the previous kernel
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], S::STAGE_BYTES);
        uint8_t* a = sT + stage * S::STAGE_BYTES;
        uint8_t* b = a + S::A_BYTES;
        const int ptile = rev ? num_tiles - 1 - tile : tile;
        for (int k = 0; k < S::AS; ++k) tma_load_2d(a + k * (S::BR * 128), &tmA, &full[stage], a_col0 + k * 64, ptile * S::BR);
        for (int k = 0; k < S::BS; ++k) tma_load_2d(b + k * (S::BR * 128), &tmB, &full[stage], b_col0 + k * 64, ptile * S::BR);
        if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_bf16(128, KK, 1, 1);  // both operands MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sT + stage * S::STAGE_BYTES);
        const uint32_t b_base = a_base + S::A_BYTES;
#pragma unroll
        for (int kk = 0; kk < S::BR / 16; ++kk) {
          const uint64_t db = smem_desc_sw128(b_base + kk * 2048, S::BR * 128, 1024);
#pragma unroll
          for (int mb = 0; mb < S::MB; ++mb) {
            const uint64_t da = smem_desc_sw128(a_base + mb * 2 * (S::BR * 128) + kk * 2048, S::BR * 128, 1024);
            umma_bf16(tmem_base + mb * KK, da, db, idesc, (tile > t_begin || kk > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty[stage]);
        if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(done);
    }
  } else {
    // ---- warps 2..5: weighted column sums during the main loop, then the TMEM -> global epilogue ----
    // Only the first 128 columns of A carry a bias gradient (all of dZA; the dP half of [dP|dQ]).
    // Thread (cl, rl): 16-byte chunk cl (8 columns) of rows rl, rl+8, ... ; the 8 row lanes are
    // combined once, in lane order, after the last tile.
    const int t = threadIdx.x - 64;  // 0..127
    const int cl = t & 15, rl = t >> 4;
    float bsum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bsum[i] = 0.f;
    if (want_bias) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int r = (rev ? num_tiles - 1 - tile : tile) * S::BR + t;
        float w = 0.f;
        if (r < rows) {
          if (rowptr) {
            const int deg = __ldg(rowptr + r + 1) - __ldg(rowptr + r);
            w = aggr == GCB_AGGR_MEAN ? (deg > 0 ? 1.f : 0.f) : (float)deg;
          } else w = 1.f;
        }
        named_bar_sync(1, 128);  // previous tile's readers are done with w_s
        w_s[t] = w;
        mbar_wait(&full[stage], phase);
        named_bar_sync(1, 128);
        const uint8_t* slab = sT + stage * S::STAGE_BYTES + (cl >> 3) * (S::BR * 128);
        const int chunk = cl & 7;
#pragma unroll 4
        for (int rr = rl; rr < S::BR; rr += 8) {
          const uint4 raw = *reinterpret_cast<const uint4*>(slab + rr * 128 + ((chunk ^ (rr & 7)) << 4));
          const float wr = w_s[rr];
          const uint32_t wd[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            bsum[2 * i] = fmaf(wr, __uint_as_float(wd[i] << 16), bsum[2 * i]);
            bsum[2 * i + 1] = fmaf(wr, __uint_as_float(wd[i] & 0xFFFF0000u), bsum[2 * i + 1]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
      }
      // combine the 8 row lanes in order (red_s: [8][128] floats, after w_s)
      float* red_s = w_s + 128;
#pragma unroll
      for (int i = 0; i < 8; ++i) red_s[rl * 128 + cl * 8 + i] = bsum[i];
      named_bar_sync(1, 128);
      float tot = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) tot += red_s[k * 128 + t];
      float* bslot = cta_partial + (size_t)out_nn * out_kk + a_col0;
      bslot[t] = acc_partial ? bslot[t] + tot : tot;
      if constexpr (!BLK) {
#pragma unroll
        for (int i = 1; i < NN / 128; ++i) bslot[i * 128 + t] = 0.f;
      }
    } else if (BLK && want_bias_i != 0 && bias_owner) {
      cta_partial[(size_t)out_nn * out_kk + a_col0 + t] = 0.f;  // an A window without a bias gradient (the dQ half)
    }
    mbar_wait(done, 0);
    tc_fence_after();
    const int q = warp & 3;
    float* out = cta_partial + (size_t)a_col0 * out_kk + b_col0;
#pragma unroll 1
    for (int mb = 0; mb < S::MB; ++mb) {
      const int m = mb * 128 + q * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < KK; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + mb * KK + c0, v);
        tmem_ld_wait();
        float4* o = reinterpret_cast<float4*>(out + (size_t)m * out_kk + c0);
        if (acc_partial) {  // keep summing this CTA's partial across launches (layers); one reduction at the end
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 old = o[j];
            o[j] = make_float4(old.x + __uint_as_float(v[4 * j]), old.y + __uint_as_float(v[4 * j + 1]),
                               old.z + __uint_as_float(v[4 * j + 2]), old.w + __uint_as_float(v[4 * j + 3]));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

constexpr int kTnMaxCtas = kNumSMs;

template <int NN, int KK>
inline int launch_gemm_tn_tc_impl(const bf16* A, int lda, const bf16* B, int ldb, int rows, float* partial,
                                  bool want_bias, const int32_t* rowptr, int aggr, int rev, int* grid_out, cudaStream_t s,
                                  int acc_partial) {
  using S = TnShape<NN, KK>;
  static bool attr_set = false;
  if (!attr_set) {
    GCB_CUDA(cudaFuncSetAttribute(gemm_tn_tc_kernel<NN, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap tmA, tmB;
  if (make_tmap_bf16(&tmA, A, rows, NN, lda, S::BR) || make_tmap_bf16(&tmB, B, rows, KK, ldb, S::BR)) return GCB_ERR_CUDA;
  const int num_tiles = (rows + S::BR - 1) / S::BR;
  // A/B switch GCB_TN_CTAS: fewer weight-gradient CTAs leave whole SMs to the kernels of the main chain they overlap
  // (40-56 of 148 CTAs measured best on B200 at batch 8192: train step 1.547 -> 1.504 ms; 32 makes the side stream the critical path; gpurun_out/r2_ab_s.jsonl, r2_ab_t.jsonl)
  static const int tn_cap = [] {
    const char* e = std::getenv("GCB_TN_CTAS");
    const int n = e ? std::atoi(e) : 48;
    return n < 1 ? 1 : (n > kTnMaxCtas ? kTnMaxCtas : n);
  }();
  int grid = num_tiles < tn_cap ? num_tiles : tn_cap;
  const int per = (num_tiles + grid - 1) / grid;
  grid = (num_tiles + per - 1) / per;  // every CTA owns at least one tile
  *grid_out = grid;
  GCB_KERNEL(NN == 128 ? "gemm_tn_tc_n128k256" : "gemm_tn_tc_n256k128", s, launch_pdl(gemm_tn_tc_kernel<NN, KK>, grid, S::THREADS, S::SMEM_BYTES, s,
      tmA, tmB, rows, partial, want_bias ? 1 : 0, rowptr, aggr, rev, acc_partial, TnBlocks()));
  return GCB_OK;
}

// Wide shapes (NN a multiple of 128, KK a multiple of 256, beyond one [128 x 256] block): the column-blocked form.
// *grid_out = number of ROW RANGES (= partial blocks the caller reduces, each [NN*KK (+ NN)] floats).
inline int launch_gemm_tn_tc_blocked(const bf16* A, int lda, const bf16* B, int ldb, int rows, int NN, int KK, float* partial,
                                     bool want_bias, int bias_cols, const int32_t* rowptr, int aggr, int rev, int* grid_out,
                                     cudaStream_t s) {
  using S = TnShape<128, 256>;
  static bool attr_set = false;
  if (!attr_set) {
    GCB_CUDA(cudaFuncSetAttribute(gemm_tn_tc_kernel<128, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap tmA, tmB;
  if (make_tmap_bf16(&tmA, A, rows, NN, lda, S::BR) || make_tmap_bf16(&tmB, B, rows, KK, ldb, S::BR)) return GCB_ERR_CUDA;
  TnBlocks blk;
  blk.blocks_m = NN / 128; blk.blocks_n = KK / 256; blk.nn_total = NN; blk.kk_total = KK; blk.bias_cols = bias_cols;
  const int nblk = blk.blocks_m * blk.blocks_n;
  const int num_tiles = (rows + S::BR - 1) / S::BR;
  static const int tn_cap = [] {
    const char* e = std::getenv("GCB_TN_CTAS_WIDE");
    const int n = e ? std::atoi(e) : 144;
    return n < 1 ? 1 : (n > kTnMaxCtas ? kTnMaxCtas : n);
  }();
  int ranges = tn_cap / nblk;
  if (ranges < 1) ranges = 1;
  if (ranges > num_tiles) ranges = num_tiles;
  const int per = (num_tiles + ranges - 1) / ranges;
  ranges = (num_tiles + per - 1) / per;  // every row range owns at least one tile
  *grid_out = ranges;
  GCB_KERNEL("gemm_tn_tc_blocked", s, launch_pdl(gemm_tn_tc_kernel<128, 256, true>, ranges * nblk, S::THREADS, S::SMEM_BYTES, s,
      tmA, tmB, rows, partial, want_bias ? 1 : 0, rowptr, aggr, rev, 0, blk));
  return GCB_OK;
}

// Writes per-CTA partial blocks [grid][NN*KK (+ NN bias sums)] and reports the CTA count.
inline int launch_gemm_tn_tc(const bf16* A, int lda, const bf16* B, int ldb, int rows, int NN, int KK, float* partial,
                             bool want_bias, const int32_t* rowptr, int aggr, int rev, int* grid_out,
                             cudaStream_t s, int acc_partial = 0) {
  if (rows < 128) return GCB_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) return GCB_ERR_UNSUPPORTED;
  if (lda % 8 || ldb % 8) return GCB_ERR_UNSUPPORTED;
  if (NN == 128 && KK == 256) return launch_gemm_tn_tc_impl<128, 256>(A, lda, B, ldb, rows, partial, want_bias, rowptr, aggr, rev, grid_out, s, acc_partial);
  if (NN == 256 && KK == 128) return launch_gemm_tn_tc_impl<256, 128>(A, lda, B, ldb, rows, partial, want_bias, rowptr, aggr, rev, grid_out, s, acc_partial);
  // wide models (embedding_dim 256: [256 x 512] and [512 x 256]); the bias gradient lives in the first min(NN, KK) columns
  // of A (all of dZA, the dP half of [dP|dQ]) -- the same convention as the two shapes above
  static const bool wide_off = std::getenv("GCB_NO_TN_WIDE") && std::getenv("GCB_NO_TN_WIDE")[0] == '1';  // A/B: CUDA-core fallback
  if (!wide_off && !acc_partial && NN % 128 == 0 && KK % 256 == 0 && (NN / 128) * (KK / 256) <= kTnMaxCtas && (int64_t)NN * KK <= 512 * 512)
    return launch_gemm_tn_tc_blocked(A, lda, B, ldb, rows, NN, KK, partial, want_bias, NN < KK ? NN : KK, rowptr, aggr, rev, grid_out, s);
  return GCB_ERR_UNSUPPORTED;
}

}  // namespace tc
}  // namespace gcb
