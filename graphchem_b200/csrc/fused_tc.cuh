// Fused "linear transform + message step" kernels on tcgen05 (bf16, D = 128).
//
// Both message-passing steps that follow a [rows,128] x [128,256] product consume the product
// row-locally: every neighbour a row needs lives in the same molecule.  The packed batch therefore
// carries CLOSED row tiles (<= 128 rows, molecules never split, include/graphchem_b200.h
// GCB_H_TILE_PTR), and one CTA can do the whole step for a tile without the [rows,256] intermediate
// ever reaching HBM:
//
//   TMA (A tile, 128 x 128 bf16)  ->  tcgen05.mma into TMEM (128 x 256 fp32, double buffered)
//   -> 16 epilogue warps: tcgen05.ld, (+bias), round to bf16, stage the tile in shared memory
//   -> the same warps run the row phase out of shared memory (16 lanes per row, 16-byte chunks)
//
//   BOND_FWD  (reference gcn.py:163-164, EdgeConv on the bond matrix + act):
//       [P|Q] = B_{l-1} [U;V]^T + [b|0];   B_l[r] = act(P[r] + max_{e->r} Q[src e]),  act(0) without in-edges;
//       training also records the per-edge tie bits / per-channel tie counts of the max.
//   ATOM_BWD  (autograd of gcn.py:172-173, GeneralConv + act):
//       dS = dZA_l [W_m|W_e];   dZA_{l-1}[j] = (dZA_l[j] + sum_{e: src e = j} dS[dst e, :D]) * act'(A_{l-1}[j]);
//       the second half dS[:, D:] (consumed across molecules by the bond backward) is written out.
//
// The arithmetic (rounding points, summation order, tie handling) is exactly that of the unfused
// pipeline gemm_nt_tc_kernel -> bond_update_kernel / atom_bwd_gather_kernel, so both paths give
// bit-identical results (tests/test_fused_gpu.py).
#pragma once
#include "gemm_tc.cuh"
#include "kernels_v2.cuh"
#include "tile_rows.cuh"

namespace gcb {
namespace tc {

enum FusedMode { FUSED_BOND_FWD = 0, FUSED_ATOM_BWD = 1 };

struct FusedArgs {
  const int32_t* tile_ptr = nullptr;  // [n_tiles + 1]
  int n_tiles = 0;
  int rows = 0;                       // live rows (M for the bond step, N for the atom step)
  int act = 0;
  // CSR of the row phase: dst-sorted (in_*) for BOND_FWD, src-sorted (out_*) for ATOM_BWD
  const int32_t* ptr = nullptr;       // [rows + 1]
  const int32_t* nbr = nullptr;       // [E] neighbour row of every CSR entry
  const int4* nbr4 = nullptr;         // [rows] ELL-4 view of nbr
  // BOND_FWD
  const float* bias = nullptr;        // [256]: b_bond | 0
  bf16* Bout = nullptr;               // [M, 128]
  uint8_t* tiemask = nullptr;         // [E, 16] or null (inference)
  uint8_t* tiecnt = nullptr;          // [M, 128] or null
  // ATOM_BWD
  const bf16* dZA = nullptr;          // [N, 128] (also the GEMM operand)
  const bf16* A_prev = nullptr;       // [N, 128]
  bf16* dZA_prev = nullptr;           // [N, 128]
  bf16* dS = nullptr;                 // [N, 256]; only columns 128.. are written
};

struct FusedShape {
  static constexpr int D = 128, K = 128, N = 256, BM = 128;
  static constexpr int KS = K / 64;
  static constexpr int STAGES = 2;
  static constexpr int W_BYTES = N * K * 2;         // 64 KB, resident
  static constexpr int A_STAGE_BYTES = BM * K * 2;  // 32 KB
  static constexpr int ROW_BYTES = N * 2;           // one staged result row: 512 B = 32 sixteen-byte chunks
  static constexpr int T_BYTES = BM * ROW_BYTES;    // 64 KB staged bf16 result tile
  static constexpr int BAR_BYTES = 256 + N * 4;
  static constexpr int SMEM_BYTES = 1024 + W_BYTES + STAGES * A_STAGE_BYTES + T_BYTES + BAR_BYTES;
  static constexpr int TMEM_COLS = 512;             // two 256-column accumulators
  static constexpr int EPI_WARPS = 16;
  static constexpr int EPI_THREADS = EPI_WARPS * 32;
  static constexpr int THREADS = 64 + EPI_THREADS;
  static constexpr int PASSES = BM / (EPI_THREADS / 16);  // 16 lanes per row -> 32 rows per pass
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

// Staged tile addressing: row `rr`, 16-byte chunk `k` (0..31); the XOR keeps both the per-row
// TMEM drain (32 rows, one chunk) and the row phase (one row, 16 chunks) free of avoidable conflicts.
__device__ __forceinline__ uint32_t fused_chunk_off(int rr, int k) {
  return (uint32_t)rr * FusedShape::ROW_BYTES + (uint32_t)((k ^ (rr & 7)) << 4);
}

template <int MODE, bool TIES>
__global__ void __launch_bounds__(FusedShape::THREADS, 1)
fused_rows_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, FusedArgs fa) {
  using S = FusedShape;
  using R = v2::Raw<bf16>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sA = sW + S::W_BYTES;
  uint8_t* sT = sA + S::STAGES * S::A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sT + S::T_BYTES);
  uint64_t* full = bars;                   // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + S::STAGES;      // [STAGES]  MMA -> TMA
  uint64_t* tfull = bars + 2 * S::STAGES;  // [2]       MMA -> epilogue
  uint64_t* tempty = tfull + 2;            // [2]       epilogue -> MMA
  uint64_t* wfull = tempty + 2;            // [1]       weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = fa.n_tiles;

  if (MODE == FUSED_BOND_FWD) {
    for (int i = threadIdx.x; i < S::N; i += S::THREADS) sBias[i] = fa.bias ? __ldg(fa.bias + i) : 0.f;
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmW);
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
  const uint32_t sTa = smem_u32(sT);  // staged tile, 32-bit shared address (explicit ld/st.shared)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull, S::W_BYTES);
      for (int ks = 0; ks < S::KS; ++ks) tma_load_2d(sW + ks * (S::N * 128), &tmW, wfull, ks * 64, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int r0 = __ldg(fa.tile_ptr + tile);
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], S::A_STAGE_BYTES);
        uint8_t* dst = sA + stage * S::A_STAGE_BYTES;
        // 128 rows from r0: rows past the tile end belong to the next tile (ignored by the row
        // phase), rows past the matrix end are zero-filled by TMA
        for (int ks = 0; ks < S::KS; ++ks) tma_load_2d(dst + ks * (S::BM * 128), &tmA, &full[stage], ks * 64, r0);
        if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_bf16(S::BM, S::N, 0, 0);
      mbar_wait(wfull, 0);
      const uint32_t w_base = smem_u32(sW);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + stage * S::A_STAGE_BYTES);
#pragma unroll
        for (int kk = 0; kk < S::K / 16; ++kk) {
          const uint32_t ks = kk >> 2, ko = (kk & 3) * 32;
          const uint64_t da = smem_desc_sw128(a_base + ks * (S::BM * 128) + ko, 16, 1024);
          const uint64_t dw = smem_desc_sw128(w_base + ks * (S::N * 128) + ko, 16, 1024);
          umma_bf16(tmem_base + acc * S::N, da, dw, idesc, kk > 0 ? 1u : 0u);
        }
        umma_commit(&empty[stage]);
        umma_commit(&tfull[acc]);
        if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue + row phase (16 warps) =====================
    const int ew = warp - 2;           // 0..15
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int part = ew >> 2;          // 64-column part drained by this warp
    const int drain_row = q * 32 + lane;
    const int ep_tid = threadIdx.x - 64;
    const int grp = ep_tid >> 4;       // row slot within a pass (0..31); a warp owns two adjacent rows
    const int lig = ep_tid & 15;       // 16-byte chunk of the 128-channel row
    const int c = lig * 8;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int r0 = __ldg(fa.tile_ptr + tile);
      const int r1 = min(__ldg(fa.tile_ptr + tile + 1), fa.rows);
      // row-phase indices of all passes are fetched before the accumulator is waited for
      int4 n4[S::PASSES];
      int e0[S::PASSES], e1[S::PASSES];
#pragma unroll
      for (int ps = 0; ps < S::PASSES; ++ps) {
        const int r = r0 + ps * 32 + grp;
        n4[ps] = make_int4(-1, -1, -1, -1);
        e0[ps] = e1[ps] = 0;
        if (r < r1) {
          n4[ps] = __ldg(fa.nbr4 + r);
          e0[ps] = __ldg(fa.ptr + r);
          e1[ps] = __ldg(fa.ptr + r + 1);
        }
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      // ---- drain: TMEM -> (+bias) -> bf16 -> staged tile ----
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int c0 = part * 64 + ci * 32;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * S::N + c0, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (MODE == FUSED_BOND_FWD) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sBias + c0 + j);
            f[j] = fmaf(1.f, b4.x, f[j]); f[j + 1] = fmaf(1.f, b4.y, f[j + 1]);
            f[j + 2] = fmaf(1.f, b4.z, f[j + 2]); f[j + 3] = fmaf(1.f, b4.w, f[j + 3]);
          }
        }
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          uint4 t;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[j4 * 8 + 2 * i], f[j4 * 8 + 2 * i + 1]);
          sts128(sTa + fused_chunk_off(drain_row, (c0 >> 3) + j4), t);
        }
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      named_bar_sync(1, S::EPI_THREADS);  // staged tile complete

      // ---- row phase ----
      if (MODE == FUSED_BOND_FWD) {
#pragma unroll
        for (int ps = 0; ps < S::PASSES; ++ps) {
          const int rr = ps * 32 + grp;
          const int r = r0 + rr;
          if (r >= r1) continue;
          const int sidx[4] = {n4[ps].x, n4[ps].y, n4[ps].z, n4[ps].w};
          const uint4 praw = lds128(sTa + fused_chunk_off(rr, lig));
          uint4 qv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (sidx[u] >= 0) qv[u] = lds128(sTa + fused_chunk_off((sidx[u] - r0) & 127, 16 + lig));
          uint4 mraw = R::neg_inf();
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (sidx[u] >= 0) mraw = R::vmax(mraw, qv[u]);
          for (int e = e0[ps] + 4; e < e1[ps]; ++e)
            mraw = R::vmax(mraw, lds128(sTa + fused_chunk_off((__ldg(fa.nbr + e) - r0) & 127, 16 + lig)));
          if (TIES && e1[ps] > e0[ps]) {
            uint32_t cnt_lo = 0u, cnt_hi = 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (sidx[u] >= 0) fa.tiemask[(size_t)(e0[ps] + u) * 16 + lig] = (uint8_t)tile::tie_bits(qv[u], mraw, cnt_lo, cnt_hi);
            }
            for (int e = e0[ps] + 4; e < e1[ps]; ++e) {
              const uint4 qe = lds128(sTa + fused_chunk_off((__ldg(fa.nbr + e) - r0) & 127, 16 + lig));
              fa.tiemask[(size_t)e * 16 + lig] = (uint8_t)tile::tie_bits(qe, mraw, cnt_lo, cnt_hi);
            }
            *reinterpret_cast<uint2*>(fa.tiecnt + (size_t)r * S::D + c) = make_uint2(cnt_lo, cnt_hi);
          }
          float o[8];
          if (e0[ps] == e1[ps]) {
            const float z = act_fwd_t<bf16>(0.f, fa.act);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = z;
          } else {
            float p[8], mx[8];
            R::to_float(praw, p);
            R::to_float(mraw, mx);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = p[i] + mx[i];
            act_fwd_vec<bf16, 8>(o, fa.act);
          }
          store_vec(fa.Bout + (size_t)r * S::D + c, o);
        }
      } else {
        // operands of all passes first (one exposed global latency per tile), then the shared-memory gathers
        uint4 draw[S::PASSES], yraw[S::PASSES];
#pragma unroll
        for (int ps = 0; ps < S::PASSES; ++ps) {
          const int r = r0 + ps * 32 + grp;
          if (r < r1) {
            draw[ps] = R::load(fa.dZA + (size_t)r * S::D + c);
            yraw[ps] = R::load(fa.A_prev + (size_t)r * S::D + c);
          }
        }
#pragma unroll
        for (int ps = 0; ps < S::PASSES; ++ps) {
          const int rr = ps * 32 + grp;
          const int r = r0 + rr;
          if (r >= r1) continue;
          const int tidx[4] = {n4[ps].x, n4[ps].y, n4[ps].z, n4[ps].w};
          uint4 g[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (tidx[u] >= 0) g[u] = lds128(sTa + fused_chunk_off((tidx[u] - r0) & 127, lig));
          // second half of this row of dS goes to HBM for the bond backward
          *reinterpret_cast<uint4*>(fa.dS + (size_t)r * (2 * S::D) + S::D + c) =
              lds128(sTa + fused_chunk_off(rr, 16 + lig));
          float d[8], y[8];
          R::to_float(draw[ps], d);
          R::to_float(yraw[ps], y);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (tidx[u] >= 0) v2::raw_add<bf16, 8>(g[u], d);
          for (int e = e0[ps] + 4; e < e1[ps]; ++e)
            v2::raw_fma<bf16, 8>(1.f, lds128(sTa + fused_chunk_off((__ldg(fa.nbr + e) - r0) & 127, lig)), d);
          act_grad_mul_vec<bf16, 8>(d, y, fa.act);
          store_vec(fa.dZA_prev + (size_t)r * S::D + c, d);
        }
      }
      named_bar_sync(1, S::EPI_THREADS);  // everyone is done with the staged tile
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

// A: [rows,128] bf16 GEMM operand, W: [256,128] bf16.  Returns GCB_ERR_UNSUPPORTED when the tcgen05
// path cannot take the call (alignment); the caller then runs the unfused kernels.
template <int MODE, bool TIES>
inline int launch_fused_rows_tc(const char* name, const bf16* A, const bf16* W, const FusedArgs& fa, cudaStream_t s) {
  using S = FusedShape;
  if (fa.n_tiles <= 0 || fa.rows <= 0) return GCB_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) & 15) return GCB_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    GCB_CUDA(cudaFuncSetAttribute(fused_rows_tc_kernel<MODE, TIES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap tmA, tmW;
  if (make_tmap_bf16(&tmA, A, fa.rows, S::K, S::K, S::BM) || make_tmap_bf16(&tmW, W, S::N, S::K, S::K, S::N)) return GCB_ERR_CUDA;
  const int grid = fa.n_tiles < kNumSMs ? fa.n_tiles : kNumSMs;
  GCB_KERNEL(name, s, (fused_rows_tc_kernel<MODE, TIES><<<grid, S::THREADS, S::SMEM_BYTES, s>>>(tmA, tmW, fa)));
  return GCB_OK;
}

}  // namespace tc
}  // namespace gcb
