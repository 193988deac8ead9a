// Fused bond update on tcgen05 with CHANNEL-MAJOR accumulators (bf16, D = 128), third generation.
//
//   B_l[r] = act( P[r] + b + max_{e -> r} Q[src e] ),   [P|Q] = B_{l-1} [U;V]^T        (reference gcn.py:163-164)
//
// The earlier fused kernels (fused_rows.cuh) kept the accumulator row-major in TMEM (lane = bond row): the row that
// owns a lane can read its own P, but its neighbours' Q rows live in OTHER lanes, so Q had to be drained to shared
// memory and gathered thread-per-row with bank conflicts, in lock-step behind CTA-wide barriers (79 us per launch,
// slower than GEMM + tile kernel).  Here the product is issued TRANSPOSED:
//
//   D^T[c][r] = sum_k W[c][k] B[r][k]      A operand = [U;V] (M = 2 x 128 output channels), B operand = the row tile
//
// so that TMEM lane = output channel and TMEM column = bond row.  A row warp owns 32 channels x 32 rows: it reads its
// block of both accumulators with two wide tcgen05.ld (P stays in registers, indexed by row at compile time in the
// unrolled row loop), and writes its Q block to shared memory as [row][channel] fp32.  Because a warp walks the rows
// of a closed tile together, "Q of neighbour s" is the SAME row for all 32 lanes: a 128-byte contiguous, conflict-free
// shared-memory read (the thread-per-row layout gathered 32 different rows per instruction).  The max runs on the
// fp32 accumulators (no bf16 rounding of P and Q: closer to the fp32 reference than the unfused pipeline), the tie
// bits of the max are one __ballot_sync per in-edge (bit = lane = channel: exactly the 32-bit word layout of the tie
// mask), and results leave as 64-byte row segments straight from registers.  [P|Q] never exists in global memory.
// (A first version read every neighbour with its own one-column tcgen05.ld and no shared memory at all: 243 us per
// launch -- small TMEM loads cost ~16 cycles each whatever their size, profiles/r2_fused_ab.jsonl.)
//
// Warp roles: warp 0 = TMA producer (weights once, then a ring of row tiles), warp 1 = MMA issuer (two M128 x N128 x
// K128 products per tile into a double-buffered TMEM accumulator pair), warps 2..17 = row warps
// (channel quarter = warp % 4, 32 of the tile's 128 rows each).
#pragma once
#include "gemm_tc.cuh"

namespace gcb {
namespace cm {

using namespace gcb::tc;

constexpr int kRowWarps = 16;
constexpr int kThreads = 64 + kRowWarps * 32;
constexpr int kStages = 2;
constexpr int kWBytes = 256 * 128 * 2;        // [U;V]: 2 K-slabs of [256 rows x 128 B]
constexpr int kTileBytes = 128 * 128 * 2;     // one row tile: 2 K-slabs of [128 rows x 128 B]
constexpr int kIdxBytes = 32 * 16 + 32 * 8;   // per row warp: tile-local ELL-4 neighbours + CSR range of its 32 rows
constexpr int kQBytes = 128 * 128 * 4;         // the tile's Q as [row][channel] fp32
constexpr int kSmemBytes = 1024 + kWBytes + kStages * kTileBytes + kQBytes + kRowWarps * kIdxBytes + 256;

struct Args {
  const int32_t* tile_ptr = nullptr;  // [n_tiles + 1] closed row tiles
  int n_tiles = 0;
  int rows = 0;                       // live bond rows M
  int act = 0;
  const int32_t* ptr = nullptr;       // dst-sorted CSR of the bond step: [rows + 1]
  const int32_t* nbr = nullptr;       // [E] source rows
  const int4* nbr4 = nullptr;         // [rows] ELL-4 view of nbr
  const float* bias = nullptr;        // [128] b_bond
  bf16* out = nullptr;                // B_l [rows, 128]
  uint32_t* tiemask = nullptr;        // training: [E][4] words, bit c of word p = channel 32 p + c attains the max
  uint8_t* tiecnt = nullptr;          // training: [rows][128] number of in-edges that tie for the max
};

__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }
__device__ __forceinline__ uint2 lds64(uint32_t saddr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t saddr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(saddr), "r"(a), "r"(b) : "memory");
}

template <bool TIES>
__global__ void __launch_bounds__(kThreads, 1)
bond_fwd_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmW, Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sT = sW + kWBytes;
  uint8_t* sQ = sT + kStages * kTileBytes;
  uint8_t* sIdx = sQ + kQBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sIdx + kRowWarps * kIdxBytes);
  uint64_t* full = bars;                    // [kStages] TMA -> MMA
  uint64_t* empty = bars + kStages;         // [kStages] MMA -> TMA
  uint64_t* tfull = bars + 2 * kStages;     // [2] MMA -> row warps
  uint64_t* tempty = tfull + 2;             // [2] row warps -> MMA
  uint64_t* wfull = tempty + 2;             // weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = a.n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmB); prefetch_tmap(&tmW);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kRowWarps); }
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull, kWBytes);
      for (int ks = 0; ks < 2; ++ks) tma_load_2d(sW + ks * (256 * 128), &tmW, wfull, ks * 64, 0);
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int r0 = __ldg(a.tile_ptr + tile);
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], kTileBytes);
        // 128 rows from r0: rows past the tile end belong to the next tile (their columns are never read), rows
        // past the matrix end are zero-filled by TMA
        for (int ks = 0; ks < 2; ++ks) tma_load_2d(sT + stage * kTileBytes + ks * (128 * 128), &tmB, &full[stage], ks * 64, r0);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_bf16(128, 128, 0, 0);
      mbar_wait(wfull, 0);
      const uint32_t w_base = smem_u32(sW);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t t_base = smem_u32(sT + stage * kTileBytes);
#pragma unroll
        for (int m = 0; m < 2; ++m) {      // m = 0: P channels (weight rows 0..127), m = 1: Q channels (128..255)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t ks = kk >> 2, ko = (kk & 3) * 32;
            const uint64_t dw = smem_desc_sw128(w_base + ks * (256 * 128) + m * (128 * 128) + ko, 16, 1024);
            const uint64_t dt = smem_desc_sw128(t_base + ks * (128 * 128) + ko, 16, 1024);
            umma_bf16(tmem_base + acc * 256 + m * 128, dw, dt, idesc, kk > 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty[stage]);
        umma_commit(&tfull[acc]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== row warps: lane = channel 32 q + lane, rows [32 rg, 32 rg + 32) of every tile =====================
    const int q = warp & 3;
    const int rg = (warp - 2) >> 2;
    const int ch = q * 32 + lane;
    const uint32_t tq = (uint32_t)(q * 32) << 16;
    const uint32_t widx = smem_u32(sIdx) + (warp - 2) * kIdxBytes;
    const uint32_t sq = smem_u32(sQ);
    const float bias_c = a.bias ? __ldg(a.bias + ch) : 0.f;
    const float z0 = act_fwd_fast(0.f, a.act);  // rows without in-edges
    pdl_wait();
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int r0 = __ldg(a.tile_ptr + tile);
      const int nrows = min(__ldg(a.tile_ptr + tile + 1), a.rows) - r0;
      {  // this warp's 32 rows: tile-local ELL-4 neighbours and CSR ranges into the warp's scratch (lane = row)
        const int my = rg * 32 + lane;
        int4 n4 = make_int4(-1, -1, -1, -1);
        int e0 = 0, e1 = 0;
        if (my < nrows) {
          n4 = __ldg(a.nbr4 + r0 + my);
          e0 = __ldg(a.ptr + r0 + my);
          e1 = __ldg(a.ptr + r0 + my + 1);
          n4.x = n4.x >= 0 ? n4.x - r0 : -1; n4.y = n4.y >= 0 ? n4.y - r0 : -1;
          n4.z = n4.z >= 0 ? n4.z - r0 : -1; n4.w = n4.w >= 0 ? n4.w - r0 : -1;
        }
        __syncwarp();
        sts128(widx + lane * 16, make_uint4((uint32_t)n4.x, (uint32_t)n4.y, (uint32_t)n4.z, (uint32_t)n4.w));
        sts64(widx + 512 + lane * 8, (uint32_t)e0, (uint32_t)e1);
        __syncwarp();
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t tP = tmem_base + tq + acc * 256, tQ = tP + 128;
      // this warp's block of both accumulators: 32 channels (lanes) x its 32 rows (columns), two wide loads
      uint32_t p[32], v[32];
      tmem_ld_32x32(tP + rg * 32, p);
      tmem_ld_32x32(tQ + rg * 32, v);
      tmem_ld_wait();
      // accumulators read: hand the TMEM pair back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      // Q goes to shared memory as [row][channel] fp32: a warp writes, and later gathers, 128 contiguous bytes of
      // ONE row at a time (the row is warp-uniform), so neither direction has bank conflicts
      named_bar_sync(1, kRowWarps * 32);  // every warp is done gathering from the previous tile's Q
#pragma unroll
      for (int j = 0; j < 32; ++j) sts32(sq + (uint32_t)(rg * 32 + j) * 512u + (uint32_t)ch * 4u, v[j]);
      named_bar_sync(1, kRowWarps * 32);  // the tile's Q is complete
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int row = rg * 32 + j;
        if (row >= nrows) break;
        const uint4 n4 = lds128(widx + j * 16);
        const uint2 ee = lds64(widx + 512 + j * 8);
        const int nn[4] = {(int)n4.x, (int)n4.y, (int)n4.z, (int)n4.w};
        const int e0 = (int)ee.x, e1 = (int)ee.y;
        float qf[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          qf[u] = nn[u] >= 0 ? __uint_as_float(lds32(sq + (uint32_t)(nn[u] & 127) * 512u + (uint32_t)ch * 4u)) : -INFINITY;
        float m = fmaxf(fmaxf(qf[0], qf[1]), fmaxf(qf[2], qf[3]));
        if (e1 - e0 > 4) {  // rare: more than four in-edges, the rest of the CSR row
          for (int e = e0 + 4; e < e1; ++e)
            m = fmaxf(m, __uint_as_float(lds32(sq + (uint32_t)((__ldg(a.nbr + e) - r0) & 127) * 512u + (uint32_t)ch * 4u)));
        }
        if (TIES && e1 > e0) {
          uint32_t cnt = 0u;
          const int n_ell = min(e1 - e0, 4);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (u < n_ell) {
              const bool tie = qf[u] == m;
              const uint32_t w = __ballot_sync(0xFFFFFFFFu, tie);
              cnt += tie ? 1u : 0u;
              if (lane == 0) a.tiemask[(size_t)(e0 + u) * 4 + q] = w;
            }
          }
          for (int e = e0 + 4; e < e1; ++e) {
            const float t = __uint_as_float(lds32(sq + (uint32_t)((__ldg(a.nbr + e) - r0) & 127) * 512u + (uint32_t)ch * 4u));
            const bool tie = t == m;
            const uint32_t w = __ballot_sync(0xFFFFFFFFu, tie);
            cnt += tie ? 1u : 0u;
            if (lane == 0) a.tiemask[(size_t)e * 4 + q] = w;
          }
          a.tiecnt[(size_t)(r0 + row) * 128 + ch] = (uint8_t)min(cnt, 255u);
        }
        const float o = e1 > e0 ? act_fwd_fast(__uint_as_float(p[j]) + bias_c + m, a.act) : z0;
        a.out[(size_t)(r0 + row) * 128 + ch] = __float2bfloat16_rn(o);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// B_{l-1} [rows, 128] bf16, W = [U;V] [256, 128] bf16; returns GCB_ERR_UNSUPPORTED when the kernel does not apply.
template <bool TIES>
inline int launch_bond_fwd(const bf16* Bprev, const bf16* W, const Args& a, cudaStream_t s) {
  if (a.n_tiles <= 0 || a.rows <= 0) return GCB_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(Bprev) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(a.out)) & 15) return GCB_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    GCB_CUDA(cudaFuncSetAttribute(bond_fwd_kernel<TIES>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  CUtensorMap tmB, tmW;
  if (make_tmap_bf16(&tmB, Bprev, a.rows, 128, 128, 128) || make_tmap_bf16(&tmW, W, 256, 128, 128, 256)) return GCB_ERR_CUDA;
  const int grid = a.n_tiles < kNumSMs ? a.n_tiles : kNumSMs;
  GCB_KERNEL("cm_bond_fwd_kernel", s, launch_pdl(bond_fwd_kernel<TIES>, grid, kThreads, kSmemBytes, s, tmB, tmW, a));
  return GCB_OK;
}

}  // namespace cm
}  // namespace gcb
