// Tile-staged row kernels (bf16, D = 128 or 256): the message-passing gathers out of shared memory.
//
// The unfused row kernels (kernels_v2.cuh) chase  index -> neighbour row  through global memory and
// are bound by that dependent latency at the occupancy their registers allow (25-35 % of HBM
// bandwidth in the round-1 profiles).  Every neighbour of a row lives in the same molecule, and the
// packed batch carries CLOSED row tiles (<= 128 rows, no edge leaves a tile; GCB_H_TILE_PTR).  Here
// one CTA owns one tile: a single thread asks TMA for the tile's operand rows (one or more
// 128-row boxes, no swizzle: a staged row is 256 contiguous bytes), the threads stage the tile's CSR
// indices (ELL-4 views + row pointers) in shared memory meanwhile, and after ONE barrier wait all
// tile-local gathers are shared-memory reads.  16 lanes own a row (16-byte chunks), a warp owns two
// adjacent rows, so global stores stay coalesced.  CTAs are small (256 threads, <= 64 registers) so
// that 3-4 of them per SM overlap each other's load and compute phases.
//
// Arithmetic, rounding points and summation order are those of the v2 kernels: results are
// bit-identical (tests/test_fused_gpu.py).
#pragma once
#include "kernels_v2.cuh"
#include <cstdlib>

#include "tc_common.cuh"

namespace gcb {
namespace tc {
// Row-major bf16 matrix [rows, cols] (leading dimension ld), box = [box_rows x box_cols], no swizzle.
int make_tmap_bf16_plain(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                         int box_cols);
}  // namespace tc

namespace tile {

constexpr int kRows = GCB_TILE_ROWS;              // 128
constexpr int kThreads = 256;
constexpr int kIdxBytes = 2 * kRows * 16 + 544 + 1024;   // two ELL-4 arrays + 129 row pointers (padded) + pad row
constexpr int kTieEdges = 640;                           // in-edges of a tile whose tie bits bond_bwd stages (5 per row)
constexpr int kPadBytes = 1024;                          // >= the widest staged row ([P|Q] at D = 256 is staged Q-only: 512 B)

// CH = 16-byte chunks (= lanes) per row: 16 for D = 128, 32 for D = 256.
template <int CH>
struct Cfg {
  static constexpr int D = CH * 8;
  static constexpr int kSlots = kThreads / CH;        // rows per pass: 16 / 8
  static constexpr int kPasses = kRows / kSlots;      // 8 / 16
  static constexpr int kRowBytes = D * 2;             // 256 / 512
  static constexpr int kBoxBytes = kRows * kRowBytes; // one staged [128 x D] bf16 box: 32 / 64 KB
  static constexpr int kCtas = CH == 16 ? 4 : 3;      // CTAs per SM the launch bounds aim for
};

using R = v2::Raw<bf16>;

struct TileCommon {
  const int32_t* tile_ptr;  // [n_tiles + 1]
  int rows;                 // live rows
  const int32_t* ptr;       // CSR row pointer of the row phase
  const int32_t* nbr;       // CSR neighbour rows
  const int4* nbr4;         // ELL-4 view of nbr
  const int4* aux4;         // second ELL-4 view walked with the same CSR (edge ids / dst-CSR positions) or null
  const int32_t* aux;       // CSR form of aux4 (rows with more than four entries) or null
  int rev = 0;              // walk the tiles last-to-first: start on the rows the producer kernel wrote last (L2 resident)
  int pf = 0;               // set by the launchers (tile_prefetch()): bulk L2 prefetch of the tile's row-streamed operands
};

// A/B switch GCB_TILE_PREFETCH (bit mask; 0 = off): 1 bond backward's row operands (B_l, dBnext, tie counts), 2 its gathered
// dS rows, 4 / 8 atom backward's A_prev / dZA rows.  Measured on B200 (profiles/r2_ab_v_prefetch.jsonl): 3 takes the bond
// backward from 91 to 86 us (train step 1.506 -> 1.484 ms); 4 and 8 make the atom backward 1-2 us SLOWER, a prefetch one
// wave of CTAs ahead costs 4-7 us, and prefetching the by-edge-id bond rows of the atom aggregation gains 0.15 %.
inline int tile_prefetch() {
  static const int v = [] { const char* e = std::getenv("GCB_TILE_PREFETCH"); return e ? std::atoi(e) : 3; }();
  return v;
}

// One instruction asks L2 for a contiguous run of rows (the tile's [r0, r1) x row bytes, a multiple of 16): the row
// phase then finds in L2 what it would otherwise fetch from DRAM two 16-byte chunks per thread at a time behind a
// long-scoreboard stall.  L2 is the coherence point, so a prefetch is safe at any time (also before griddepcontrol.wait).
__device__ __forceinline__ void prefetch_l2_rows(const void* p, int bytes) {
  if (bytes > 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l2_line(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p) : "memory");
}

// Tile of this CTA.  CTAs are dispatched in blockIdx order, so `rev` makes the kernel's first wave read the END of
// its operand -- which is what the previous kernel of the chain (walking the other way) produced last.
__device__ __forceinline__ int tile_of_cta(const TileCommon& cm) {
  return cm.rev ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x;
}

__device__ __forceinline__ uint4 lds16(uint32_t saddr) { return gcb::tc::lds128(saddr); }
__device__ __forceinline__ int lds32(uint32_t saddr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t saddr, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }

// Tile prologue: [r0, r1) of this CTA's tile; one thread issues the TMA boxes (ISSUE, BOX_BYTES in
// total, 0 = none) while all threads stage the indices; ends with everything visible.
//
// Branch-free neighbours: the ELL-4 view is staged as BYTE OFFSETS of the neighbours' staged rows
// ((nbr - r0) * STRIDE), and a missing slot (-1) points at a PAD ROW behind the index arrays that holds the
// identity of the reduction (PADWORD: -inf for max, 0 for sums).  The row loops then load and reduce all four
// slots unconditionally -- the `if (nidx[u] >= 0)` tests cost a quarter of the instructions of these
// issue-bound kernels (BSSY / BSYNC / BRA / ISETP per slot).
#define GCB_TILE_PROLOGUE(BOX_BYTES, ISSUE, STRIDE, PADWORD) GCB_TILE_PROLOGUE_PF(BOX_BYTES, ISSUE, STRIDE, PADWORD, (void)0)
#define GCB_TILE_PROLOGUE_PF(BOX_BYTES, ISSUE, STRIDE, PADWORD, PREFETCH)                           \
  constexpr int kD = Cfg<CH>::D, kSlots = Cfg<CH>::kSlots, kPasses = Cfg<CH>::kPasses;              \
  constexpr int kRowBytes = Cfg<CH>::kRowBytes, kBoxBytes = Cfg<CH>::kBoxBytes;                     \
  (void)kD; (void)kSlots; (void)kPasses; (void)kRowBytes; (void)kBoxBytes;                          \
  extern __shared__ uint8_t smem_raw[];                                                             \
  const uint32_t smem = (gcb::tc::smem_u32(smem_raw) + 127u) & ~127u; /* 32-bit shared address */   \
  const uint32_t s_n4 = smem + (BOX_BYTES), s_a4 = s_n4 + kRows * 16, s_ptr = s_a4 + kRows * 16;    \
  const uint32_t s_pad = s_ptr + 544; /* one staged row of PADWORD (kPadBytes) */                   \
  constexpr uint32_t kPadOff = (uint32_t)(BOX_BYTES) + kRows * 32 + 544;                            \
  __shared__ uint64_t bar;                                                                          \
  const int tb_ = tile_of_cta(cm);                                                                  \
  const int r0 = __ldg(cm.tile_ptr + tb_);                                                          \
  const int r1 = min(__ldg(cm.tile_ptr + tb_ + 1), cm.rows);                                        \
  pdl_launch_dependents();                                                                          \
  if (threadIdx.x == 32 && cm.pf) { PREFETCH; } /* another warp than the TMA issuer */              \
  if ((BOX_BYTES) > 0 && threadIdx.x == 0) {                                                        \
    gcb::tc::mbar_init(&bar, 1);                                                                    \
    gcb::tc::fence_barrier_init();                                                                  \
    pdl_wait(); /* the boxes hold activations of the previous kernel */                             \
    gcb::tc::mbar_arrive_expect_tx(&bar, (BOX_BYTES));                                              \
    ISSUE;                                                                                          \
  }                                                                                                 \
  for (int i = threadIdx.x; i < kPadBytes / 16; i += kThreads)                                      \
    gcb::tc::sts128(s_pad + i * 16, make_uint4((PADWORD), (PADWORD), (PADWORD), (PADWORD)));        \
  for (int i = threadIdx.x; i <= kRows; i += kThreads) {                                            \
    const int r = r0 + i;                                                                           \
    if (i < kRows) {                                                                                \
      int4 v = make_int4(-1, -1, -1, -1), w = v;                                                    \
      if (r < r1) {                                                                                 \
        v = __ldg(cm.nbr4 + r);                                                                     \
        if (cm.aux4) w = __ldg(cm.aux4 + r);                                                        \
      }                                                                                             \
      const uint32_t o0 = v.x >= 0 ? (uint32_t)(v.x - r0) * (STRIDE) : kPadOff;                      \
      const uint32_t o1 = v.y >= 0 ? (uint32_t)(v.y - r0) * (STRIDE) : kPadOff;                      \
      const uint32_t o2 = v.z >= 0 ? (uint32_t)(v.z - r0) * (STRIDE) : kPadOff;                      \
      const uint32_t o3 = v.w >= 0 ? (uint32_t)(v.w - r0) * (STRIDE) : kPadOff;                      \
      gcb::tc::sts128(s_n4 + i * 16, make_uint4(o0, o1, o2, o3));                                   \
      gcb::tc::sts128(s_a4 + i * 16, make_uint4(w.x, w.y, w.z, w.w));                               \
    }                                                                                               \
    sts32(s_ptr + i * 4, r <= r1 ? __ldg(cm.ptr + r) : 0);                                          \
  }                                                                                                 \
  const int grp = threadIdx.x / CH, lig = threadIdx.x % CH, c = lig * 8;                            \
  (void)c; (void)s_a4;                                                                              \
  pdl_wait(); /* indices are immutable batch data and were staged above; activations come after this */ \
  __syncthreads(); /* indices staged, barrier initialised */                                        \
  if ((BOX_BYTES) > 0) gcb::tc::mbar_wait(&bar, 0);

// noff[u]: byte offset (from `smem`) of the staged row of ELL slot u, or of the pad row
#define GCB_TILE_ROW(PS)                                                                            \
  const int rr = (PS) * kSlots + grp;                                                               \
  const int r = r0 + rr;                                                                            \
  if (r >= r1) continue;                                                                            \
  const uint4 n4_ = lds16(s_n4 + rr * 16);                                                          \
  const uint32_t noff[4] = {n4_.x, n4_.y, n4_.z, n4_.w};                                            \
  const int e0 = lds32(s_ptr + rr * 4), e1 = lds32(s_ptr + rr * 4 + 4);

// Tie bookkeeping of the max aggregation on packed words.  q, m: 8 bf16 channels (channel 2i = low
// half of word i).  Returns bit i = (q_i == m_i) and adds 1 to byte counter i of cnt_lo (channels
// 0-3) / cnt_hi (channels 4-7); exact up to 255 in-edges per row, like the bf16 counters of the v2
// kernels.
__device__ __forceinline__ uint32_t tie_bits(const uint4& q, const uint4& m, uint32_t& cnt_lo, uint32_t& cnt_hi) {
  auto eqm = [](uint32_t a, uint32_t b) -> uint32_t {  // 0xFFFF per equal half
    return __heq2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  };
  // bytes 0/2 of each mask word = lo/hi half flags -> one flag byte per channel
  const uint32_t f_lo = __byte_perm(eqm(q.x, m.x), eqm(q.y, m.y), 0x6420);  // channels 0,1,2,3
  const uint32_t f_hi = __byte_perm(eqm(q.z, m.z), eqm(q.w, m.w), 0x6420);  // channels 4,5,6,7
  cnt_lo += f_lo & 0x01010101u;
  cnt_hi += f_hi & 0x01010101u;
  const uint32_t t = (f_lo & 0x08040201u) | (f_hi & 0x80402010u);
  return (t * 0x01010101u) >> 24;
}

// ---- bond update: Bout[r] = act(P[r] + max_{e->r} Q[src e]);  staged: [P|Q] rows of the tile (512 B each) ----
// QONLY: stage only the Q half (32 KB, more CTAs per SM) and stream the row's own P chunk from global.
template <int CH, bool TIES, bool QONLY>
__global__ void __launch_bounds__(kThreads, (QONLY && CH == 16) ? 4 : 3)
bond_update_kernel(const __grid_constant__ CUtensorMap tmPQ, TileCommon cm, const bf16* __restrict__ PQ,
                   bf16* __restrict__ Bout, uint8_t* __restrict__ tiemask, uint8_t* __restrict__ tiecnt, int act) {
  constexpr int kQRow = QONLY ? Cfg<CH>::kRowBytes : 2 * Cfg<CH>::kRowBytes;  // bytes per staged row
  constexpr int kQOff = QONLY ? 0 : Cfg<CH>::kRowBytes;                        // offset of the Q half inside a staged row
  GCB_TILE_PROLOGUE((QONLY ? 1 : 2) * Cfg<CH>::kBoxBytes, gcb::tc::tma_load_2d_s(smem, &tmPQ, &bar, QONLY ? Cfg<CH>::D : 0, r0),
                    kQRow, 0xFF80FF80u)
#pragma unroll 2
  for (int ps = 0; ps < kPasses; ++ps) {
    GCB_TILE_ROW(ps)
    const uint4 praw = QONLY ? R::load(PQ + (size_t)r * (2 * kD) + c) : lds16(smem + rr * kQRow + lig * 16);
    uint4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = lds16(smem + noff[u] + kQOff + lig * 16);  // pad row: -inf
    uint4 mraw = R::vmax(R::vmax(q[0], q[1]), R::vmax(q[2], q[3]));
    for (int e = e0 + 4; e < e1; ++e)
      mraw = R::vmax(mraw, lds16(smem + ((__ldg(cm.nbr + e) - r0) & 127) * kQRow + kQOff + lig * 16));
    if (TIES && e1 > e0) {
      // per in-edge: which of the 8 channels attain the max (one byte); per channel: how many edges tie
      // (a pad slot is -inf and never equals the max of a row that has an in-edge)
      uint32_t cnt_lo = 0u, cnt_hi = 0u;  // four byte counters each: channels 0-3 / 4-7
      const int deg = e1 - e0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t b = tie_bits(q[u], mraw, cnt_lo, cnt_hi);
        if (u < deg) tiemask[(size_t)(e0 + u) * CH + lig] = (uint8_t)b;
      }
      for (int e = e0 + 4; e < e1; ++e) {
        const uint4 qe = lds16(smem + ((__ldg(cm.nbr + e) - r0) & 127) * kQRow + kQOff + lig * 16);
        tiemask[(size_t)e * CH + lig] = (uint8_t)tie_bits(qe, mraw, cnt_lo, cnt_hi);
      }
      *reinterpret_cast<uint2*>(tiecnt + (size_t)r * kD + c) = make_uint2(cnt_lo, cnt_hi);
    }
    float o[8];
    if (e0 == e1) {
      const float z = act_fwd_t<bf16>(0.f, act);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = z;
    } else {
      float p[8], mx[8];
      R::to_float(praw, p);
      R::to_float(mraw, mx);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = p[i] + mx[i];
      act_fwd_vec<bf16, 8>(o, act);
    }
    store_vec(Bout + (size_t)r * kD + c, o);
  }
}

// ---- first bond update straight from the token table (embedding lookup fused with the first message step) ----
// In message step 1 the bond matrix is a row copy out of the pre-activated table, B_0[r] = tab_bond[tok r] (reference
// gcn.py:156-157), so [P|Q]_1[r] = tabPQ[tok r] with tabPQ = tab_bond [U;V]^T + [b|0]: a [bond_vocab x 2D] table that
// the step's tcgen05 GEMM produces once per weight refresh (gcb_prepare_weights).  The kernel stages the table
// (<= 64 KB) and the tile's tokens instead of a [128 x 2D] tile of a [rows x 2D] matrix that then never exists:
// B_0, the first [P|Q] GEMM and their 5 N D s bytes of HBM traffic disappear.  Same arithmetic, same rounding
// points and the same tie bookkeeping as bond_update_kernel.
template <int CH, bool TIES, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
bond_update_tok_kernel(TileCommon cm, const bf16* __restrict__ tabPQ, const int32_t* __restrict__ tok, int vocab,
                       bf16* __restrict__ Bout, uint8_t* __restrict__ tiemask, uint8_t* __restrict__ tiecnt, int act) {
  constexpr int kD = Cfg<CH>::D, kSlots = Cfg<CH>::kSlots, kPasses = Cfg<CH>::kPasses;
  constexpr int kRowBytes = 2 * Cfg<CH>::kRowBytes;  // one table row: [P | Q]
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (gcb::tc::smem_u32(smem_raw) + 127u) & ~127u;  // table [vocab][2 D] bf16
  // table rows [0, vocab) + one pad row of -inf (what a missing ELL slot reads), then per row: the byte offsets of its
  // four neighbours' TABLE rows (token looked up once, here), its own table-row offset, its CSR range
  const uint32_t s_n4 = smem + (vocab + 1) * kRowBytes, s_tok = s_n4 + kRows * 16, s_ptr = s_tok + kRows * 4;
  const uint32_t pad_off = (uint32_t)vocab * kRowBytes;
  const int tb_ = tile_of_cta(cm);
  const int r0 = __ldg(cm.tile_ptr + tb_);
  const int r1 = min(__ldg(cm.tile_ptr + tb_ + 1), cm.rows);
  pdl_launch_dependents();
  // indices and tokens are immutable batch data; the table was written by the weight refresh at the end of the
  // previous step (complete and visible before this kernel's pdl_wait returns)
  for (int i = threadIdx.x; i <= kRows; i += kThreads) {
    const int r = r0 + i;
    if (i < kRows) {
      int4 v = make_int4(-1, -1, -1, -1);
      int t = 0;
      if (r < r1) { v = __ldg(cm.nbr4 + r); t = __ldg(tok + r); }
      const uint32_t o0 = v.x >= 0 ? (uint32_t)__ldg(tok + v.x) * kRowBytes : pad_off, o1 = v.y >= 0 ? (uint32_t)__ldg(tok + v.y) * kRowBytes : pad_off;
      const uint32_t o2 = v.z >= 0 ? (uint32_t)__ldg(tok + v.z) * kRowBytes : pad_off, o3 = v.w >= 0 ? (uint32_t)__ldg(tok + v.w) * kRowBytes : pad_off;
      gcb::tc::sts128(s_n4 + i * 16, make_uint4(o0, o1, o2, o3));
      sts32(s_tok + i * 4, t * kRowBytes);
    }
    sts32(s_ptr + i * 4, r <= r1 ? __ldg(cm.ptr + r) : 0);
  }
  pdl_wait();
  {
    const uint4* src = reinterpret_cast<const uint4*>(tabPQ);
    const int n16 = vocab * (kRowBytes / 16);
    for (int i = threadIdx.x; i < n16; i += kThreads) gcb::tc::sts128(smem + i * 16, __ldg(src + i));
    for (int i = threadIdx.x; i < kRowBytes / 16; i += kThreads)
      gcb::tc::sts128(smem + pad_off + i * 16, make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u));
  }
  const int grp = threadIdx.x / CH, lig = threadIdx.x % CH, c = lig * 8;
  __syncthreads();
  constexpr int kQOff = Cfg<CH>::kRowBytes;
  auto qtail = [&](int nbr_row) -> uint32_t {  // CSR tail (more than four in-edges): token of a tile-local neighbour row
    return smem + (uint32_t)__ldg(tok + nbr_row) * kRowBytes + kQOff + lig * 16;
  };
#pragma unroll 2
  for (int ps = 0; ps < kPasses; ++ps) {
    GCB_TILE_ROW(ps)
    const uint4 praw = lds16(smem + lds32(s_tok + rr * 4) + lig * 16);
    uint4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = lds16(smem + noff[u] + kQOff + lig * 16);
    uint4 mraw = R::vmax(R::vmax(q[0], q[1]), R::vmax(q[2], q[3]));
    for (int e = e0 + 4; e < e1; ++e) mraw = R::vmax(mraw, lds16(qtail(__ldg(cm.nbr + e))));
    if (TIES && e1 > e0) {
      uint32_t cnt_lo = 0u, cnt_hi = 0u;
      const int deg = e1 - e0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t bm = tie_bits(q[u], mraw, cnt_lo, cnt_hi);
        if (u < deg) tiemask[(size_t)(e0 + u) * CH + lig] = (uint8_t)bm;
      }
      for (int e = e0 + 4; e < e1; ++e) {
        const uint4 qe = lds16(qtail(__ldg(cm.nbr + e)));
        tiemask[(size_t)e * CH + lig] = (uint8_t)tie_bits(qe, mraw, cnt_lo, cnt_hi);
      }
      *reinterpret_cast<uint2*>(tiecnt + (size_t)r * kD + c) = make_uint2(cnt_lo, cnt_hi);
    }
    float o[8];
    if (e0 == e1) {
      const float z = act_fwd_t<bf16>(0.f, act);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = z;
    } else {
      float p[8], mx[8];
      R::to_float(praw, p);
      R::to_float(mraw, mx);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = p[i] + mx[i];
      act_fwd_vec<bf16, 8>(o, act);
    }
    store_vec(Bout + (size_t)r * kD + c, o);
  }
}

// ---- atom aggregation: S[i] = [ sum_{e->i} A[src e] | sum_{e->i} Bnew[eid e] ]  (Bnew[e >= M] = act(0)) ----
// staged: A rows of the tile (neighbours are tile-local); the bond rows are addressed by EDGE id and
// live in other tiles, so they are gathered from global memory, issued before the shared-memory half.
template <int CH, int UN>
__global__ void __launch_bounds__(kThreads, Cfg<CH>::kCtas)
atom_gather_kernel(const __grid_constant__ CUtensorMap tmA, TileCommon cm, const bf16* __restrict__ Bnew, int M,
                   bf16* __restrict__ S, int act) {
  GCB_TILE_PROLOGUE(Cfg<CH>::kBoxBytes, gcb::tc::tma_load_2d_s(smem, &tmA, &bar, 0, r0), kRowBytes, 0u)
  const float cst = round_to(act_fwd_t<bf16>(0.f, act), (const bf16*)nullptr);
  // act(0) as a packed bf16 pair: what a dead bond row (edge id >= M) contributes to every channel
  const uint32_t cst2 = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(cst)) * 0x00010001u;
#pragma unroll (UN)
  for (int ps = 0; ps < kPasses; ++ps) {
    GCB_TILE_ROW(ps)
    const uint4 e4_ = lds16(s_a4 + rr * 16);
    const int eidx[4] = {(int)e4_.x, (int)e4_.y, (int)e4_.z, (int)e4_.w};
    // bond rows by EDGE id live in other tiles: global gathers, issued first.  Branch-free: a missing slot (-1) or
    // a dead row (>= M) reads row 0 and is replaced by 0 / act(0) afterwards (same values, same order of addition)
    uint4 b[4], a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool live = (unsigned)eidx[u] < (unsigned)M;
      b[u] = R::load(Bnew + (size_t)(live ? eidx[u] : 0) * kD + c);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = lds16(smem + noff[u] + lig * 16);  // pad row: zeros
    float sa[8], sb[8];
    v2::vzero<bf16, 8>(sa);
    v2::vzero<bf16, 8>(sb);
#pragma unroll
    for (int u = 0; u < 4; ++u) v2::raw_add<bf16, 8>(a[u], sa);
    for (int e = e0 + 4; e < e1; ++e)
      v2::raw_add<bf16, 8>(lds16(smem + ((__ldg(cm.nbr + e) - r0) & 127) * kRowBytes + lig * 16), sa);
    store_vec(S + (size_t)r * (2 * kD) + c, sa);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool live = (unsigned)eidx[u] < (unsigned)M;
      const uint32_t fill = eidx[u] < 0 ? 0u : cst2;
      uint4 t = b[u];
      t.x = live ? t.x : fill; t.y = live ? t.y : fill; t.z = live ? t.z : fill; t.w = live ? t.w : fill;
      v2::raw_add<bf16, 8>(t, sb);
    }
    for (int e = e0 + 4; e < e1; ++e) {
      const int eid = __ldg(cm.aux + e);
      if (eid < M) v2::raw_add<bf16, 8>(R::load(Bnew + (size_t)eid * kD + c), sb);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) sb[i] += cst;
      }
    }
    store_vec(S + (size_t)r * (2 * kD) + kD + c, sb);
  }
}

// ---- first atom aggregation straight from the token table (embedding lookup fused with the first message step) ----
// In message step 1 the atom matrix is a row copy out of the pre-activated table, A_0[i] = tab_atom[tok i] (reference
// gcn.py:152-153), so S_1[i] = [ sum_{e->i} tab_atom[tok(src e)] | sum_{e->i} B_1[eid e] ].  The kernel stages the table
// (vocab x D bf16, 16 KB at 64 tokens) and looks the neighbours' tokens up while it stages the indices; A_0 is never
// written, and the residual of the step's transform GEMM and act' in the backward pass read the table through the token
// as well (TcEpilogue::res_idx, atom_bwd_kernel's prev_idx).  Same arithmetic and order of addition as atom_gather_kernel.
template <int CH, int UN>
__global__ void __launch_bounds__(kThreads, Cfg<CH>::kCtas)
atom_gather_tok_kernel(TileCommon cm, const bf16* __restrict__ tabA, const int32_t* __restrict__ tok, int vocab,
                       const bf16* __restrict__ Bnew, int M, bf16* __restrict__ S, int act) {
  constexpr int kD = Cfg<CH>::D, kSlots = Cfg<CH>::kSlots, kPasses = Cfg<CH>::kPasses, kRowBytes = Cfg<CH>::kRowBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (gcb::tc::smem_u32(smem_raw) + 127u) & ~127u;  // table [vocab][D] bf16, then one zero pad row
  const uint32_t s_n4 = smem + (vocab + 1) * kRowBytes, s_a4 = s_n4 + kRows * 16, s_ptr = s_a4 + kRows * 16;
  const uint32_t pad_off = (uint32_t)vocab * kRowBytes;
  const int tb_ = tile_of_cta(cm);
  const int r0 = __ldg(cm.tile_ptr + tb_);
  const int r1 = min(__ldg(cm.tile_ptr + tb_ + 1), cm.rows);
  pdl_launch_dependents();
  // indices and tokens are immutable batch data; the table was written by the weight refresh of the previous step
  for (int i = threadIdx.x; i <= kRows; i += kThreads) {
    const int r = r0 + i;
    if (i < kRows) {
      int4 v = make_int4(-1, -1, -1, -1), w = v;
      if (r < r1) { v = __ldg(cm.nbr4 + r); w = __ldg(cm.aux4 + r); }
      const uint32_t o0 = v.x >= 0 ? (uint32_t)__ldg(tok + v.x) * kRowBytes : pad_off, o1 = v.y >= 0 ? (uint32_t)__ldg(tok + v.y) * kRowBytes : pad_off;
      const uint32_t o2 = v.z >= 0 ? (uint32_t)__ldg(tok + v.z) * kRowBytes : pad_off, o3 = v.w >= 0 ? (uint32_t)__ldg(tok + v.w) * kRowBytes : pad_off;
      gcb::tc::sts128(s_n4 + i * 16, make_uint4(o0, o1, o2, o3));
      gcb::tc::sts128(s_a4 + i * 16, make_uint4(w.x, w.y, w.z, w.w));
    }
    sts32(s_ptr + i * 4, r <= r1 ? __ldg(cm.ptr + r) : 0);
  }
  pdl_wait();
  {
    const uint4* src = reinterpret_cast<const uint4*>(tabA);
    const int n16 = vocab * (kRowBytes / 16);
    for (int i = threadIdx.x; i < n16; i += kThreads) gcb::tc::sts128(smem + i * 16, __ldg(src + i));
    for (int i = threadIdx.x; i < kRowBytes / 16; i += kThreads) gcb::tc::sts128(smem + pad_off + i * 16, make_uint4(0u, 0u, 0u, 0u));
  }
  const int grp = threadIdx.x / CH, lig = threadIdx.x % CH, c = lig * 8;
  __syncthreads();
  const float cst = round_to(act_fwd_t<bf16>(0.f, act), (const bf16*)nullptr);
  const uint32_t cst2 = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(cst)) * 0x00010001u;
#pragma unroll (UN)
  for (int ps = 0; ps < kPasses; ++ps) {
    GCB_TILE_ROW(ps)
    const uint4 e4_ = lds16(s_a4 + rr * 16);
    const int eidx[4] = {(int)e4_.x, (int)e4_.y, (int)e4_.z, (int)e4_.w};
    uint4 b[4], a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool live = (unsigned)eidx[u] < (unsigned)M;
      b[u] = R::load(Bnew + (size_t)(live ? eidx[u] : 0) * kD + c);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = lds16(smem + noff[u] + lig * 16);  // pad row: zeros
    float sa[8], sb[8];
    v2::vzero<bf16, 8>(sa);
    v2::vzero<bf16, 8>(sb);
#pragma unroll
    for (int u = 0; u < 4; ++u) v2::raw_add<bf16, 8>(a[u], sa);
    for (int e = e0 + 4; e < e1; ++e)
      v2::raw_add<bf16, 8>(lds16(smem + (uint32_t)__ldg(tok + __ldg(cm.nbr + e)) * kRowBytes + lig * 16), sa);
    store_vec(S + (size_t)r * (2 * kD) + c, sa);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool live = (unsigned)eidx[u] < (unsigned)M;
      const uint32_t fill = eidx[u] < 0 ? 0u : cst2;
      uint4 t = b[u];
      t.x = live ? t.x : fill; t.y = live ? t.y : fill; t.z = live ? t.z : fill; t.w = live ? t.w : fill;
      v2::raw_add<bf16, 8>(t, sb);
    }
    for (int e = e0 + 4; e < e1; ++e) {
      const int eid = __ldg(cm.aux + e);
      if (eid < M) v2::raw_add<bf16, 8>(R::load(Bnew + (size_t)eid * kD + c), sb);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) sb[i] += cst;
      }
    }
    store_vec(S + (size_t)r * (2 * kD) + kD + c, sb);
  }
}

// ---- atom backward: dZA_prev[j] = (dZA[j] + sum_{e: src e = j} dS[dst e, :D]) * act'(A_prev[j]) ----
// staged: dS[:, :D] rows of the tile (neighbours); the row's own dZA / A_prev chunks stream from global.
template <int CH, int UN>
__global__ void __launch_bounds__(kThreads, Cfg<CH>::kCtas)
atom_bwd_kernel(const __grid_constant__ CUtensorMap tmdS, TileCommon cm, const bf16* __restrict__ dZA,
                const bf16* __restrict__ A_prev, const int32_t* __restrict__ prev_idx, bf16* __restrict__ dZA_prev, int act) {
  // prev_idx (message step 1 with the embedding lookup fused in): A_prev is the pre-activated token table and row r
  // reads A_prev[prev_idx[r]] -- A_0 is never materialised
  GCB_TILE_PROLOGUE_PF(Cfg<CH>::kBoxBytes, gcb::tc::tma_load_2d_s(smem, &tmdS, &bar, 0, r0), kRowBytes, 0u,
                       (prefetch_l2_rows(prev_idx ? nullptr : A_prev + (size_t)r0 * Cfg<CH>::D, prev_idx || !(cm.pf & 4) ? 0 : (r1 - r0) * Cfg<CH>::kRowBytes),
                        prefetch_l2_rows(dZA + (size_t)r0 * Cfg<CH>::D, (cm.pf & 8) ? (r1 - r0) * Cfg<CH>::kRowBytes : 0)))
#pragma unroll (UN)
  for (int ps = 0; ps < kPasses; ++ps) {
    GCB_TILE_ROW(ps)
    const uint4 draw = R::load(dZA + (size_t)r * kD + c);
    const uint4 yraw = R::load(A_prev + (size_t)(prev_idx ? __ldg(prev_idx + r) : r) * kD + c);
    uint4 g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) g[u] = lds16(smem + noff[u] + lig * 16);  // pad row: zeros
    float d[8], y[8];
    R::to_float(draw, d);
    R::to_float(yraw, y);
#pragma unroll
    for (int u = 0; u < 4; ++u) v2::raw_add<bf16, 8>(g[u], d);
    for (int e = e0 + 4; e < e1; ++e)
      v2::raw_add<bf16, 8>(lds16(smem + ((__ldg(cm.nbr + e) - r0) & 127) * kRowBytes + lig * 16), d);
    act_grad_mul_vec<bf16, 8>(d, y, act);
    store_vec(dZA_prev + (size_t)r * kD + c, d);
  }
}

// ---- bond backward, source side: dQ[j] = sum_{e: src e = j} tie(e) * gsplit[dst e];  staged: gsplit rows ----
// tie bits come from tiemask[dst-CSR position of e] (global; 16 contiguous bytes per edge).
template <int CH>
__global__ void __launch_bounds__(kThreads, Cfg<CH>::kCtas)
bond_scatter_kernel(const __grid_constant__ CUtensorMap tmG, TileCommon cm, const uint8_t* __restrict__ tiemask,
                    bf16* __restrict__ dPQ) {
  GCB_TILE_PROLOGUE(Cfg<CH>::kBoxBytes, gcb::tc::tma_load_2d_s(smem, &tmG, &bar, 0, r0), kRowBytes, 0u)
#pragma unroll 2
  for (int ps = 0; ps < kPasses; ++ps) {
    GCB_TILE_ROW(ps)
    const uint4 o4_ = lds16(s_a4 + rr * 16);
    const int oidx[4] = {(int)o4_.x, (int)o4_.y, (int)o4_.z, (int)o4_.w};
    uint4 g[4];
    uint32_t bits[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      bits[u] = __ldg(tiemask + (size_t)max(oidx[u], 0) * CH + lig);  // a missing slot reads the zero pad row: its bits do not matter
      g[u] = lds16(smem + noff[u] + lig * 16);
    }
    float acc[8];
    v2::vzero<bf16, 8>(acc);
#pragma unroll
    for (int u = 0; u < 4; ++u) v2::raw_add_masked(g[u], bits[u], acc);
    for (int e = e0 + 4; e < e1; ++e)
      v2::raw_add_masked(lds16(smem + ((__ldg(cm.nbr + e) - r0) & 127) * kRowBytes + lig * 16),
                         __ldg(tiemask + (size_t)__ldg(cm.aux + e) * CH + lig), acc);
    store_vec(dPQ + (size_t)r * (2 * kD) + kD + c, acc);
  }
}

// ---- bond backward, both sides in one pass over the tile -------------------------------------------
//   phase 1 (destination side, = v2::bond_bwd_tie_kernel):
//       dzb[r] = (dS[dst r, D:2D] + dBnext[r] + d_out_bond[r]) * act'(B_l[r])      (0 when r has no in-edge)
//       dPQ[r, :D] = dzb[r]  (= dP);   g[r] = dzb[r] / tie count   -> shared memory (bf16, as gsplit would be)
//   phase 2 (source side, = bond_scatter_kernel):
//       dPQ[j, D:] = sum_{e: src e = j} tie(e) * g[dst e]          (dst e is tile-local)
// The split gradient never goes to HBM.  `dst` is the edge list's destination array: bond row r is
// EDGE r, and it collects the gradient of the atom it points to (gcn.py:163 quirk).
template <int CH, int UN>
__global__ void __launch_bounds__(kThreads, Cfg<CH>::kCtas)
bond_bwd_kernel(TileCommon cm, const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ dst,
                const bf16* __restrict__ dS, const bf16* __restrict__ dBnext, const bf16* __restrict__ d_out_bond,
                const bf16* __restrict__ B_l, const uint8_t* __restrict__ tiecnt, const uint8_t* __restrict__ tiemask,
                bf16* __restrict__ dPQ, int act, int ds_plane_rows) {
  // tiecnt == nullptr: the tie counts are derived from the per-edge tie bits (the fused forward records only those);
  // ds_plane_rows > 0: dS[:, D:] arrives as D/32 planes of [ds_plane_rows x 32] (fused atom backward), else as the
  // second half of the [rows, 2D] matrix.
  constexpr int kD = Cfg<CH>::D, kSlots = Cfg<CH>::kSlots, kPasses = Cfg<CH>::kPasses;
  constexpr int kRowBytes = Cfg<CH>::kRowBytes, kBoxBytes = Cfg<CH>::kBoxBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem = (gcb::tc::smem_u32(smem_raw) + 127u) & ~127u;  // g tile [128][2 D bytes]
  const uint32_t s_n4 = smem + kBoxBytes, s_a4 = s_n4 + kRows * 16, s_ptr = s_a4 + kRows * 16;
  const uint32_t s_dst = s_ptr + 544;  // [128] destination atom of bond row r, or -1 when r has no in-edge
  const uint32_t s_inv = s_dst + kRows * 4;  // [256] 1 / tie count (1 for counts 0 and 1), the factor of the even tie split
  const uint32_t s_ie = s_inv + 256 * 4;     // [129] dst-CSR row pointers of the tile's rows (tie bits of a row's in-edges)
  const uint32_t s_pad = s_ie + 544;         // one zero row: what a missing ELL slot gathers
  const uint32_t s_tm = s_pad + kRowBytes;   // tie bits of the tile's in-edges (CH bytes per edge, at most kTieEdges edges)
  const uint32_t kPadOff = s_pad - smem;
  for (int i = threadIdx.x; i < kRowBytes / 16; i += kThreads) gcb::tc::sts128(s_pad + i * 16, make_uint4(0u, 0u, 0u, 0u));
  const int tb_ = tile_of_cta(cm);
  const int r0 = __ldg(cm.tile_ptr + tb_);
  const int r1 = min(__ldg(cm.tile_ptr + tb_ + 1), cm.rows);
  pdl_launch_dependents();
  if ((cm.pf & 1) && threadIdx.x == 32 && r1 > r0) {  // phase 1 streams these rows of the tile: ask L2 for them now
    prefetch_l2_rows(B_l + (size_t)r0 * kD, (r1 - r0) * kRowBytes);
    if (dBnext) prefetch_l2_rows(dBnext + (size_t)r0 * kD, (r1 - r0) * kRowBytes);
    if (d_out_bond) prefetch_l2_rows(d_out_bond + (size_t)r0 * kD, (r1 - r0) * kRowBytes);
    if (tiecnt) prefetch_l2_rows(tiecnt + (size_t)r0 * kD, (r1 - r0) * kD);
  }
  // The out-edges of a closed tile's rows are in-edges of rows of the SAME tile, and those are contiguous in the
  // dst-sorted CSR: [ie_lo, ie_hi).  Their tie bits (written by the forward pass) are staged with coalesced loads, so
  // phase 2 reads them from shared memory instead of four dependent one-byte global loads per row chunk.
  const int ie_lo = __ldg(in_ptr + r0), ie_hi = __ldg(in_ptr + max(r1, r0));
  const bool tm_staged = ie_hi - ie_lo <= kTieEdges;
  if (tm_staged) {
    const uint4* src = reinterpret_cast<const uint4*>(tiemask + (size_t)ie_lo * CH);
    const int n16 = (ie_hi - ie_lo) * (CH / 16);
    for (int i = threadIdx.x; i < n16; i += kThreads) gcb::tc::sts128(s_tm + i * 16, __ldg(src + i));
  }
  // x / cnt of the tie split as x * (1 / cnt): the same MUFU reciprocal __fdividef() uses, looked up instead of recomputed
  for (int i = threadIdx.x; i < 256; i += kThreads) sts32(s_inv + i * 4, __float_as_int(i <= 1 ? 1.f : __fdividef(1.f, (float)i)));
  for (int i = threadIdx.x; i <= kRows; i += kThreads) {
    const int r = r0 + i;
    if (i < kRows) {
      int4 v = make_int4(-1, -1, -1, -1), w = v;
      int t = -1;
      if (r < r1) {
        v = __ldg(cm.nbr4 + r);
        w = __ldg(cm.aux4 + r);
        if (__ldg(in_ptr + r + 1) > __ldg(in_ptr + r)) t = __ldg(dst + r);
        if ((cm.pf & 2) && t >= 0 && ds_plane_rows <= 0) {  // the row of another molecule this bond row collects its gradient from
          const bf16* q = dS + (size_t)t * (2 * kD) + kD;
          prefetch_l2_line(q);
          if (kRowBytes > 128) prefetch_l2_line(q + 64);
          if (kRowBytes > 256) { prefetch_l2_line(q + 128); prefetch_l2_line(q + 192); }
        }
      }
      // neighbours as byte offsets of their split-gradient rows; a missing slot points at the zero pad row
      const uint32_t o0 = v.x >= 0 ? (uint32_t)(v.x - r0) * kRowBytes : kPadOff, o1 = v.y >= 0 ? (uint32_t)(v.y - r0) * kRowBytes : kPadOff;
      const uint32_t o2 = v.z >= 0 ? (uint32_t)(v.z - r0) * kRowBytes : kPadOff, o3 = v.w >= 0 ? (uint32_t)(v.w - r0) * kRowBytes : kPadOff;
      gcb::tc::sts128(s_n4 + i * 16, make_uint4(o0, o1, o2, o3));
      gcb::tc::sts128(s_a4 + i * 16, make_uint4(w.x, w.y, w.z, w.w));
      sts32(s_dst + i * 4, t);
    }
    sts32(s_ptr + i * 4, r <= r1 ? __ldg(cm.ptr + r) : 0);
    sts32(s_ie + i * 4, r <= r1 ? __ldg(in_ptr + r) : 0);
  }
  const int grp = threadIdx.x / CH, lig = threadIdx.x % CH, c = lig * 8;
  pdl_wait();
  __syncthreads();
  // ---- phase 1 ----
#pragma unroll (UN)
  for (int ps = 0; ps < kPasses; ++ps) {
    const int rr = ps * kSlots + grp;
    const int r = r0 + rr;
    if (r >= r1) continue;
    const int t = lds32(s_dst + rr * 4);
    float d[8], g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = g[i] = 0.f;
    if (t >= 0) {
      const uint4 sraw = ds_plane_rows > 0 ? R::load(dS + ((size_t)(c >> 5) * ds_plane_rows + t) * 32 + (c & 31))
                                           : R::load(dS + (size_t)t * (2 * kD) + kD + c);
      const uint4 yraw = R::load(B_l + (size_t)r * kD + c);
      uint2 cw = make_uint2(0x01010101u, 0x01010101u);
      if (tiecnt) {
        cw = __ldg(reinterpret_cast<const uint2*>(tiecnt + (size_t)r * kD + c));
      } else {
        // count per channel how many in-edges tie for the max: channels seen twice are rare after the first layer
        const int ie0 = lds32(s_ie + rr * 4), ie1 = lds32(s_ie + rr * 4 + 4);
        uint32_t seen = 0u, twice = 0u;
        for (int e = ie0; e < ie1; ++e) {
          const uint32_t b = __ldg(tiemask + (size_t)e * CH + lig);
          twice |= seen & b;
          seen |= b;
        }
        if (twice) {
          uint32_t cnt[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) cnt[i] = 0u;
          for (int e = ie0; e < ie1; ++e) {
            const uint32_t b = __ldg(tiemask + (size_t)e * CH + lig);
#pragma unroll
            for (int i = 0; i < 8; ++i) cnt[i] += (b >> i) & 1u;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) cnt[i] = cnt[i] > 255u ? 255u : cnt[i];
          cw.x = cnt[0] | (cnt[1] << 8) | (cnt[2] << 16) | (cnt[3] << 24);
          cw.y = cnt[4] | (cnt[5] << 8) | (cnt[6] << 16) | (cnt[7] << 24);
        }
      }
      uint4 nraw = R::zero(), oraw = R::zero();
      if (dBnext) nraw = R::load(dBnext + (size_t)r * kD + c);
      if (d_out_bond) oraw = R::load(d_out_bond + (size_t)r * kD + c);
      R::to_float(sraw, d);
      if (dBnext) v2::raw_add<bf16, 8>(nraw, d);
      if (d_out_bond) v2::raw_add<bf16, 8>(oraw, d);
      float y[8];
      R::to_float(yraw, y);
      act_grad_mul_vec<bf16, 8>(d, y, act);
      const uint32_t cwv[2] = {cw.x, cw.y};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t cnt = (cwv[i >> 2] >> ((i & 3) * 8)) & 0xFFu;
        g[i] = d[i] * __int_as_float(lds32(s_inv + cnt * 4));
      }
    }
    store_vec(dPQ + (size_t)r * (2 * kD) + c, d);
    uint4 gp;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&gp);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(g[2 * i], g[2 * i + 1]);
    gcb::tc::sts128(smem + rr * kRowBytes + lig * 16, gp);
  }
  __syncthreads();
  // ---- phase 2 ----
#pragma unroll (UN)
  for (int ps = 0; ps < kPasses; ++ps) {
    const int rr = ps * kSlots + grp;
    const int r = r0 + rr;
    if (r >= r1) continue;
    const uint4 n4_ = lds16(s_n4 + rr * 16);
    const uint32_t noff[4] = {n4_.x, n4_.y, n4_.z, n4_.w};
    const int e0 = lds32(s_ptr + rr * 4), e1 = lds32(s_ptr + rr * 4 + 4);
    const uint4 o4_ = lds16(s_a4 + rr * 16);
    const int oidx[4] = {(int)o4_.x, (int)o4_.y, (int)o4_.z, (int)o4_.w};
    uint4 gq[4];
    uint32_t bits[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      // a missing slot (-1) gathers the zero pad row: whatever bits it reads do not matter
      if (tm_staged) bits[u] = lds8(s_tm + (uint32_t)max(oidx[u] - ie_lo, 0) * CH + lig);
      else bits[u] = __ldg(tiemask + (size_t)max(oidx[u], 0) * CH + lig);
      gq[u] = lds16(smem + noff[u] + lig * 16);
    }
    float acc[8];
    v2::vzero<bf16, 8>(acc);
#pragma unroll
    for (int u = 0; u < 4; ++u) v2::raw_add_masked(gq[u], bits[u], acc);
    for (int e = e0 + 4; e < e1; ++e)
      v2::raw_add_masked(lds16(smem + ((__ldg(cm.nbr + e) - r0) & 127) * kRowBytes + lig * 16),
                         __ldg(tiemask + (size_t)__ldg(cm.aux + e) * CH + lig), acc);
    store_vec(dPQ + (size_t)r * (2 * kD) + kD + c, acc);
  }
}

// ---- launchers (D = 128 -> CH 16, D = 256 -> CH 32) ------------------------------------------------
inline bool width_ok(int D) { return D == 128 || D == 256; }
// A/B switch (GCB_TILE_UNROLL=4): passes unrolled four deep in the latency-bound kernels (twice the loads in flight per thread)
inline int tile_unroll() {
  const char* e = std::getenv("GCB_TILE_UNROLL");
  return (e && e[0] == '4') ? 4 : 2;
}

template <int CH>
inline int smem_for(int boxes) { return boxes * Cfg<CH>::kBoxBytes + kIdxBytes + 128; }

template <typename K>
inline int set_smem(K kernel, int bytes) {
  GCB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return GCB_OK;
}

template <int CH>
inline int launch_bond_update_w(const bf16* PQ, const TileCommon& cm, int n_tiles, bf16* Bout, uint8_t* tiemask,
                                uint8_t* tiecnt, int act, cudaStream_t s) {
  constexpr int D = Cfg<CH>::D;
  // Q-only staging makes this kernel ~6 % faster in isolation but the whole step 1 % slower on B200 at D = 128 (the P
  // rows it then streams through L1 compete with the next kernel's operands in L2): opt-in there.  At D = 256 the
  // [P|Q] box (512 columns) exceeds a TMA box, so Q-only is the only form.
  static const bool q_env = std::getenv("GCB_TILE_Q_ONLY") && std::getenv("GCB_TILE_Q_ONLY")[0] == '1';
  const bool qonly = q_env || CH == 32;
  CUtensorMap tm;
  if (gcb::tc::make_tmap_bf16_plain(&tm, PQ, cm.rows, 2 * D, 2 * D, kRows, qonly ? D : 2 * D)) return GCB_ERR_CUDA;
  const int smem = smem_for<CH>(qonly ? 1 : 2);
  static bool attr = false;
  if (!attr) {
    GCB_TRY(set_smem(bond_update_kernel<CH, true, true>, smem_for<CH>(1))); GCB_TRY(set_smem(bond_update_kernel<CH, false, true>, smem_for<CH>(1)));
    if constexpr (CH == 16) {
      GCB_TRY(set_smem(bond_update_kernel<CH, true, false>, smem_for<CH>(2))); GCB_TRY(set_smem(bond_update_kernel<CH, false, false>, smem_for<CH>(2)));
    }
    attr = true;
  }
  if (qonly) {
    if (tiemask) GCB_KERNEL("tile_bond_update_kernel", s, launch_pdl(bond_update_kernel<CH, true, true>, n_tiles, kThreads, smem, s, tm, cm, PQ, Bout, tiemask, tiecnt, act));
    else GCB_KERNEL("tile_bond_update_kernel", s, launch_pdl(bond_update_kernel<CH, false, true>, n_tiles, kThreads, smem, s, tm, cm, PQ, Bout, tiemask, tiecnt, act));
  } else if constexpr (CH == 16) {
    if (tiemask) GCB_KERNEL("tile_bond_update_kernel", s, launch_pdl(bond_update_kernel<CH, true, false>, n_tiles, kThreads, smem, s, tm, cm, PQ, Bout, tiemask, tiecnt, act));
    else GCB_KERNEL("tile_bond_update_kernel", s, launch_pdl(bond_update_kernel<CH, false, false>, n_tiles, kThreads, smem, s, tm, cm, PQ, Bout, tiemask, tiecnt, act));
  }
  return GCB_OK;
}
inline int launch_bond_update(int D, const bf16* PQ, const TileCommon& cm, int n_tiles, bf16* Bout, uint8_t* tiemask,
                              uint8_t* tiecnt, int act, cudaStream_t s) {
  return D == 128 ? launch_bond_update_w<16>(PQ, cm, n_tiles, Bout, tiemask, tiecnt, act, s)
                  : launch_bond_update_w<32>(PQ, cm, n_tiles, Bout, tiemask, tiecnt, act, s);
}

// Shared memory of bond_update_tok_kernel; the caller falls back to GEMM + bond_update when the table does not fit.
inline int bond_tok_smem(int D, int vocab) { return (vocab + 1) * 4 * D + kRows * 16 + kRows * 4 + 544 + 128; }
inline bool bond_tok_ok(int D, int vocab) { return width_ok(D) && bond_tok_smem(D, vocab) <= 72 * 1024; }

template <int CH>
inline int launch_bond_update_tok_w(const TileCommon& cm, int n_tiles, const bf16* tabPQ, const int32_t* tok, int vocab,
                                    bf16* Bout, uint8_t* tiemask, uint8_t* tiecnt, int act, cudaStream_t s) {
  const int smem = bond_tok_smem(Cfg<CH>::D, vocab);
  static bool attr = false;
  if (!attr) {
    GCB_TRY(set_smem(bond_update_tok_kernel<CH, true, 4>, 72 * 1024)); GCB_TRY(set_smem(bond_update_tok_kernel<CH, false, 4>, 72 * 1024));
    attr = true;
  }
  // four resident CTAs (<= 64 registers): 42.5 vs 46.5 us with three (profiles/r2_fused_ab.jsonl)
  if (tiemask) GCB_KERNEL("tile_bond_update_tok_kernel", s, launch_pdl(bond_update_tok_kernel<CH, true, 4>, n_tiles, kThreads, smem, s, cm, tabPQ, tok, vocab, Bout, tiemask, tiecnt, act));
  else GCB_KERNEL("tile_bond_update_tok_kernel", s, launch_pdl(bond_update_tok_kernel<CH, false, 4>, n_tiles, kThreads, smem, s, cm, tabPQ, tok, vocab, Bout, tiemask, tiecnt, act));
  return GCB_OK;
}
inline int launch_bond_update_tok(int D, const TileCommon& cm, int n_tiles, const bf16* tabPQ, const int32_t* tok, int vocab,
                                  bf16* Bout, uint8_t* tiemask, uint8_t* tiecnt, int act, cudaStream_t s) {
  return D == 128 ? launch_bond_update_tok_w<16>(cm, n_tiles, tabPQ, tok, vocab, Bout, tiemask, tiecnt, act, s)
                  : launch_bond_update_tok_w<32>(cm, n_tiles, tabPQ, tok, vocab, Bout, tiemask, tiecnt, act, s);
}

template <int CH>
inline int launch_atom_gather_w(const bf16* A, const bf16* Bnew, int M, const TileCommon& cm, int n_tiles, bf16* S, int act,
                                cudaStream_t s) {
  constexpr int D = Cfg<CH>::D;
  CUtensorMap tm;
  if (gcb::tc::make_tmap_bf16_plain(&tm, A, cm.rows, D, D, kRows, D)) return GCB_ERR_CUDA;
  const int smem = smem_for<CH>(1);
  static bool attr = false;
  if (!attr) { GCB_TRY(set_smem(atom_gather_kernel<CH, 2>, smem)); GCB_TRY(set_smem(atom_gather_kernel<CH, 4>, smem)); attr = true; }
  if (tile_unroll() == 4) GCB_KERNEL("tile_atom_gather_kernel", s, launch_pdl(atom_gather_kernel<CH, 4>, n_tiles, kThreads, smem, s, tm, cm, Bnew, M, S, act));
  else GCB_KERNEL("tile_atom_gather_kernel", s, launch_pdl(atom_gather_kernel<CH, 2>, n_tiles, kThreads, smem, s, tm, cm, Bnew, M, S, act));
  return GCB_OK;
}
inline int launch_atom_gather(int D, const bf16* A, const bf16* Bnew, int M, const TileCommon& cm, int n_tiles, bf16* S, int act,
                              cudaStream_t s) {
  return D == 128 ? launch_atom_gather_w<16>(A, Bnew, M, cm, n_tiles, S, act, s)
                  : launch_atom_gather_w<32>(A, Bnew, M, cm, n_tiles, S, act, s);
}

// Shared memory of atom_gather_tok_kernel; the caller materialises A_0 when the table does not fit.
inline int atom_tok_smem(int D, int vocab) { return (vocab + 1) * 2 * D + 2 * kRows * 16 + 544 + 128; }
inline bool atom_tok_ok(int D, int vocab) { return width_ok(D) && atom_tok_smem(D, vocab) <= 48 * 1024; }

template <int CH>
inline int launch_atom_gather_tok_w(const TileCommon& cm, int n_tiles, const bf16* tabA, const int32_t* tok, int vocab, const bf16* Bnew,
                                    int M, bf16* S, int act, cudaStream_t s) {
  const int smem = atom_tok_smem(Cfg<CH>::D, vocab);
  static bool attr = false;
  if (!attr) { GCB_TRY(set_smem(atom_gather_tok_kernel<CH, 2>, 48 * 1024)); attr = true; }
  GCB_KERNEL("tile_atom_gather_tok_kernel", s, launch_pdl(atom_gather_tok_kernel<CH, 2>, n_tiles, kThreads, smem, s, cm, tabA, tok, vocab, Bnew, M, S, act));
  return GCB_OK;
}
inline int launch_atom_gather_tok(int D, const TileCommon& cm, int n_tiles, const bf16* tabA, const int32_t* tok, int vocab, const bf16* Bnew,
                                  int M, bf16* S, int act, cudaStream_t s) {
  return D == 128 ? launch_atom_gather_tok_w<16>(cm, n_tiles, tabA, tok, vocab, Bnew, M, S, act, s)
                  : launch_atom_gather_tok_w<32>(cm, n_tiles, tabA, tok, vocab, Bnew, M, S, act, s);
}

template <int CH>
inline int launch_atom_bwd_w(const bf16* dS, const bf16* dZA, const bf16* A_prev, const int32_t* prev_idx, const TileCommon& cm, int n_tiles,
                             bf16* dZA_prev, int act, cudaStream_t s) {
  constexpr int D = Cfg<CH>::D;
  CUtensorMap tmS;
  if (gcb::tc::make_tmap_bf16_plain(&tmS, dS, cm.rows, D, 2 * D, kRows, D)) return GCB_ERR_CUDA;
  const int smem = smem_for<CH>(1);
  static bool attr = false;
  if (!attr) { GCB_TRY(set_smem(atom_bwd_kernel<CH, 2>, smem)); GCB_TRY(set_smem(atom_bwd_kernel<CH, 4>, smem)); attr = true; }
  TileCommon cp = cm;
  cp.pf = tile_prefetch();
  if (tile_unroll() == 4) GCB_KERNEL("tile_atom_bwd_kernel", s, launch_pdl(atom_bwd_kernel<CH, 4>, n_tiles, kThreads, smem, s, tmS, cp, dZA, A_prev, prev_idx, dZA_prev, act));
  else GCB_KERNEL("tile_atom_bwd_kernel", s, launch_pdl(atom_bwd_kernel<CH, 2>, n_tiles, kThreads, smem, s, tmS, cp, dZA, A_prev, prev_idx, dZA_prev, act));
  return GCB_OK;
}
inline int launch_atom_bwd(int D, const bf16* dS, const bf16* dZA, const bf16* A_prev, const int32_t* prev_idx, const TileCommon& cm, int n_tiles,
                           bf16* dZA_prev, int act, cudaStream_t s) {
  return D == 128 ? launch_atom_bwd_w<16>(dS, dZA, A_prev, prev_idx, cm, n_tiles, dZA_prev, act, s)
                  : launch_atom_bwd_w<32>(dS, dZA, A_prev, prev_idx, cm, n_tiles, dZA_prev, act, s);
}

template <int CH>
inline int launch_bond_scatter_w(const bf16* gsplit, const TileCommon& cm, int n_tiles, const uint8_t* tiemask, bf16* dPQ,
                                 cudaStream_t s) {
  constexpr int D = Cfg<CH>::D;
  CUtensorMap tm;
  if (gcb::tc::make_tmap_bf16_plain(&tm, gsplit, cm.rows, D, D, kRows, D)) return GCB_ERR_CUDA;
  const int smem = smem_for<CH>(1);
  static bool attr = false;
  if (!attr) { GCB_TRY(set_smem(bond_scatter_kernel<CH>, smem)); attr = true; }
  GCB_KERNEL("tile_bond_scatter_kernel", s, launch_pdl(bond_scatter_kernel<CH>, n_tiles, kThreads, smem, s, tm, cm, tiemask, dPQ));
  return GCB_OK;
}
inline int launch_bond_scatter(int D, const bf16* gsplit, const TileCommon& cm, int n_tiles, const uint8_t* tiemask, bf16* dPQ,
                               cudaStream_t s) {
  return D == 128 ? launch_bond_scatter_w<16>(gsplit, cm, n_tiles, tiemask, dPQ, s)
                  : launch_bond_scatter_w<32>(gsplit, cm, n_tiles, tiemask, dPQ, s);
}

template <int CH>
inline int launch_bond_bwd_w(const TileCommon& cm, int n_tiles, const int32_t* in_ptr, const int32_t* dst, const bf16* dS,
                             const bf16* dBnext, const bf16* d_out_bond, const bf16* B_l, const uint8_t* tiecnt,
                             const uint8_t* tiemask, bf16* dPQ, int act, int ds_plane_rows, cudaStream_t s) {
  const int smem = smem_for<CH>(1) + kRows * 4 + 256 * 4 + 544 + Cfg<CH>::kRowBytes + kTieEdges * CH;
  static bool attr = false;
  if (!attr) { GCB_TRY(set_smem(bond_bwd_kernel<CH, 2>, smem)); GCB_TRY(set_smem(bond_bwd_kernel<CH, 4>, smem)); attr = true; }
  TileCommon cp = cm;
  cp.pf = tile_prefetch();
  if (tile_unroll() == 4) GCB_KERNEL("tile_bond_bwd_kernel", s, launch_pdl(bond_bwd_kernel<CH, 4>, n_tiles, kThreads, smem, s, cp, in_ptr, dst, dS, dBnext, d_out_bond, B_l,
                                                                        tiecnt, tiemask, dPQ, act, ds_plane_rows));
  else GCB_KERNEL("tile_bond_bwd_kernel", s, launch_pdl(bond_bwd_kernel<CH, 2>, n_tiles, kThreads, smem, s, cp, in_ptr, dst, dS, dBnext, d_out_bond, B_l,
                                                          tiecnt, tiemask, dPQ, act, ds_plane_rows));
  return GCB_OK;
}
inline int launch_bond_bwd(int D, const TileCommon& cm, int n_tiles, const int32_t* in_ptr, const int32_t* dst, const bf16* dS,
                           const bf16* dBnext, const bf16* d_out_bond, const bf16* B_l, const uint8_t* tiecnt,
                           const uint8_t* tiemask, bf16* dPQ, int act, int ds_plane_rows, cudaStream_t s) {
  return D == 128 ? launch_bond_bwd_w<16>(cm, n_tiles, in_ptr, dst, dS, dBnext, d_out_bond, B_l, tiecnt, tiemask, dPQ, act, ds_plane_rows, s)
                  : launch_bond_bwd_w<32>(cm, n_tiles, in_ptr, dst, dS, dBnext, d_out_bond, B_l, tiecnt, tiemask, dPQ, act, ds_plane_rows, s);
}

}  // namespace tile
}  // namespace gcb
