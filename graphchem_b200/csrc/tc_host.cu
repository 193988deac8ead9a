// Host-side helpers of the tcgen05 path: TMA tensor-map encoding through the driver entry point
// (resolved at run time: the library links only the static CUDA runtime, never libcuda directly).
#include <cstdlib>
#include <mutex>

#include "tc_common.cuh"

namespace gcb {

bool pdl_enabled() {
  static const bool off = std::getenv("GCB_NO_PDL") && std::getenv("GCB_NO_PDL")[0] == '1';
  return !off;
}

namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_cuda_error(cudaErrorNotSupported, __FILE__, __LINE__);
    return 1;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  const cuuint32_t estride[2] = {1u, 1u};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_cuda_error(cudaErrorInvalidValue, __FILE__, __LINE__);
    return 1;
  }
  return 0;
}

int make_tmap_bf16_plain(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                         int box_cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_cuda_error(cudaErrorNotSupported, __FILE__, __LINE__);
    return 1;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estride[2] = {1u, 1u};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_cuda_error(cudaErrorInvalidValue, __FILE__, __LINE__);
    return 1;
  }
  return 0;
}

bool tc_enabled() {
  const char* e = std::getenv("GCB_DISABLE_TC");
  return !(e && e[0] == '1');
}

// Fused tcgen05 GEMM + message kernels (fused_rows.cuh): OPT-IN.  Measured on B200 (batch 8192, D = 128, 4 steps,
// profiles/r2_fused_ab.jsonl) they lose to the tile kernels + stand-alone GEMMs (2.15 vs 1.77 ms per train step):
// their 16 row warps run drain -> gather -> activation -> copy-out in lock-step at 28 % issue utilisation.
// GCB_FUSED=1 switches all three on, GCB_FUSED_BOND / GCB_FUSED_ATOM / GCB_FUSED_ABWD one of them (which = 0 / 1 / 2);
// GCB_NO_FUSED=1 and GCB_NO_FUSED_BOND / _ATOM / _ABWD veto (A/B switches).
bool fused_enabled(int which) {
  static const char* on_names[3] = {"GCB_FUSED_BOND", "GCB_FUSED_ATOM", "GCB_FUSED_ABWD"};
  static const char* off_names[3] = {"GCB_NO_FUSED_BOND", "GCB_NO_FUSED_ATOM", "GCB_NO_FUSED_ABWD"};
  auto set = [](const char* name) { const char* e = std::getenv(name); return e && e[0] == '1'; };
  if (set("GCB_NO_FUSED") || set(off_names[which])) return false;
  return set("GCB_FUSED") || set(on_names[which]);
}
// GCB_FUSED_EXACT=1: the fused bond forward rounds P to bf16 as the unfused pipeline stores it (bit-identical
// results, tests/test_fused_gpu.py); by default P stays in fp32 (closer to the reference, fewer instructions).
bool fused_exact() {
  const char* e = std::getenv("GCB_FUSED_EXACT");
  return e && e[0] == '1';
}

// Tile-staged row kernels (tile_rows.cuh); on unless GCB_NO_TILE=1.
bool tile_enabled() {
  const char* e = std::getenv("GCB_NO_TILE");
  return !(e && e[0] == '1');
}

// One tile kernel for both sides of the bond backward (GCB_NO_BOND_MERGE=1: tie kernel + tile scatter).
bool bond_bwd_merged_enabled() {
  const char* e = std::getenv("GCB_NO_BOND_MERGE");
  return !(e && e[0] == '1');
}

bool overlap_enabled() {
  const char* e = std::getenv("GCB_NO_OVERLAP");
  return !(e && e[0] == '1');
}

bool pingpong_enabled() {
  const char* e = std::getenv("GCB_NO_PINGPONG");
  return !(e && e[0] == '1');
}

}  // namespace tc
}  // namespace gcb
