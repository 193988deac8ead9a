// Orchestration of the MoleculeGCN hot path behind the C-ABI of include/graphchem_b200.h:
// parameter layout, derived weights, workspace plan, forward, backward, loss and optimiser.
//
// Reference path being replaced: graphchem/nn/gcn.py:120-202 (forward), autograd through it
// (examples/predict_cn.ipynb:250), F.mse_loss (:249) and torch.optim.Adam.step (:251).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "tile_rows.cuh"
#include "fused_rows.cuh"
#include "kernels.cuh"
#include "readout.cuh"
#include "rowops.cuh"
#include "aggr_max.cuh"

namespace gcb {

// ---- error string (thread-local; the only state the library keeps) ------------------------------
static thread_local char g_err[512] = "";
void set_cuda_error(cudaError_t e, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "%s (%s) at %s:%d", cudaGetErrorName(e), cudaGetErrorString(e), file, line);
}

// ---- kernel timing / launch counting (thread-local debug facility) -------------------------------
struct TimingSlot { const char* name; cudaEvent_t e0, e1; };
static thread_local bool g_timing = false;
static thread_local int64_t g_launches = 0;
static thread_local std::vector<TimingSlot>* g_slots = nullptr;
KernelTimer::KernelTimer(const char* name, cudaStream_t s) : name_(name), s_(s), slot_(-1) {
  ++g_launches;
  if (g_timing) {
    if (!g_slots) g_slots = new std::vector<TimingSlot>();
    TimingSlot t{name, nullptr, nullptr};
    if (cudaEventCreate(&t.e0) == cudaSuccess && cudaEventCreate(&t.e1) == cudaSuccess) {
      cudaEventRecord(t.e0, s);
      g_slots->push_back(t);
      slot_ = (int)g_slots->size() - 1;
    }
  }
}
void KernelTimer::stop() {
  if (slot_ >= 0) cudaEventRecord((*g_slots)[slot_].e1, s_);
}

// ---- GEMM dispatch: tcgen05 kernel for bf16 hot shapes, CUDA-core kernel otherwise ---------------
namespace tc { bool tc_enabled(); bool pingpong_enabled(); bool overlap_enabled(); bool fused_enabled(int which); bool fused_exact(); bool tile_enabled(); bool bond_bwd_merged_enabled(); }
template <typename T>
static int gemm_nt_auto(const T* A, int lda, const T* W, int ldw, bool w_kn, T* C, int ldc, int M, int N, int K,
                        const Epilogue& ep, cudaStream_t s) {
  if constexpr (sizeof(T) == 2) {
    if (!w_kn && ldw == K && !ep.actgrad_y && tc::tc_enabled()) {
      tc::TcEpilogue te;
      te.bias = ep.bias; te.rowptr = ep.rowptr; te.aggr = ep.aggr;
      te.res = (const bf16*)ep.res; te.ld_res = ep.ld_res; te.res_idx = ep.res_idx; te.act = ep.act; te.rev = ep.rev;
      const int rc = tc::launch_gemm_nt_tc(A, lda, W, C, ldc, M, N, K, te, s);
      if (rc != GCB_ERR_UNSUPPORTED) return rc;
    }
  }
  return launch_gemm_nt<T>(A, lda, W, ldw, w_kn, C, ldc, M, N, K, ep, s);
}

// `defer_partial` (tcgen05 path only): keep this GEMM's per-CTA partials in that buffer, summed across
// calls (`defer_first` starts a new sum), and report the CTA count in *defer_grid; the caller runs ONE
// ordered reduction after the last call instead of one per call.  *defer_grid stays 0 when the
// CUDA-core path (which reduces immediately into dW/db with `accumulate`) had to take the call.
template <typename T>
static int gemm_tn_auto(const T* A, int lda, const T* B, int ldb, int rows, int Nn, int Kk, float* partial, float* dW,
                        float* db, const int32_t* rowptr, int aggr, int accumulate, cudaStream_t s, int rev = 0,
                        float* defer_partial = nullptr, int* defer_grid = nullptr, int defer_first = 1) {
  if constexpr (sizeof(T) == 2) {
    if (tc::tc_enabled()) {
      int grid = 0;
      // measured on B200 (batch 8192): the read-modify-write of the partials costs more than the reductions it
      // saves (1.867 vs 1.852 ms/step), so this stays opt-in
      static const bool defer_on = std::getenv("GCB_TN_DEFER") && std::getenv("GCB_TN_DEFER")[0] == '1';
      const bool defer = defer_on && defer_partial != nullptr && defer_grid != nullptr && db != nullptr && db == dW + (int64_t)Nn * Kk;
      const int rc = tc::launch_gemm_tn_tc(A, lda, B, ldb, rows, Nn, Kk, defer ? defer_partial : partial, db != nullptr, rowptr,
                                           aggr, rev, &grid, s, defer && !defer_first ? 1 : 0);
      if (rc == GCB_OK) {
        if (defer) { *defer_grid = grid; return GCB_OK; }
        const int64_t wcount = (int64_t)Nn * Kk, stride = wcount + (db ? Nn : 0);
        if (db && db == dW + wcount) {  // weight and bias accumulators are contiguous: one ordered reduction
          GCB_TRY(launch_reduce_partials(partial, grid, stride, dW, accumulate, s, stride));
        } else {
          GCB_TRY(launch_reduce_partials(partial, grid, wcount, dW, accumulate, s, stride));
          if (db) GCB_TRY(launch_reduce_partials(partial + wcount, grid, Nn, db, accumulate, s, stride));
        }
        return GCB_OK;
      }
      if (rc != GCB_ERR_UNSUPPORTED) return rc;
    }
  }
  return launch_gemm_tn<T>(A, lda, B, ldb, rows, Nn, Kk, partial, dW, db, rowptr, aggr, accumulate, s);
}

// ---- side stream for off-critical-path work (weight-gradient GEMMs of the backward pass) ----------
// Internal, per host thread, created lazily; joins the caller's stream through events only, so it is
// captured together with it when the caller is recording a CUDA graph.
struct SideStream {
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> events;
  bool ok = false;
  cudaEvent_t event(size_t i) {
    while (events.size() <= i) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
      events.push_back(e);
    }
    return events[i];
  }
};
static SideStream* side_stream() {
  static thread_local SideStream ss;
  if (!ss.ok) {
    if (cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    ss.ok = true;
  }
  return &ss;
}

// ---- parameter layout ---------------------------------------------------------------------------
struct ParamLayout {
  int n_tensors = 0;
  int64_t off[GCB_MAX_PARAM_TENSORS + 1];
  // named offsets
  int64_t emb_atom, emb_bond, w_msg, b_msg, w_edge, b_edge, w_bond, b_bond;
  int64_t r_w[32], r_b[32];
  int r_in[32], r_out[32];
  int n_lin = 0;  // number of readout Linear layers (n_readout + 1, or 0)
  int64_t total = 0;
};

static int make_layout(const gcb_model_cfg& c, ParamLayout& p) {
  if (c.D <= 0 || c.atom_vocab <= 0 || c.bond_vocab <= 0 || c.L < 0 || c.n_readout < 0) return GCB_ERR_INVALID_ARG;
  if (c.n_readout > 30) return GCB_ERR_UNSUPPORTED;
  if (c.n_readout > 0 && (c.R <= 0 || c.out_dim <= 0)) return GCB_ERR_INVALID_ARG;
  const int64_t D = c.D;
  int64_t o = 0;
  int n = 0;
  auto add = [&](int64_t& slot, int64_t count) { slot = o; p.off[n++] = o; o += count; };
  add(p.emb_atom, (int64_t)c.atom_vocab * D);
  add(p.emb_bond, (int64_t)c.bond_vocab * D);
  add(p.w_msg, D * D); add(p.b_msg, D);
  add(p.w_edge, D * D); add(p.b_edge, D);
  add(p.w_bond, D * 2 * D); add(p.b_bond, D);
  p.n_lin = c.n_readout > 0 ? c.n_readout + 1 : 0;
  for (int k = 0; k < p.n_lin; ++k) {
    p.r_in[k] = k == 0 ? c.D : c.R;
    p.r_out[k] = k == p.n_lin - 1 ? c.out_dim : c.R;
    add(p.r_w[k], (int64_t)p.r_in[k] * p.r_out[k]);
    add(p.r_b[k], p.r_out[k]);
  }
  p.off[n] = o;
  p.n_tensors = n;
  p.total = o;
  return GCB_OK;
}

static int check_cfg(const gcb_model_cfg& c) {
  if (c.dtype != GCB_F32 && c.dtype != GCB_BF16) return GCB_ERR_INVALID_ARG;
  if (c.act < 0 || c.act > GCB_ACT_LEAKY_RELU) return GCB_ERR_INVALID_ARG;
  if (c.aggr != GCB_AGGR_ADD && c.aggr != GCB_AGGR_MEAN && c.aggr != GCB_AGGR_MAX) return GCB_ERR_UNSUPPORTED;
  if (c.pool != GCB_POOL_ADD && c.pool != GCB_POOL_MEAN) return GCB_ERR_INVALID_ARG;
  if (c.D % 8 != 0) return GCB_ERR_UNSUPPORTED;  // 16-byte channel chunks in both dtypes
  if (!(c.p_dropout >= 0.f && c.p_dropout < 1.f)) return GCB_ERR_UNSUPPORTED;
  if (c.L > 250) return GCB_ERR_UNSUPPORTED;
  return GCB_OK;
}

// ---- derived weights ----------------------------------------------------------------------------
struct Derived {
  char* base;
  int64_t wpq, wat, watT, wpqT, bpq, bat, tab_atom, tab_bond, tab_pq, total;
};
static Derived derived_layout(const gcb_model_cfg& c, void* base) {
  const int64_t D = c.D, es = c.dtype == GCB_BF16 ? 2 : 4;
  Derived d;
  d.base = (char*)base;
  int64_t o = 0;
  auto add = [&](int64_t& slot, int64_t bytes) { slot = o; o += align_up(bytes, 256); };
  add(d.wpq, 2 * D * D * es); add(d.wat, 2 * D * D * es); add(d.watT, 2 * D * D * es); add(d.wpqT, 2 * D * D * es);
  add(d.bpq, 2 * D * 4); add(d.bat, D * 4);
  // pre-activated embedding tables act(emb)[vocab, D] in the activation dtype: the forward lookup is a row copy
  add(d.tab_atom, (int64_t)c.atom_vocab * D * es); add(d.tab_bond, (int64_t)c.bond_vocab * D * es);
  // [P|Q] of every bond token, tab_bond [U;V]^T + [b|0]: message step 1 reads it instead of a [rows, 2D] matrix
  add(d.tab_pq, (int64_t)c.bond_vocab * 2 * D * es);
  d.total = o;
  return d;
}

template <typename T>
__global__ void prepare_weights_kernel(const float* __restrict__ w_msg, const float* __restrict__ b_msg,
                                       const float* __restrict__ w_edge, const float* __restrict__ b_edge,
                                       const float* __restrict__ w_bond, const float* __restrict__ b_bond,
                                       T* __restrict__ wpq, T* __restrict__ wat, T* __restrict__ watT,
                                       T* __restrict__ wpqT, float* __restrict__ bpq, float* __restrict__ bat, int D,
                                       const float* __restrict__ emb_atom, const float* __restrict__ emb_bond,
                                       T* __restrict__ tab_atom, T* __restrict__ tab_bond, int n_atom, int n_bond, int act) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  // act(embedding) rounded to the storage type, exactly what embed + act produced per row (gcn.py:152-157)
  if (idx < n_atom) tab_atom[idx] = from_float<T>(act_fwd_t<T>(emb_atom[idx], act));
  if (idx < n_bond) tab_bond[idx] = from_float<T>(act_fwd_t<T>(emb_bond[idx], act));
  const int DD2 = 2 * D * D;
  if (idx < DD2) {
    // Wpq[n][k], n in [0,2D), k in [0,D):  U = W1 - W2 (n < D), V = W2 (n >= D);  W_b = [W1 | W2], [D, 2D]
    const int n = idx / D, k = idx % D;
    const float pq = n < D ? w_bond[(size_t)n * 2 * D + k] - w_bond[(size_t)n * 2 * D + D + k]
                           : w_bond[(size_t)(n - D) * 2 * D + D + k];
    wpq[idx] = from_float<T>(pq);
    wpqT[(size_t)k * 2 * D + n] = from_float<T>(pq);
    // Wat[o][j], o in [0,D), j in [0,2D): [W_msg | W_edge]
    const int o = idx / (2 * D), j = idx % (2 * D);
    const float at = j < D ? w_msg[(size_t)o * D + j] : w_edge[(size_t)o * D + (j - D)];
    wat[idx] = from_float<T>(at);
    watT[(size_t)j * D + o] = from_float<T>(at);
  }
  if (idx < 2 * D) bpq[idx] = idx < D ? b_bond[idx] : 0.f;
  if (idx < D) bat[idx] = b_msg[idx] + b_edge[idx];
}

// acc buffers -> flat gradient of the conv parameters.
__global__ void finalize_conv_grads_kernel(const float* __restrict__ accWpq, const float* __restrict__ accBpq,
                                           const float* __restrict__ accWat, const float* __restrict__ accBat,
                                           float* __restrict__ g_wmsg, float* __restrict__ g_bmsg,
                                           float* __restrict__ g_wedge, float* __restrict__ g_bedge,
                                           float* __restrict__ g_wbond, float* __restrict__ g_bbond, int D,
                                           int accumulate, int zero) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  auto put = [&](float* dst, float v) { *dst = accumulate ? *dst + v : v; };
  if (idx < 2 * D * D) {
    {  // bond weight [D, 2D]: dW1 = dU, dW2 = dV - dU
      const int o = idx / (2 * D), j = idx % (2 * D);
      float v = 0.f;
      if (!zero) v = j < D ? accWpq[(size_t)o * D + j] : accWpq[(size_t)(D + o) * D + (j - D)] - accWpq[(size_t)o * D + (j - D)];
      put(g_wbond + idx, v);
    }
    {  // [W_msg | W_edge] accumulated as [D, 2D]
      const int o = idx / (2 * D), j = idx % (2 * D);
      const float v = zero ? 0.f : accWat[idx];
      if (j < D) put(g_wmsg + (size_t)o * D + j, v);
      else put(g_wedge + (size_t)o * D + (j - D), v);
    }
  }
  if (idx < D) {
    const float vb = zero ? 0.f : accBat[idx];
    put(g_bmsg + idx, vb);
    put(g_bedge + idx, vb);
    put(g_bbond + idx, zero ? 0.f : accBpq[idx]);
  }
}

// Backward of message step 1 through the bond-token table.  tok[t][n] = sum over the bond rows with token t of
// [dP|dQ]_1 (per-token sums in bucket order, embed_bwd_partial + ordered reduction).  With B_0[r] = tab_bond[tok r]:
//   dW_pq   += tok^T tab_bond                    (what gemm_tn([dP|dQ]_1, B_0) sums row by row)
//   db_pq   += column sums of the dP half
//   d emb_bond[t] = (tok[t] W_pq) * act'(tab_bond[t])   (what the dB GEMM + embedding backward produce row by row)
// All sums run over t or n in index order: deterministic.
template <typename T>
__global__ void __launch_bounds__(256)
bond_table_bwd_kernel(const float* __restrict__ tok, const T* __restrict__ tab_bond, const T* __restrict__ wpq,
                      float* __restrict__ accWpq, float* __restrict__ accBpq, float* __restrict__ g_emb_bond, int bv, int D,
                      int act, int acc_first, int accumulate, int nb1) {
  const int D2 = 2 * D;
  if ((int)blockIdx.x < nb1) {
    // dW_pq and db_pq: one output per thread, bv terms each (coalesced over k)
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < D2 * D) {
      const int n = idx / D, k = idx % D;
      float s = 0.f;
#pragma unroll 8
      for (int t = 0; t < bv; ++t) s = fmaf(__ldg(tok + (size_t)t * D2 + n), to_float(tab_bond[(size_t)t * D + k]), s);
      accWpq[idx] = acc_first ? s : accWpq[idx] + s;
    }
    if (idx < D2) {
      float s = 0.f;
      if (idx < D) {
#pragma unroll 8
        for (int t = 0; t < bv; ++t) s += __ldg(tok + (size_t)t * D2 + idx);
      }
      accBpq[idx] = acc_first ? s : accBpq[idx] + s;
    }
    return;
  }
  // d emb_bond: one CTA per (token, 32 output columns); the 2D terms of every output are split over 8 thread rows
  // (32 consecutive n each) and combined in row order: a fixed summation tree
  __shared__ float red[8][33];
  const int b = blockIdx.x - nb1, kgroups = D / 32;
  const int t = b / kgroups, k = (b % kgroups) * 32 + (threadIdx.x & 31), sub = threadIdx.x >> 5;
  const int per = D2 / 8;
  float s = 0.f;
#pragma unroll 8
  for (int n = sub * per; n < (sub + 1) * per; ++n) s = fmaf(__ldg(tok + (size_t)t * D2 + n), to_float(wpq[(size_t)n * D + k]), s);
  red[sub][threadIdx.x & 31] = s;
  __syncthreads();
  if (sub == 0) {
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) tot += red[j][threadIdx.x];
    const int idx = t * D + k;
    tot *= act_grad_t<T>(to_float(tab_bond[idx]), act);
    g_emb_bond[idx] = accumulate ? g_emb_bond[idx] + tot : tot;
  }
}

// ---- workspace plan -----------------------------------------------------------------------------
constexpr int kMaxLayers = 256;

struct Plan {
  int64_t total = 0;
  char* base = nullptr;
  // forward (saved) tensors
  void* A[kMaxLayers];
  void* B[kMaxLayers];
  void* PQ[kMaxLayers];
  void* S[kMaxLayers];
  float* pooled = nullptr;
  float* H[33];
  // backward temporaries
  void* dZA[2];
  void* dS = nullptr;
  void* dBn[2];
  void* dPQ = nullptr;
  void* dPQ2[2];  // [dP|dQ] of alternate layers (lets the weight-gradient GEMM of layer l overlap layer l-1)
  void* mx = nullptr;
  void* gsplit = nullptr;
  // per layer (training): [E][32] per-edge, per-lane "this edge attains the max" bits; [M][D] tie counts
  uint8_t* tiemask[kMaxLayers];
  uint8_t* tiecnt[kMaxLayers];
  float* dH[2];
  float* dPool = nullptr;
  float* partial = nullptr;
  float *partialAt = nullptr, *partialPq = nullptr;  // per-CTA partials of the two conv weight gradients, summed over layers
  float *accWpq = nullptr, *accBpq = nullptr, *accWat = nullptr, *accBat = nullptr;
  float* tokPQ = nullptr;  // [bond_vocab][2D] per-token sums of step 1's [dP|dQ]
  // aggr = "max" (aggr_max.cuh): per-edge rows in dst-CSR order
  void* Msg[kMaxLayers];    // [E, D] messages of every layer (saved for the amax backward)
  void* Zrow = nullptr;     // [E, 2D] [A[src] | B'[eid]] (recomputed in backward)
  void* dMsg = nullptr;     // [E, D]
  void* dZrow = nullptr;    // [E, 2D]
  int32_t* einpos = nullptr;  // [E] dst-CSR position of edge e
};

static int make_plan(const gcb_model_cfg& c, const ParamLayout& pl, int64_t N, int64_t E, int64_t G, int64_t M,
                     int flags, void* base, Plan& p) {
  const int64_t D = c.D, es = c.dtype == GCB_BF16 ? 2 : 4;
  const bool save = (flags & GCB_SAVE_FOR_BWD) != 0;
  const int L = c.L;
  if (L + 1 > kMaxLayers) return GCB_ERR_UNSUPPORTED;
  const bool amax = c.aggr == GCB_AGGR_MAX;  // per-edge messages instead of the aggregated [N, 2D] operands
  const int64_t NS = amax ? 0 : N;           // rows of S / dS
  int64_t o = 0;
  auto take = [&](int64_t bytes) -> void* {
    void* ptr = base ? (char*)base + o : nullptr;
    o += align_up(bytes > 0 ? bytes : 1, 256);
    return ptr;
  };
  p.base = (char*)base;
  if (save) {
    for (int l = 0; l <= L; ++l) p.A[l] = take(N * D * es);
    for (int l = 0; l <= L; ++l) p.B[l] = take(M * D * es);
    for (int l = 1; l <= L; ++l) p.PQ[l] = take(M * 2 * D * es);
    for (int l = 1; l <= L; ++l) p.S[l] = take(NS * 2 * D * es);
    if (amax) for (int l = 1; l <= L; ++l) p.Msg[l] = take(E * D * es);
  } else {
    void* a2[2] = {take(N * D * es), take(N * D * es)};
    void* b2[2] = {take(M * D * es), take(M * D * es)};
    void* pq = take(M * 2 * D * es);
    void* s = take(NS * 2 * D * es);
    void* msg = amax ? take(E * D * es) : nullptr;
    for (int l = 0; l <= L; ++l) { p.A[l] = a2[l & 1]; p.B[l] = b2[l & 1]; p.PQ[l] = pq; p.S[l] = s; p.Msg[l] = msg; }
  }
  if (amax) p.Zrow = take(E * 2 * D * es);
  p.pooled = (float*)take(G * D * 4);
  const int64_t R = c.n_readout > 0 ? c.R : 0;
  p.H[0] = p.pooled;
  for (int k = 1; k <= c.n_readout; ++k) p.H[k] = (float*)take(G * R * 4);
  if (save) {
    p.dZA[0] = take(N * D * es); p.dZA[1] = take(N * D * es);
    p.dS = take(NS * 2 * D * es);
    if (amax) {
      p.dMsg = take(E * D * es);
      p.dZrow = take(E * 2 * D * es);
      p.einpos = (int32_t*)take(E * 4);
    }
    p.dBn[0] = take(M * D * es); p.dBn[1] = take(M * D * es);
    p.dPQ2[0] = take(M * 2 * D * es); p.dPQ2[1] = take(M * 2 * D * es);
    p.dPQ = p.dPQ2[0];
    p.mx = take(M * D * es);
    p.gsplit = take(M * D * es);
    for (int l = 1; l <= L; ++l) { p.tiemask[l] = (uint8_t*)take(E * 32); p.tiecnt[l] = (uint8_t*)take(M * D); }
    const int64_t wmax = D > R ? D : R;
    p.dH[0] = (float*)take(G * wmax * 4); p.dH[1] = (float*)take(G * wmax * 4);
    p.dPool = (float*)take(G * D * 4);
    // partial sums of the TN GEMMs / embedding gradients
    int64_t pf = (int64_t)tc::kTnMaxCtas * (2 * D * D + 2 * D);
    for (int k = 0; k < pl.n_lin; ++k) {
      const int64_t need = (int64_t)tn_splits((int)G) * pl.r_out[k] * (pl.r_in[k] + 1);
      if (need > pf) pf = need;
    }
    if (pl.n_lin > 0) {
      const int64_t need = ceil_div(G, readout_gb((int)G)) * ((pl.total - pl.r_w[0] + 3) & ~int64_t(3));  // 16-byte aligned per-CTA blocks
      if (need > pf) pf = need;
    }
    const int64_t vmax = c.atom_vocab > c.bond_vocab ? c.atom_vocab : c.bond_vocab;
    if (kEmbedSplits * vmax * D > pf) pf = kEmbedSplits * vmax * D;
    if ((int64_t)kEmbedSplits * c.bond_vocab * 2 * D > pf) pf = (int64_t)kEmbedSplits * c.bond_vocab * 2 * D;
    p.partial = (float*)take(pf * 4);
    p.partialAt = (float*)take((int64_t)tc::kTnMaxCtas * (2 * D * D + 2 * D) * 4);
    p.partialPq = (float*)take((int64_t)tc::kTnMaxCtas * (2 * D * D + 2 * D) * 4);
    // weight accumulator immediately followed by its bias accumulator (one reduction covers both)
    p.accWpq = (float*)take((2 * D * D + 2 * D) * 4);
    p.accBpq = p.accWpq + 2 * D * D;
    p.accWat = (float*)take((2 * D * D + D) * 4);
    p.accBat = p.accWat + 2 * D * D;
    p.tokPQ = (float*)take((int64_t)c.bond_vocab * 2 * D * 4);
  }
  p.total = o;
  return GCB_OK;
}

// The fused readout kernels win in the latency regime (a few CTAs instead of ~13 launches); for large
// batches the tiled GEMM path is faster, so the fused path is used up to 1024 graphs.
static bool make_readout_desc(const ParamLayout& pl, ReadoutDesc& rd, int G) {
  static const int fused_max = std::getenv("GCB_READOUT_FUSED_MAX") ? std::atoi(std::getenv("GCB_READOUT_FUSED_MAX")) : 1024;
  if (pl.n_lin <= 0 || pl.n_lin > kReadoutMaxLin || G > fused_max) return false;
  rd.n_lin = pl.n_lin;
  rd.max_dim = 0;
  for (int k = 0; k < pl.n_lin; ++k) {
    rd.in_dim[k] = pl.r_in[k]; rd.out_dim[k] = pl.r_out[k];
    rd.w_off[k] = pl.r_w[k]; rd.b_off[k] = pl.r_b[k];
    rd.p_off[k] = pl.r_w[k] - pl.r_w[0];
    rd.max_dim = rd.max_dim > pl.r_in[k] ? rd.max_dim : pl.r_in[k];
    rd.max_dim = rd.max_dim > pl.r_out[k] ? rd.max_dim : pl.r_out[k];
  }
  rd.partial_stride = pl.total - pl.r_w[0];
  rd.max_w = 0;
  for (int k = 0; k < pl.n_lin; ++k) {
    const int w = (pl.r_in[k] + 1) * (pl.r_out[k] + 1);
    rd.max_w = rd.max_w > w ? rd.max_w : w;
  }
  // shared memory: 3 activation panels of 32 x (max_dim + 1) floats + one weight matrix
  return (3 * kReadoutGB * (rd.max_dim + 1) + rd.max_w) * 4 <= 200 * 1024;
}

// Graphs per CTA of the register-tiled readout: 64.  32 (two CTAs per SM at 8192 graphs, twice the resident warps) measured
// SLOWER on B200 -- forward 31 vs 26 us, backward 66 vs 48 us, train step 1.506 vs 1.482 ms (profiles/r2_ab_x_readout.jsonl):
// every CTA stages the layer's weights again, and that, not the per-graph chain, is what the kernel waits for.
// GCB_READOUT_GB=32 keeps the variant reachable for tests / other shapes.
static int readout2_gb(int G) {
  const char* e = std::getenv("GCB_READOUT_GB");  // read per call: the tests switch it inside one process
  const int forced = e ? std::atoi(e) : 0;
  if (forced == 32 || forced == 64) return forced;
  (void)G;
  return 64;
}
// Register-tiled readout (readout.cuh, second half): large batches, every layer input a multiple of 4 wide.
static bool readout2_ok(const ParamLayout& pl, ReadoutDesc& rd, int G) {
  const char* e = std::getenv("GCB_NO_READOUT2");
  const bool off = e && e[0] == '1';
  static const int min_g = std::getenv("GCB_READOUT2_MIN_G") ? std::atoi(std::getenv("GCB_READOUT2_MIN_G")) : 1024;
  if (off || pl.n_lin <= 0 || pl.n_lin > kReadoutMaxLin || G <= min_g) return false;
  rd.n_lin = pl.n_lin;
  rd.max_dim = 0;
  for (int k = 0; k < pl.n_lin; ++k) {
    if (pl.r_in[k] % 4) return false;
    rd.in_dim[k] = pl.r_in[k]; rd.out_dim[k] = pl.r_out[k];
    rd.w_off[k] = pl.r_w[k]; rd.b_off[k] = pl.r_b[k];
    rd.p_off[k] = pl.r_w[k] - pl.r_w[0];
    rd.max_dim = rd.max_dim > pl.r_in[k] ? rd.max_dim : pl.r_in[k];
    rd.max_dim = rd.max_dim > pl.r_out[k] ? rd.max_dim : pl.r_out[k];
  }
  rd.max_dim = (rd.max_dim + 3) & ~3;
  rd.partial_stride = (pl.total - pl.r_w[0] + 3) & ~int64_t(3);  // 16-byte aligned per-CTA partial blocks
  rd.max_w = 0;
  return readout2_bwd_smem(rd) <= 200 * 1024;
}

struct Graph {
  int N, E, G, M;
  const int32_t *x_tok, *e_tok, *src, *dst, *in_ptr, *in_eid, *in_src, *out_ptr, *out_eid, *out_dst, *graph_ptr,
      *graph_node, *atok_ptr, *atok_node, *btok_ptr, *btok_row, *node_graph, *in_src4, *in_eid4, *out_dst4, *out_inpos4, *out_inpos,
      *tile_ptr, *bond_pos;
  int n_tiles;  // closed row tiles (0: none, use the unfused kernels)
  int pack_flags;
};
static int make_graph(const gcb_model_cfg& c, const int32_t* packed, const int32_t* h, Graph& g) {
  if (!packed || !h || h[GCB_H_MAGIC] != GCB_PACK_MAGIC) return GCB_ERR_INVALID_ARG;
  g.N = h[GCB_H_N]; g.E = h[GCB_H_E]; g.G = h[GCB_H_G]; g.M = h[GCB_H_M];
  if (h[GCB_H_AV] != c.atom_vocab || h[GCB_H_BV] != c.bond_vocab) return GCB_ERR_INVALID_ARG;
  const int expectM = c.L >= 1 ? (g.N < g.E ? g.N : g.E) : g.E;
  if (g.M != expectM) return GCB_ERR_INVALID_ARG;
  g.x_tok = packed + h[GCB_H_X_TOK]; g.e_tok = packed + h[GCB_H_E_TOK];
  g.src = packed + h[GCB_H_SRC]; g.dst = packed + h[GCB_H_DST];
  g.in_ptr = packed + h[GCB_H_IN_PTR]; g.in_eid = packed + h[GCB_H_IN_EID]; g.in_src = packed + h[GCB_H_IN_SRC];
  g.out_ptr = packed + h[GCB_H_OUT_PTR]; g.out_eid = packed + h[GCB_H_OUT_EID]; g.out_dst = packed + h[GCB_H_OUT_DST];
  g.graph_ptr = packed + h[GCB_H_GRAPH_PTR]; g.graph_node = packed + h[GCB_H_GRAPH_NODE];
  g.atok_ptr = packed + h[GCB_H_ATOK_PTR]; g.atok_node = packed + h[GCB_H_ATOK_NODE];
  g.btok_ptr = packed + h[GCB_H_BTOK_PTR]; g.btok_row = packed + h[GCB_H_BTOK_ROW];
  g.node_graph = packed + h[GCB_H_NODE_GRAPH];
  g.pack_flags = h[GCB_H_FLAGS];
  g.bond_pos = (g.pack_flags & GCB_PACK_EACH) && h[GCB_H_BOND_POS] > 0 ? packed + h[GCB_H_BOND_POS] : nullptr;
  if ((g.pack_flags & GCB_PACK_EACH) && !g.bond_pos) return GCB_ERR_INVALID_ARG;
  g.tile_ptr = packed + h[GCB_H_TILE_PTR];
  g.n_tiles = h[GCB_H_TILE_PTR] > 0 ? h[GCB_H_NTILES] : 0;
  g.in_src4 = packed + h[GCB_H_IN_SRC4]; g.in_eid4 = packed + h[GCB_H_IN_EID4]; g.out_dst4 = packed + h[GCB_H_OUT_DST4]; g.out_inpos4 = packed + h[GCB_H_OUT_INPOS4]; g.out_inpos = packed + h[GCB_H_OUT_INPOS];
  return GCB_OK;
}

static inline Drop make_drop(const gcb_model_cfg& c, int flags, uint64_t seed, uint32_t tag) {
  Drop d;
  if ((flags & GCB_TRAINING) && c.p_dropout > 0.f) { d.p = c.p_dropout; d.seed = seed; d.tag = tag; }
  return d;
}
static inline uint32_t tag_bond(int l) { return 4u * (uint32_t)l; }
static inline uint32_t tag_atom(int l) { return 4u * (uint32_t)l + 1u; }
static inline uint32_t tag_readout(int k) { return 100000u + (uint32_t)k; }

// The fused GEMM + message kernels (fused_rows.cuh) take bf16, D = 128, add aggregation, closed row tiles, every
// atom row a live bond row, and no active dropout; anything else runs on the unfused kernels.  `which`: 0 bond
// forward, 1 atom forward, 2 atom backward (opt-in: GCB_FUSED=1 / GCB_FUSED_BOND / _ATOM / _ABWD, see tc_host.cu).
// The fused bond forward records tie bits only (no tie counts) and the fused atom backward emits dS[:, D:] as
// planes: both are understood by tile::bond_bwd_kernel only, hence the merged-bond-backward condition.
template <typename T>
static bool fused_ok(const gcb_model_cfg& c, const Graph& g, int flags, int which) {
  if constexpr (sizeof(T) != 2) return false;
  return tc::tc_enabled() && tc::tile_enabled() && tc::fused_enabled(which) && c.D == 128 && c.aggr == GCB_AGGR_ADD &&
         g.n_tiles > 0 && g.N > 0 && g.M == g.N && tc::bond_bwd_merged_enabled() &&
         !((flags & GCB_TRAINING) && c.p_dropout > 0.f);
}

// Tile-staged row kernels (tile_rows.cuh): same preconditions, without the tensor-core coupling.
template <typename T>
static bool tiles_ok(const gcb_model_cfg& c, const Graph& g, int flags) {
  if constexpr (sizeof(T) != 2) return false;
  return tc::tile_enabled() && tile::width_ok(c.D) && g.n_tiles > 0 && g.N > 0 && g.M == g.N &&
         !((flags & GCB_TRAINING) && c.p_dropout > 0.f);
}

// Message step 1 straight from the bond-token table (tile::bond_update_tok_kernel and, in the backward pass, the
// per-token sums of [dP|dQ]): the tile preconditions plus a table that fits shared memory.  GCB_NO_TOK1=1 switches
// it off (A/B switch; the bit-for-bit comparisons with the plain row kernels use it, because summing [dP|dQ] per
// token BEFORE the transform rounds differently from transforming every row).
static bool tok1_enabled() {
  const char* e = std::getenv("GCB_NO_TOK1");
  return !(e && e[0] == '1');
}
template <typename T>
static bool tok1_ok(const gcb_model_cfg& c, const Graph& g, int flags) {
  if constexpr (sizeof(T) != 2) return false;
  return tok1_enabled() && tc::tc_enabled() && c.L >= 1 && tiles_ok<T>(c, g, flags) && tile::bond_tok_ok(c.D, c.bond_vocab) &&
         !fused_ok<T>(c, g, flags, 0);
}

// Atom side of the same fusion (reference gcn.py:152-153 -> 172-173): A_0 is never materialised, step 1 gathers the
// neighbours' rows out of the pre-activated atom table, and the residual of its transform GEMM and act' of its backward
// read the table through the token.  bf16 tile path with add aggregation only; GCB_NO_ATOK1=1 switches it off.
template <typename T>
static bool atok1_ok(const gcb_model_cfg& c, const Graph& g, int flags) {
  if constexpr (sizeof(T) != 2) return false;
  const char* e = std::getenv("GCB_NO_ATOK1");
  const bool off = e && e[0] == '1';
  return !off && tok1_enabled() && tc::tc_enabled() && c.L >= 1 && c.aggr == GCB_AGGR_ADD && tiles_ok<T>(c, g, flags) &&
         tile::atom_tok_ok(c.D, c.atom_vocab) && !fused_ok<T>(c, g, flags, 1) && !fused_ok<T>(c, g, flags, 2);
}

// ---- forward ------------------------------------------------------------------------------------
template <typename T>
static int forward_impl(const gcb_model_cfg& c, const ParamLayout& pl, const Graph& g, const float* params,
                        const Derived& dv, const Plan& p, int flags, uint64_t seed, float* out, T* out_atom,
                        T* out_bond, cudaStream_t s) {
  constexpr int V = Vec<T>::N;
  const int D = c.D, L = c.L, N = g.N, E = g.E, G = g.G, M = g.M;
  const T* wpq = (const T*)(dv.base + dv.wpq);
  const T* wat = (const T*)(dv.base + dv.wat);
  const float* bpq = (const float*)(dv.base + dv.bpq);
  const float* bat = (const float*)(dv.base + dv.bat);

  const bool atok1 = atok1_ok<T>(c, g, flags);  // A_0 is then never materialised: step 1 reads the atom token table
  if (N > 0 && !atok1) {
    GCB_KERNEL("embed_rows_kernel", s, embed_rows_kernel<T><<<row_chunk_blocks(N, D, V), 256, 0, s>>>((const T*)(dv.base + dv.tab_atom), g.x_tok, (T*)p.A[0], N, D / V));
  }
  const bool tok1 = tok1_ok<T>(c, g, flags);  // B_0 is then never materialised: step 1 reads the token table
  if (M > 0 && !tok1) {
    GCB_KERNEL("embed_rows_kernel", s, embed_rows_kernel<T><<<row_chunk_blocks(M, D, V), 256, 0, s>>>((const T*)(dv.base + dv.tab_bond), g.e_tok, (T*)p.B[0], M, D / V));
  }
  // Boustrophedon traversal: every large kernel walks its rows in the direction opposite to the
  // previous one, so it starts on the rows its producer wrote LAST (still L2 resident).
  int rev = 0;
  const bool pingpong = tc::pingpong_enabled();
  auto flip = [&]() { if (pingpong) rev ^= 1; return rev; };
  for (int l = 1; l <= L; ++l) {
    const Drop db = make_drop(c, flags, seed, tag_bond(l));
    const Drop da = make_drop(c, flags, seed, tag_atom(l));
    bool bond_done = false;
    if constexpr (sizeof(T) == 2) {
      if (M > 0 && fused_ok<T>(c, g, flags, 0)) {
        const bool save = (flags & GCB_SAVE_FOR_BWD) != 0;
        fr::RowArgs ra;
        ra.tile_ptr = g.tile_ptr; ra.n_tiles = g.n_tiles; ra.rows = M; ra.act = c.act;
        ra.ptr = g.in_ptr; ra.nbr = g.in_src; ra.nbr4 = (const int4*)g.in_src4;
        ra.bias = bpq; ra.out = (bf16*)p.B[l];
        ra.tiemask = save ? reinterpret_cast<uint32_t*>(p.tiemask[l]) : nullptr;
        const bool exact = tc::fused_exact();
        int rc;
        if (save) rc = exact ? fr::launch_rows<fr::BOND_FWD, true, true>("fused_bond_fwd_kernel", (const bf16*)p.B[l - 1], (const bf16*)wpq, nullptr, ra, s)
                             : fr::launch_rows<fr::BOND_FWD, true, false>("fused_bond_fwd_kernel", (const bf16*)p.B[l - 1], (const bf16*)wpq, nullptr, ra, s);
        else rc = exact ? fr::launch_rows<fr::BOND_FWD, false, true>("fused_bond_fwd_kernel", (const bf16*)p.B[l - 1], (const bf16*)wpq, nullptr, ra, s)
                        : fr::launch_rows<fr::BOND_FWD, false, false>("fused_bond_fwd_kernel", (const bf16*)p.B[l - 1], (const bf16*)wpq, nullptr, ra, s);
        if (rc == GCB_OK) bond_done = true;
        else if (rc != GCB_ERR_UNSUPPORTED) return rc;
      }
    }
    if constexpr (sizeof(T) == 2) {
      if (M > 0 && !bond_done && l == 1 && tok1) {
        const bool save = (flags & GCB_SAVE_FOR_BWD) != 0;
        const tile::TileCommon cm{g.tile_ptr, M, g.in_ptr, g.in_src, (const int4*)g.in_src4, nullptr, nullptr, flip()};
        GCB_TRY(tile::launch_bond_update_tok(D, cm, g.n_tiles, (const bf16*)(dv.base + dv.tab_pq), g.e_tok, c.bond_vocab,
                                             (bf16*)p.B[l], save ? p.tiemask[l] : nullptr, save ? p.tiecnt[l] : nullptr, c.act, s));
        bond_done = true;
      }
    }
    if (M > 0 && !bond_done) {
      Epilogue ep;
      ep.bias = bpq; ep.rev = flip();
      GCB_TRY(gemm_nt_auto<T>((const T*)p.B[l - 1], D, wpq, D, false, (T*)p.PQ[l], 2 * D, M, 2 * D, D, ep, s));
      bool tiled = false;
      if constexpr (sizeof(T) == 2) {
        if (tiles_ok<T>(c, g, flags)) {
          const bool save = (flags & GCB_SAVE_FOR_BWD) != 0;
          const tile::TileCommon cm{g.tile_ptr, M, g.in_ptr, g.in_src, (const int4*)g.in_src4, nullptr, nullptr, flip()};
          GCB_TRY(tile::launch_bond_update(D, (const bf16*)p.PQ[l], cm, g.n_tiles, (bf16*)p.B[l], save ? p.tiemask[l] : nullptr,
                                           save ? p.tiecnt[l] : nullptr, c.act, s));
          tiled = true;
        }
      }
      if (!tiled)
      GCB_TRY(launch_bond_update<T>((const T*)p.PQ[l], g.in_ptr, g.in_src, g.in_eid, g.in_src4, g.in_eid4, (T*)p.B[l],
                                    (flags & GCB_SAVE_FOR_BWD) ? p.tiemask[l] : nullptr,
                                    (flags & GCB_SAVE_FOR_BWD) ? p.tiecnt[l] : nullptr, M, D, c.act, db, flip(), s));
    }
    bool atom_done = false;
    if constexpr (sizeof(T) == 2) {
      if (N > 0 && fused_ok<T>(c, g, flags, 1)) {
        const bool save = (flags & GCB_SAVE_FOR_BWD) != 0;
        fr::AtomArgs aa;
        aa.tile_ptr = g.tile_ptr; aa.n_tiles = g.n_tiles; aa.rows = N; aa.M = M; aa.act = c.act;
        aa.in_ptr = g.in_ptr; aa.in_src = g.in_src; aa.in_src4 = (const int4*)g.in_src4;
        aa.in_eid = g.in_eid; aa.in_eid4 = (const int4*)g.in_eid4;
        aa.Bnew = (const bf16*)p.B[l]; aa.bias = bat; aa.Aout = (bf16*)p.A[l]; aa.S = save ? (bf16*)p.S[l] : nullptr;
        const int rc = save ? fr::launch_atom_fwd<true>((const bf16*)p.A[l - 1], (const bf16*)wat, aa, s)
                            : fr::launch_atom_fwd<false>((const bf16*)p.A[l - 1], (const bf16*)wat, aa, s);
        if (rc == GCB_OK) atom_done = true;
        else if (rc != GCB_ERR_UNSUPPORTED) return rc;
      }
    }
    if (N > 0 && !atom_done && c.aggr == GCB_AGGR_MAX) {
      // transform-then-aggregate (aggr_max.cuh): per-edge rows -> messages -> segmented max + residual + act
      GCB_TRY(launch_edge_rows<T>((const T*)p.A[l - 1], (const T*)p.B[l], g.in_src, g.in_eid, (T*)p.Zrow, E, M, D, c.act, db, s));
      if (E > 0) {
        Epilogue ep;
        ep.bias = bat;
        GCB_TRY(gemm_nt_auto<T>((const T*)p.Zrow, 2 * D, wat, 2 * D, false, (T*)p.Msg[l], D, E, D, 2 * D, ep, s));
      }
      GCB_TRY(launch_atom_max_update<T>((const T*)p.Msg[l], g.in_ptr, (const T*)p.A[l - 1], (T*)p.A[l], N, D, c.act, s));
      if (da.p > 0.f) {
        GCB_KERNEL("dropout_inplace_kernel", s, dropout_inplace_kernel<T><<<row_chunk_blocks(N, D, V), 256, 0, s>>>((T*)p.A[l], N, D, da));
      }
      atom_done = true;
    }
    if (N > 0 && !atom_done) {
      bool a_tiled = false;
      if constexpr (sizeof(T) == 2) {
        if (c.aggr == GCB_AGGR_ADD && tiles_ok<T>(c, g, flags)) {
          const tile::TileCommon cm{g.tile_ptr, N, g.in_ptr, g.in_src, (const int4*)g.in_src4, (const int4*)g.in_eid4, g.in_eid, flip()};
          if (l == 1 && atok1)
            GCB_TRY(tile::launch_atom_gather_tok(D, cm, g.n_tiles, (const bf16*)(dv.base + dv.tab_atom), g.x_tok, c.atom_vocab, (const bf16*)p.B[l], M,
                                                 (bf16*)p.S[l], c.act, s));
          else
            GCB_TRY(tile::launch_atom_gather(D, (const bf16*)p.A[l - 1], (const bf16*)p.B[l], M, cm, g.n_tiles, (bf16*)p.S[l], c.act, s));
          a_tiled = true;
        }
      }
      if (!a_tiled)
      GCB_TRY(launch_atom_gather<T>((const T*)p.A[l - 1], (const T*)p.B[l], g.in_ptr, g.in_src, g.in_eid, g.in_src4, g.in_eid4, (T*)p.S[l], N, M, D,
                                    c.act, c.aggr, db, flip(), s));
      Epilogue ep;
      ep.bias = bat; ep.rowptr = g.in_ptr; ep.aggr = c.aggr;
      ep.res = p.A[l - 1]; ep.ld_res = D; ep.act = c.act; ep.rev = flip();
      if (l == 1 && atok1) { ep.res = dv.base + dv.tab_atom; ep.res_idx = g.x_tok; }  // residual A_0[i] = tab_atom[tok i]
      GCB_TRY(gemm_nt_auto<T>((const T*)p.S[l], 2 * D, wat, 2 * D, false, (T*)p.A[l], D, N, D, 2 * D, ep, s));
      if (da.p > 0.f) {
        GCB_KERNEL("dropout_inplace_kernel", s, dropout_inplace_kernel<T><<<row_chunk_blocks(N, D, V), 256, 0, s>>>((T*)p.A[l], N, D, da));
      }
    }
  }
  // pooling + readout (fp32)
  if (G > 0) {
    GCB_KERNEL("pool_kernel", s, pool_kernel<T><<<row_chunk_blocks(G, D, V), 256, 0, s>>>((const T*)p.A[L], g.graph_ptr, g.graph_node, p.pooled, G, D, c.pool == GCB_POOL_MEAN ? 1 : 0));
    if (pl.n_lin == 0) {
      if (out) GCB_CUDA(cudaMemcpyAsync(out, p.pooled, (size_t)G * D * 4, cudaMemcpyDeviceToDevice, s));
    } else if (ReadoutDesc rd2; readout2_ok(pl, rd2, G)) {
      ReadoutPtrs hp;
      for (int k = 0; k <= pl.n_lin; ++k) hp.h[k] = k < pl.n_lin ? p.H[k] : nullptr;
      static bool attr2 = false;
      if (!attr2) {
        GCB_CUDA(cudaFuncSetAttribute(readout2_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        GCB_CUDA(cudaFuncSetAttribute(readout2_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr2 = true;
      }
      Drop drb = make_drop(c, flags, seed, tag_readout(0));
      if (readout2_gb(G) == 32)
        GCB_KERNEL("readout_fwd_kernel", s, readout2_fwd_kernel<32><<<(unsigned)ceil_div(G, 32), kReadout2Threads, readout2_fwd_smem(rd2, 32), s>>>(rd2, params, hp, out, G, c.act, drb));
      else
        GCB_KERNEL("readout_fwd_kernel", s, readout2_fwd_kernel<64><<<(unsigned)ceil_div(G, 64), kReadout2Threads, readout2_fwd_smem(rd2, 64), s>>>(rd2, params, hp, out, G, c.act, drb));
    } else if (ReadoutDesc rd; make_readout_desc(pl, rd, G)) {
      ReadoutPtrs hp;
      for (int k = 0; k <= pl.n_lin; ++k) hp.h[k] = k < pl.n_lin ? p.H[k] : nullptr;
      const int gb = readout_gb(G);
      const int smem = (2 * gb * (rd.max_dim + 1) + rd.max_w) * (int)sizeof(float);
      static bool attr = false;
      if (!attr) { GCB_CUDA(cudaFuncSetAttribute(readout_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
      Drop drb = make_drop(c, flags, seed, tag_readout(0));
      GCB_KERNEL("readout_fwd_kernel", s, readout_fwd_kernel<<<(unsigned)ceil_div(G, gb), kReadoutThreads, smem, s>>>(rd, params, hp, out, G, c.act, drb, gb));
    } else {
      for (int k = 0; k < pl.n_lin; ++k) {
        const bool last = k == pl.n_lin - 1;
        Epilogue ep;
        ep.bias = params + pl.r_b[k];
        ep.act = last ? -1 : c.act;
        float* dst = last ? out : p.H[k + 1];
        if (!dst) break;
        GCB_TRY(launch_gemm_nt<float>(p.H[k], pl.r_in[k], params + pl.r_w[k], pl.r_in[k], false, dst, pl.r_out[k], G,
                                      pl.r_out[k], pl.r_in[k], ep, s));
        const Drop dr = make_drop(c, flags, seed, tag_readout(k + 1));
        if (!last && dr.p > 0.f) {
          const int64_t cnt = (int64_t)G * pl.r_out[k];
          GCB_KERNEL("dropout_scalar_kernel", s, dropout_scalar_kernel<<<(unsigned)ceil_div(cnt, 256), 256, 0, s>>>(dst, cnt, dr));
        }
      }
    }
  }
  if (out_atom && N > 0) GCB_CUDA(cudaMemcpyAsync(out_atom, p.A[L], (size_t)N * D * sizeof(T), cudaMemcpyDeviceToDevice, s));
  if (out_bond && E > 0 && g.bond_pos) {
    // "each molecule alone" layout: edge e's row lives at bond_pos[e] (or is the constant act(0) row)
    GCB_KERNEL("gather_bond_rows_kernel", s, gather_bond_rows_kernel<T><<<row_chunk_blocks(E, D, V), 256, 0, s>>>((const T*)p.B[L], g.bond_pos, out_bond, E, D, c.act));
  } else if (out_bond && E > 0) {
    if (M > 0) GCB_CUDA(cudaMemcpyAsync(out_bond, p.B[L], (size_t)M * D * sizeof(T), cudaMemcpyDeviceToDevice, s));
    if (E > M) {
      GCB_KERNEL("fill_const_rows_kernel", s, fill_const_rows_kernel<T><<<row_chunk_blocks(E - M, D, V), 256, 0, s>>>(out_bond, M, E - M, D, c.act,
                                                                             make_drop(c, flags, seed, tag_bond(L))));
    }
  }
  return GCB_OK;
}

// ---- backward -----------------------------------------------------------------------------------
template <typename T>
static int backward_impl(const gcb_model_cfg& c, const ParamLayout& pl, const Graph& g, const float* params,
                         const Derived& dv, const Plan& p, int flags, uint64_t seed, const float* d_out,
                         const T* d_out_atom, const T* d_out_bond, float* grad, int accumulate, cudaStream_t s) {
  constexpr int V = Vec<T>::N;
  const int D = c.D, L = c.L, N = g.N, G = g.G, M = g.M;
  const T* watT = (const T*)(dv.base + dv.watT);
  const T* wpqT = (const T*)(dv.base + dv.wpqT);

  // Weight-gradient GEMMs (and their partial reductions) feed nothing until the optimiser step, so
  // they run on a side stream next to the latency-bound row kernels of the same / the next layer:
  //   W_at(l) = dZA_l^T S_l        may start when layer l starts; dZA_l is overwritten by atom_bwd(l-1)
  //   W_pq(l) = [dP|dQ]_l^T B_l-1  may start after scatter(l);   dPQ[l&1] is overwritten by tie(l-2)
  const bool tok1 = tok1_ok<T>(c, g, flags);
  SideStream* ss = tc::overlap_enabled() ? side_stream() : nullptr;
  const bool overlap = ss != nullptr;
  cudaStream_t sw = overlap ? ss->stream : s;
  size_t evi = 0;
  cudaEvent_t done_at[kMaxLayers + 2], done_pq[kMaxLayers + 2];
  for (int i = 0; i < kMaxLayers + 2; ++i) done_at[i] = done_pq[i] = nullptr;
  auto fork = [&]() -> int {  // side stream continues from the caller stream's current point
    if (!overlap) return GCB_OK;
    cudaEvent_t e = ss->event(evi++);
    if (!e) return GCB_ERR_CUDA;
    GCB_CUDA(cudaEventRecord(e, s));
    GCB_CUDA(cudaStreamWaitEvent(sw, e, 0));
    return GCB_OK;
  };
  auto mark = [&](cudaEvent_t& slot) -> int {  // remember "side stream reached this point"
    if (!overlap) return GCB_OK;
    slot = ss->event(evi++);
    if (!slot) return GCB_ERR_CUDA;
    GCB_CUDA(cudaEventRecord(slot, sw));
    return GCB_OK;
  };
  auto join = [&](cudaEvent_t e) -> int {
    if (overlap && e) GCB_CUDA(cudaStreamWaitEvent(s, e, 0));
    return GCB_OK;
  };
  // ---- readout ----
  const float* dPool = nullptr;
  if (d_out && G > 0) {
    if (pl.n_lin == 0) {
      dPool = d_out;
    } else if (ReadoutDesc rd2; readout2_ok(pl, rd2, G)) {
      ReadoutPtrs hp;
      for (int k = 0; k <= pl.n_lin; ++k) hp.h[k] = k < pl.n_lin ? p.H[k] : nullptr;
      static bool attr2 = false;
      if (!attr2) {
        GCB_CUDA(cudaFuncSetAttribute(readout2_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        GCB_CUDA(cudaFuncSetAttribute(readout2_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr2 = true;
      }
      const int gb = readout2_gb(G);
      const int blocks = (int)ceil_div(G, gb);
      Drop drb = make_drop(c, flags, seed, tag_readout(0));
      if (gb == 32)
        GCB_KERNEL("readout_bwd_kernel", s, readout2_bwd_kernel<32><<<blocks, kReadout2Threads, readout2_bwd_smem(rd2, 32), s>>>(rd2, params, hp, d_out, p.partial, p.dPool, G, c.act, drb));
      else
        GCB_KERNEL("readout_bwd_kernel", s, readout2_bwd_kernel<64><<<blocks, kReadout2Threads, readout2_bwd_smem(rd2, 64), s>>>(rd2, params, hp, d_out, p.partial, p.dPool, G, c.act, drb));
      // the ordered sum of the per-CTA readout weight gradients feeds nothing before the optimiser: on the side stream
      // (in front of the weight-gradient GEMMs, which reuse p.partial after it) pool_bwd does not wait for it
      const bool red_side = overlap && L >= 1;
      if (red_side) GCB_TRY(fork());
      GCB_TRY(launch_reduce_partials(p.partial, blocks, pl.total - pl.r_w[0], grad + pl.r_w[0], accumulate, red_side ? sw : s, rd2.partial_stride));
      dPool = p.dPool;
    } else if (ReadoutDesc rd; make_readout_desc(pl, rd, G)) {
      ReadoutPtrs hp;
      for (int k = 0; k <= pl.n_lin; ++k) hp.h[k] = k < pl.n_lin ? p.H[k] : nullptr;
      const int gb = readout_gb(G);
      const int smem = (3 * gb * (rd.max_dim + 1) + rd.max_w) * (int)sizeof(float);
      static bool attr = false;
      if (!attr) { GCB_CUDA(cudaFuncSetAttribute(readout_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
      const int blocks = (int)ceil_div(G, gb);
      Drop drb = make_drop(c, flags, seed, tag_readout(0));
      // a single CTA's partial block IS the readout gradient (same layout as the flat parameters): no reduction launch
      const bool direct = blocks == 1 && !accumulate;
      GCB_KERNEL("readout_bwd_kernel", s, readout_bwd_kernel<<<blocks, kReadoutThreads, smem, s>>>(rd, params, hp, d_out, direct ? grad + pl.r_w[0] : p.partial, p.dPool, G, c.act, drb, gb));
      if (!direct) GCB_TRY(launch_reduce_partials(p.partial, blocks, rd.partial_stride, grad + pl.r_w[0], accumulate, s));
      dPool = p.dPool;
    } else {
      const float* dcur = d_out;
      for (int k = pl.n_lin - 1; k >= 0; --k) {
        GCB_TRY(launch_gemm_tn<float>(dcur, pl.r_out[k], p.H[k], pl.r_in[k], G, pl.r_out[k], pl.r_in[k], p.partial,
                                      grad + pl.r_w[k], grad + pl.r_b[k], nullptr, GCB_AGGR_ADD, accumulate, s));
        float* dnext = k > 0 ? p.dH[k & 1] : p.dPool;
        Epilogue ep;
        if (k > 0) {
          ep.actgrad_y = p.H[k]; ep.ld_y = pl.r_in[k]; ep.act = c.act;
          const Drop dr = make_drop(c, flags, seed, tag_readout(k));
          ep.p_drop = dr.p; ep.seed = dr.seed; ep.drop_tag = dr.tag;
        }
        GCB_TRY(launch_gemm_nt<float>(dcur, pl.r_out[k], params + pl.r_w[k], pl.r_in[k], true, dnext, pl.r_in[k], G,
                                      pl.r_in[k], pl.r_out[k], ep, s));
        dcur = dnext;
      }
      dPool = p.dPool;
    }
  } else if (pl.n_lin > 0 && !accumulate) {
    for (int k = 0; k < pl.n_lin; ++k) {
      GCB_CUDA(cudaMemsetAsync(grad + pl.r_w[k], 0, sizeof(float) * (size_t)(pl.off[8 + 2 * k + 2] - pl.off[8 + 2 * k]), s));
    }
  }

  // ---- pool backward -> dZA_L ----
  int cur = 0;
  if (N > 0) {
    if (G > 0) {
      GCB_TRY(launch_pool_bwd<T>(dPool, d_out_atom, (const T*)p.A[L], g.graph_ptr, g.graph_node, g.node_graph, (T*)p.dZA[cur],
                                 N, G, D, c.act, L >= 1 ? make_drop(c, flags, seed, tag_atom(L)) : Drop(), c.pool == GCB_POOL_MEAN ? 1 : 0, s));
    } else {
      GCB_CUDA(cudaMemsetAsync(p.dZA[cur], 0, (size_t)N * D * sizeof(T), s));
    }
  }
  // ---- message-passing layers, last to first ----
  const T* dBnext = nullptr;
  int bcur = 0;
  int rev = 0;
  const bool pingpong = tc::pingpong_enabled();
  auto flip = [&]() { if (pingpong) rev ^= 1; return rev; };
  int grid_at = 0, grid_pq = 0;  // > 0: partials deferred in p.partialAt / p.partialPq
  for (int l = L; l >= 1; --l) {
    const int acc_w = l != L;
    T* dPQ_l = (T*)p.dPQ2[l & 1];
    bool atom_done = false;
    if constexpr (sizeof(T) == 2) {
      if (N > 0 && M > 0 && fused_ok<T>(c, g, flags, 2)) {
        GCB_TRY(join(done_at[l + 1]));  // W_at(l+1) read the buffer this layer's dZA_{l-1} goes to
        fr::RowArgs ra;
        ra.tile_ptr = g.tile_ptr; ra.n_tiles = g.n_tiles; ra.rows = N; ra.act = c.act;
        ra.ptr = g.out_ptr; ra.nbr = g.out_dst; ra.nbr4 = (const int4*)g.out_dst4;
        ra.out = (bf16*)p.dZA[cur ^ 1]; ra.dS2p = (bf16*)p.dS;
        const int rc = fr::launch_rows<fr::ATOM_BWD, false, false>("fused_atom_bwd_kernel", (const bf16*)p.dZA[cur], (const bf16*)watT,
                                                                   (const bf16*)p.A[l - 1], ra, s);
        if (rc == GCB_OK) atom_done = true;
        else if (rc != GCB_ERR_UNSUPPORTED) return rc;
      }
    }
    const bool amax = c.aggr == GCB_AGGR_MAX;
    if (amax) {
      // aggr = "max": the gradient goes to the in-edges that attained the maximum (even tie split), then through the
      // per-edge linear maps; dZrow [E, 2D] takes the place of dS for the source-side kernels (aggr_max.cuh)
      GCB_TRY(join(done_at[l + 1]));  // W_at(l+1) read Zrow / dMsg / dZA[cur ^ 1]
      if (l == L) GCB_TRY(launch_edge_inpos<T>(g.in_eid, p.einpos, g.E, s));
      GCB_TRY(launch_edge_rows<T>((const T*)p.A[l - 1], (const T*)p.B[l], g.in_src, g.in_eid, (T*)p.Zrow, g.E, M, D, c.act,
                                  make_drop(c, flags, seed, tag_bond(l)), s));
      GCB_TRY(launch_atom_max_bwd<T>((const T*)p.Msg[l], g.in_ptr, (const T*)p.dZA[cur], (T*)p.dMsg, N, D, s));
      if (g.E > 0) {
        Epilogue ep;
        GCB_TRY(gemm_nt_auto<T>((const T*)p.dMsg, D, watT, D, false, (T*)p.dZrow, 2 * D, g.E, 2 * D, D, ep, s));
      }
      GCB_TRY(fork());
      GCB_TRY(gemm_tn_auto<T>((const T*)p.dMsg, D, (const T*)p.Zrow, 2 * D, g.E, D, 2 * D, p.partial, p.accWat,
                                p.accBat, nullptr, GCB_AGGR_ADD, acc_w, sw));
      GCB_TRY(mark(done_at[l]));
      if (N > 0) {
        GCB_TRY(launch_atom_bwd_gather<T>((const T*)p.dZA[cur], (const T*)p.dZrow, g.out_ptr, g.out_inpos, g.out_inpos4, g.in_ptr,
                                          (const T*)p.A[l - 1], (T*)p.dZA[cur ^ 1], N, D, c.act, GCB_AGGR_ADD,
                                          l - 1 >= 1 ? make_drop(c, flags, seed, tag_atom(l - 1)) : Drop(), flip(), s));
      }
    } else if (atom_done) {
      GCB_TRY(fork());
      GCB_TRY(gemm_tn_auto<T>((const T*)p.dZA[cur], D, (const T*)p.S[l], 2 * D, N, D, 2 * D, p.partial, p.accWat,
                                p.accBat, g.in_ptr, c.aggr, acc_w, sw, 0, p.partialAt, &grid_at, l == L));
      GCB_TRY(mark(done_at[l]));
    } else {
      if (N > 0) {
        Epilogue ep;
        ep.rev = flip();
        GCB_TRY(gemm_nt_auto<T>((const T*)p.dZA[cur], D, watT, D, false, (T*)p.dS, 2 * D, N, 2 * D, D, ep, s));
      }
      GCB_TRY(fork());
      GCB_TRY(gemm_tn_auto<T>((const T*)p.dZA[cur], D, (const T*)p.S[l], 2 * D, N, D, 2 * D, p.partial, p.accWat,
                                p.accBat, g.in_ptr, c.aggr, acc_w, sw, 0, p.partialAt, &grid_at, l == L));
      GCB_TRY(mark(done_at[l]));
      GCB_TRY(join(done_at[l + 1]));  // W_at(l+1) read the buffer atom_bwd(l) is about to overwrite
      bool tiled = false;
      if constexpr (sizeof(T) == 2) {
        if (N > 0 && c.aggr == GCB_AGGR_ADD && tiles_ok<T>(c, g, flags)) {
          const tile::TileCommon cm{g.tile_ptr, N, g.out_ptr, g.out_dst, (const int4*)g.out_dst4, nullptr, nullptr, flip()};
          const bool from_tab = l == 1 && atok1_ok<T>(c, g, flags);  // A_0 lives in the token table
          GCB_TRY(tile::launch_atom_bwd(D, (const bf16*)p.dS, (const bf16*)p.dZA[cur],
                                        from_tab ? (const bf16*)(dv.base + dv.tab_atom) : (const bf16*)p.A[l - 1], from_tab ? g.x_tok : nullptr, cm,
                                        g.n_tiles, (bf16*)p.dZA[cur ^ 1], c.act, s));
          tiled = true;
        }
      }
      if (N > 0 && !tiled) {
        GCB_TRY(launch_atom_bwd_gather<T>((const T*)p.dZA[cur], (const T*)p.dS, g.out_ptr, g.out_dst, g.out_dst4, g.in_ptr, (const T*)p.A[l - 1],
                                          (T*)p.dZA[cur ^ 1], N, D, c.act, c.aggr,
                                          l - 1 >= 1 ? make_drop(c, flags, seed, tag_atom(l - 1)) : Drop(), flip(), s));
      }
    }
    GCB_TRY(join(done_pq[l + 2]));  // W_pq(l+2) read dPQ[l&1]
    bool bond_tiled = false;
    if constexpr (sizeof(T) == 2) {
      if (M > 0 && c.aggr == GCB_AGGR_ADD && tiles_ok<T>(c, g, flags) && tc::bond_bwd_merged_enabled()) {
        const tile::TileCommon cm{g.tile_ptr, M, g.out_ptr, g.out_dst, (const int4*)g.out_dst4, (const int4*)g.out_inpos4, g.out_inpos, flip()};
        GCB_TRY(tile::launch_bond_bwd(D, cm, g.n_tiles, g.in_ptr, g.dst, (const bf16*)p.dS, (const bf16*)dBnext,
                                      l == L ? (const bf16*)d_out_bond : nullptr, (const bf16*)p.B[l],
                                      fused_ok<T>(c, g, flags, 0) ? nullptr : p.tiecnt[l], p.tiemask[l],
                                      (bf16*)dPQ_l, c.act, atom_done ? N : 0, s));
        bond_tiled = true;
      }
    }
    if (M > 0 && !bond_tiled) {
      // bond row r = edge r feeds the atom it points to: dS[dst r, D:] (add / mean), its own message row (max)
      GCB_TRY(launch_bond_bwd_tie<T>(amax ? (const T*)p.dZrow : (const T*)p.dS, dBnext, l == L ? d_out_bond : nullptr, (const T*)p.B[l], (const T*)p.PQ[l],
                                     amax ? p.einpos : g.dst, g.in_ptr, g.in_src, g.in_eid, g.in_src4, g.in_eid4, dPQ_l, (T*)p.mx, (T*)p.gsplit,
                                     p.tiemask[l], p.tiecnt[l], M, D, c.act, amax ? (int)GCB_AGGR_ADD : c.aggr,
                                     make_drop(c, flags, seed, tag_bond(l)), flip(), s));
      bool tiled = false;
      if constexpr (sizeof(T) == 2) {
        if (tiles_ok<T>(c, g, flags)) {
          const tile::TileCommon cm{g.tile_ptr, M, g.out_ptr, g.out_dst, (const int4*)g.out_dst4, (const int4*)g.out_inpos4, g.out_inpos, flip()};
          GCB_TRY(tile::launch_bond_scatter(D, (const bf16*)p.gsplit, cm, g.n_tiles, p.tiemask[l], (bf16*)dPQ_l, s));
          tiled = true;
        }
      }
      if (!tiled)
      GCB_TRY(launch_bond_bwd_scatter<T>((const T*)p.PQ[l], (const T*)p.mx, (const T*)p.gsplit, g.out_ptr, g.out_dst, g.out_inpos, g.out_dst4, g.out_inpos4,
                                         p.tiemask[l], dPQ_l, M, D, flip(), s));
    }
    if (l == 1 && tok1) {  // step 1 read the token table: its weight / embedding gradients go through per-token sums below
      cur ^= 1;
      continue;
    }
    GCB_TRY(fork());
    GCB_TRY(gemm_tn_auto<T>((const T*)dPQ_l, 2 * D, (const T*)p.B[l - 1], D, M, 2 * D, D, p.partial, p.accWpq,
                              p.accBpq, nullptr, GCB_AGGR_ADD, acc_w, sw, 0, p.partialPq, &grid_pq, l == L));
    GCB_TRY(mark(done_pq[l]));
    if (M > 0) {
      Epilogue ep;
      ep.rev = flip();
      GCB_TRY(gemm_nt_auto<T>((const T*)dPQ_l, 2 * D, wpqT, 2 * D, false, (T*)p.dBn[bcur], D, M, D, 2 * D, ep, s));
      dBnext = (const T*)p.dBn[bcur];
      bcur ^= 1;
    }
    cur ^= 1;
  }
  if (grid_at > 0) GCB_TRY(launch_reduce_partials(p.partialAt, grid_at, (int64_t)2 * D * D + D, p.accWat, 0, sw, (int64_t)2 * D * D + D));
  if (grid_pq > 0) GCB_TRY(launch_reduce_partials(p.partialPq, grid_pq, (int64_t)2 * D * D + 2 * D, p.accWpq, 0, sw, (int64_t)2 * D * D + 2 * D));
  // ---- embeddings ----
  // The two embedding gradients are independent, low-occupancy bucket walks: the bond one runs on the internal stream
  // behind the last weight-gradient GEMM (it needs the summed conv accumulators anyway), the atom one on the caller's
  // stream, each with its own partial-sum scratch; the streams meet again in front of finalize_conv_grads_kernel.
  {
    const int av = c.atom_vocab, bv = c.bond_vocab;
    const int64_t scratch_cap = (int64_t)tc::kTnMaxCtas * (2 * D * D + 2 * D);
    const bool split = overlap && L >= 1 && grid_at == 0 && grid_pq == 0 && (int64_t)kEmbedSplits * av * D <= scratch_cap &&
                       (int64_t)kEmbedSplits * bv * 2 * D <= scratch_cap;
    cudaStream_t se = split ? sw : s;             // stream of the bond half
    float* part_a = split ? p.partialAt : p.partial;
    float* part_b = split ? p.partialPq : p.partial;
    if (split) {  // the bond half reads what the caller's stream produced last ([dP|dQ]_1 or dB_0)
      cudaEvent_t e = ss->event(evi++);
      if (!e) return GCB_ERR_CUDA;
      GCB_CUDA(cudaEventRecord(e, s));
      GCB_CUDA(cudaStreamWaitEvent(sw, e, 0));
    } else if (overlap && L >= 1) {  // everything on the side stream must be done before its results are used
      cudaEvent_t last = nullptr;
      GCB_TRY(mark(last));
      GCB_TRY(join(last));
    }
    if (N > 0) {
      GCB_TRY(launch_embed_bwd_partial<T>((const T*)p.dZA[cur], nullptr, g.atok_ptr, g.atok_node, part_a, av, D, c.act, s));
    }
    GCB_TRY(launch_reduce_partials(part_a, N > 0 ? embed_splits() : 0, (int64_t)av * D, grad + pl.emb_atom, accumulate, s));
    if (tok1 && M > 0) {
      // per-token sums of step 1's [dP|dQ] (buckets in row order, ordered split reduction), then the three small
      // products of bond_table_bwd_kernel
      if constexpr (sizeof(T) == 2) {
        GCB_TRY(launch_embed_bwd_partial<T>((const T*)p.dPQ2[1], nullptr, g.btok_ptr, g.btok_row, part_b, bv, 2 * D, c.act, se));
        GCB_TRY(launch_reduce_partials(part_b, embed_splits(), (int64_t)bv * 2 * D, p.tokPQ, 0, se));
        const int nb1 = (int)ceil_div(2 * D * D, 256);
        GCB_KERNEL("bond_table_bwd_kernel", se, bond_table_bwd_kernel<T><<<(unsigned)(nb1 + bv * (D / 32)), 256, 0, se>>>(
            p.tokPQ, (const T*)(dv.base + dv.tab_bond), (const T*)(dv.base + dv.wpq), p.accWpq, p.accBpq, grad + pl.emb_bond, bv, D,
            c.act, L == 1, accumulate, nb1));
      }
    } else {
      const T* dB0 = L >= 1 ? dBnext : d_out_bond;
      const bool have = dB0 != nullptr && M > 0;
      if (have) {
        GCB_TRY(launch_embed_bwd_partial<T>(dB0, (const T*)p.B[0], g.btok_ptr, g.btok_row, part_b, bv, D, c.act, se));
      }
      GCB_TRY(launch_reduce_partials(part_b, have ? embed_splits() : 0, (int64_t)bv * D, grad + pl.emb_bond, accumulate, se));
    }
    if (split) {
      cudaEvent_t last = nullptr;
      GCB_TRY(mark(last));
      GCB_TRY(join(last));
    }
  }
  // ---- conv parameters ----
  GCB_KERNEL("finalize_conv_grads_kernel", s, finalize_conv_grads_kernel<<<(unsigned)ceil_div(2 * D * D, 256), 256, 0, s>>>(
      p.accWpq, p.accBpq, p.accWat, p.accBat, grad + pl.w_msg, grad + pl.b_msg, grad + pl.w_edge, grad + pl.b_edge,
      grad + pl.w_bond, grad + pl.b_bond, D, accumulate, L == 0));
  return GCB_OK;
}

}  // namespace gcb

using namespace gcb;

// =================================================================================================
extern "C" int gcb_abi_version(void) { return GCB_ABI_VERSION; }
extern "C" const char* gcb_last_cuda_error(void) { return g_err; }
extern "C" const char* gcb_build_info(void) {
  return "graphchem_b200 sm_100a; kernels: tcgen05/TMEM/TMA GEMMs (bf16), TMA-staged tile row kernels, CSR row kernels + CUDA-core GEMMs (fp32), device-side batch packer; built " __DATE__ " " __TIME__;
}

extern "C" int gcb_param_layout(const gcb_model_cfg* cfg, int64_t* offsets) {
  if (!cfg) return GCB_ERR_INVALID_ARG;
  ParamLayout pl;
  int rc = make_layout(*cfg, pl);
  if (rc != GCB_OK) return rc;
  if (offsets) std::memcpy(offsets, pl.off, sizeof(int64_t) * (pl.n_tensors + 1));
  return pl.n_tensors;
}

extern "C" int64_t gcb_derived_bytes(const gcb_model_cfg* cfg) {
  if (!cfg || check_cfg(*cfg) != GCB_OK) return GCB_ERR_INVALID_ARG;
  return derived_layout(*cfg, nullptr).total;
}

extern "C" int gcb_prepare_weights(const gcb_model_cfg* cfg, const float* params, void* derived, void* stream) {
  if (!cfg || !params || !derived) return GCB_ERR_INVALID_ARG;
  GCB_TRY(check_cfg(*cfg));
  ParamLayout pl;
  GCB_TRY(make_layout(*cfg, pl));
  Derived d = derived_layout(*cfg, derived);
  cudaStream_t s = (cudaStream_t)stream;
  const int D = cfg->D;
  const int n_atom = cfg->atom_vocab * D, n_bond = cfg->bond_vocab * D;
  int most = 2 * D * D;
  most = most > n_atom ? most : n_atom;
  most = most > n_bond ? most : n_bond;
  const unsigned blocks = (unsigned)ceil_div(most, 256);
  if (cfg->dtype == GCB_BF16) {
    GCB_KERNEL("prepare_weights_kernel", s, prepare_weights_kernel<bf16><<<blocks, 256, 0, s>>>(
        params + pl.w_msg, params + pl.b_msg, params + pl.w_edge, params + pl.b_edge, params + pl.w_bond,
        params + pl.b_bond, (bf16*)(d.base + d.wpq), (bf16*)(d.base + d.wat), (bf16*)(d.base + d.watT),
        (bf16*)(d.base + d.wpqT), (float*)(d.base + d.bpq), (float*)(d.base + d.bat), D, params + pl.emb_atom,
        params + pl.emb_bond, (bf16*)(d.base + d.tab_atom), (bf16*)(d.base + d.tab_bond), n_atom, n_bond, cfg->act));
    if (cfg->L >= 1 && tok1_enabled() && tc::tc_enabled() && tile::bond_tok_ok(D, cfg->bond_vocab)) {
      // [P|Q] of every bond token through the SAME tcgen05 GEMM that transforms bond rows (bit-identical rows);
      // prepare_weights_kernel is a normal launch, so it is complete before this GEMM's weight loads start
      Epilogue ep;
      ep.bias = (const float*)(d.base + d.bpq);
      GCB_TRY(gemm_nt_auto<bf16>((const bf16*)(d.base + d.tab_bond), D, (const bf16*)(d.base + d.wpq), D, false,
                                 (bf16*)(d.base + d.tab_pq), 2 * D, cfg->bond_vocab, 2 * D, D, ep, s));
    }
  } else {
    GCB_KERNEL("prepare_weights_kernel", s, prepare_weights_kernel<float><<<blocks, 256, 0, s>>>(
        params + pl.w_msg, params + pl.b_msg, params + pl.w_edge, params + pl.b_edge, params + pl.w_bond,
        params + pl.b_bond, (float*)(d.base + d.wpq), (float*)(d.base + d.wat), (float*)(d.base + d.watT),
        (float*)(d.base + d.wpqT), (float*)(d.base + d.bpq), (float*)(d.base + d.bat), D, params + pl.emb_atom,
        params + pl.emb_bond, (float*)(d.base + d.tab_atom), (float*)(d.base + d.tab_bond), n_atom, n_bond, cfg->act));
  }
  return GCB_OK;
}

extern "C" int64_t gcb_workspace_bytes(const gcb_model_cfg* cfg, int32_t N, int32_t E, int32_t G, int32_t flags) {
  if (!cfg || N < 0 || E < 0 || G < 0) return GCB_ERR_INVALID_ARG;
  if (check_cfg(*cfg) != GCB_OK) return GCB_ERR_UNSUPPORTED;
  ParamLayout pl;
  if (make_layout(*cfg, pl) != GCB_OK) return GCB_ERR_INVALID_ARG;
  const int64_t M = cfg->L >= 1 ? (N < E ? N : E) : E;
  Plan p;
  if (make_plan(*cfg, pl, N, E, G, M, flags, nullptr, p) != GCB_OK) return GCB_ERR_UNSUPPORTED;
  return p.total;
}

extern "C" int gcb_forward(const gcb_model_cfg* cfg, const int32_t* packed, const int32_t* packed_header_host,
                           const float* params, const void* derived, void* workspace, int64_t workspace_bytes,
                           int32_t flags, uint64_t seed, float* out, void* out_atom, void* out_bond, void* stream) {
  if (!cfg || !params || !derived || !workspace) return GCB_ERR_INVALID_ARG;
  GCB_TRY(check_cfg(*cfg));
  ParamLayout pl;
  GCB_TRY(make_layout(*cfg, pl));
  Graph g;
  GCB_TRY(make_graph(*cfg, packed, packed_header_host, g));
  // GCB_PACK_EACH batches are for prediction: no saved activations, no dropout
  if ((g.pack_flags & GCB_PACK_EACH) && ((flags & GCB_SAVE_FOR_BWD) || ((flags & GCB_TRAINING) && cfg->p_dropout > 0.f)))
    return GCB_ERR_INVALID_ARG;
  Plan p;
  GCB_TRY(make_plan(*cfg, pl, g.N, g.E, g.G, g.M, flags, workspace, p));
  if (p.total > workspace_bytes) return GCB_ERR_WORKSPACE;
  Derived dv = derived_layout(*cfg, const_cast<void*>(derived));
  cudaStream_t s = (cudaStream_t)stream;
  if (cfg->dtype == GCB_BF16)
    return forward_impl<bf16>(*cfg, pl, g, params, dv, p, flags, seed, out, (bf16*)out_atom, (bf16*)out_bond, s);
  return forward_impl<float>(*cfg, pl, g, params, dv, p, flags, seed, out, (float*)out_atom, (float*)out_bond, s);
}

extern "C" int gcb_backward(const gcb_model_cfg* cfg, const int32_t* packed, const int32_t* packed_header_host,
                            const float* params, const void* derived, void* workspace, int64_t workspace_bytes,
                            int32_t flags, uint64_t seed, const float* d_out, const void* d_out_atom,
                            const void* d_out_bond, float* grad, int32_t accumulate, void* stream) {
  if (!cfg || !params || !derived || !workspace || !grad) return GCB_ERR_INVALID_ARG;
  if (!(flags & GCB_SAVE_FOR_BWD)) return GCB_ERR_INVALID_ARG;
  GCB_TRY(check_cfg(*cfg));
  ParamLayout pl;
  GCB_TRY(make_layout(*cfg, pl));
  Graph g;
  GCB_TRY(make_graph(*cfg, packed, packed_header_host, g));
  Plan p;
  GCB_TRY(make_plan(*cfg, pl, g.N, g.E, g.G, g.M, flags, workspace, p));
  if (p.total > workspace_bytes) return GCB_ERR_WORKSPACE;
  Derived dv = derived_layout(*cfg, const_cast<void*>(derived));
  cudaStream_t s = (cudaStream_t)stream;
  if (cfg->dtype == GCB_BF16)
    return backward_impl<bf16>(*cfg, pl, g, params, dv, p, flags, seed, d_out, (const bf16*)d_out_atom,
                               (const bf16*)d_out_bond, grad, accumulate, s);
  return backward_impl<float>(*cfg, pl, g, params, dv, p, flags, seed, d_out, (const float*)d_out_atom,
                              (const float*)d_out_bond, grad, accumulate, s);
}

extern "C" int gcb_mse_loss(const float* pred, const float* y, int64_t n, float* loss, float* d_pred,
                            float grad_scale, void* stream) {
  if (!pred || !y || n <= 0) return GCB_ERR_INVALID_ARG;
  GCB_KERNEL("mse_kernel", (cudaStream_t)stream, mse_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pred, y, n, loss, d_pred, grad_scale));
  return GCB_OK;
}

extern "C" int gcb_adam_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                             const float* hyper, int32_t* step_count, float beta1, float beta2, float eps,
                             void* stream) {
  if (!params || !grad || !exp_avg || !exp_avg_sq || !hyper || !step_count || n < 0) return GCB_ERR_INVALID_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (n > 0) {
    GCB_KERNEL("adam_kernel", s, adam_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(params, grad, exp_avg, exp_avg_sq, n, hyper, step_count,
                                                           beta1, beta2, eps));
  }
  GCB_KERNEL("increment_kernel", s, increment_kernel<<<1, 1, 0, s>>>(step_count));
  return GCB_OK;
}

// The notebook training loop over pre-packed device batches in ONE call (examples/predict_cn.ipynb:243-252):
//   for batch in loader: pred = model(batch); loss = mse_loss(pred, batch.y); loss.backward(); optimizer.step()
// Per batch exactly what gcb_forward -> gcb_mse_loss -> gcb_backward -> gcb_adam_step -> gcb_prepare_weights enqueue,
// in that order on `stream` (bit-identical to calling them one by one): the small-batch regime is bound by the host
// work per step, and this removes the caller's per-step share of it.
extern "C" int gcb_train_steps(const gcb_model_cfg* cfg, int32_t n_batches, const int32_t* const* packed,
                               const int32_t* const* packed_header_host, const float* const* y, float* params,
                               void* derived, void* workspace, int64_t workspace_bytes, uint64_t seed0, float* pred,
                               float* d_pred, float* grad, float* exp_avg, float* exp_avg_sq, const float* hyper,
                               int32_t* step_count, float beta1, float beta2, float eps, float* losses, void* stream) {
  if (!cfg || n_batches < 0 || (n_batches > 0 && (!packed || !packed_header_host || !y)) || !params || !derived ||
      !workspace || !pred || !d_pred || !grad || !losses)
    return GCB_ERR_INVALID_ARG;
  GCB_TRY(check_cfg(*cfg));
  ParamLayout pl;
  GCB_TRY(make_layout(*cfg, pl));
  const int32_t flags = GCB_TRAINING | GCB_SAVE_FOR_BWD;
  const int64_t cols = cfg->n_readout > 0 ? cfg->out_dim : cfg->D;
  for (int32_t b = 0; b < n_batches; ++b) {
    const int32_t* h = packed_header_host[b];
    if (!packed[b] || !h || !y[b] || h[GCB_H_MAGIC] != GCB_PACK_MAGIC) return GCB_ERR_INVALID_ARG;
    const uint64_t seed = seed0 + (uint64_t)b;
    GCB_TRY(gcb_forward(cfg, packed[b], h, params, derived, workspace, workspace_bytes, flags, seed, pred, nullptr, nullptr, stream));
    GCB_TRY(gcb_mse_loss(pred, y[b], (int64_t)h[GCB_H_G] * cols, losses + b, d_pred, 1.0f, stream));
    GCB_TRY(gcb_backward(cfg, packed[b], h, params, derived, workspace, workspace_bytes, flags, seed, d_pred, nullptr, nullptr,
                         grad, 0, stream));
    GCB_TRY(gcb_adam_step(params, grad, exp_avg, exp_avg_sq, pl.total, hyper, step_count, beta1, beta2, eps, stream));
    GCB_TRY(gcb_prepare_weights(cfg, params, derived, stream));
  }
  return GCB_OK;
}

// ---- debug entry: one bf16 GEMM through the dispatcher (tests/test_gemm_gpu.py) -------------------
extern "C" int gcb_debug_gemm_nt_bf16(const void* A, const void* W, void* C, int32_t rows, int32_t N, int32_t K,
                                      const float* bias, const int32_t* rowptr, int32_t aggr, const void* res,
                                      int32_t act, void* stream) {
  Epilogue ep;
  ep.bias = bias; ep.rowptr = rowptr; ep.aggr = aggr; ep.res = res; ep.ld_res = N; ep.act = act;
  return gemm_nt_auto<bf16>((const bf16*)A, K, (const bf16*)W, K, false, (bf16*)C, N, rows, N, K, ep, (cudaStream_t)stream);
}

extern "C" int gcb_debug_gemm_tn_bf16(const void* A, const void* B, int32_t rows, int32_t Nn, int32_t Kk,
                                      float* partial, float* dW, float* db, const int32_t* rowptr, int32_t aggr,
                                      void* stream) {
  return gemm_tn_auto<bf16>((const bf16*)A, Nn, (const bf16*)B, Kk, rows, Nn, Kk, partial, dW, db, rowptr, aggr, 0,
                            (cudaStream_t)stream);
}

// ---- timing / launch-count facility --------------------------------------------------------------
extern "C" int64_t gcb_launch_count(void) { return g_launches; }
extern "C" int gcb_timing_enable(int32_t enable) {
  g_timing = enable != 0;
  return GCB_OK;
}
extern "C" int64_t gcb_timing_report(char* buf, int64_t buflen) {
  // Synchronises the recorded events, aggregates by kernel name and clears the table.
  std::map<std::string, std::pair<int64_t, double>> agg;
  if (g_slots) {
    for (auto& t : *g_slots) {
      float ms = 0.f;
      if (cudaEventSynchronize(t.e1) == cudaSuccess && cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) {
        auto& a = agg[t.name];
        a.first += 1;
        a.second += ms;
      }
      cudaEventDestroy(t.e0);
      cudaEventDestroy(t.e1);
    }
    g_slots->clear();
  }
  std::string out = "[";
  bool first = true;
  for (auto& kv : agg) {
    char line[256];
    snprintf(line, sizeof(line), "%s{\"name\":\"%s\",\"launches\":%lld,\"ms\":%.6f}", first ? "" : ",", kv.first.c_str(),
             (long long)kv.second.first, kv.second.second);
    out += line;
    first = false;
  }
  out += "]";
  if (buf && buflen > 0) {
    const int64_t n = (int64_t)out.size() < buflen - 1 ? (int64_t)out.size() : buflen - 1;
    std::memcpy(buf, out.data(), (size_t)n);
    buf[n] = 0;
  }
  return (int64_t)out.size();
}
