/*
 * graphchem_b200 -- C-ABI of the B200-native MoleculeGCN hot path.
 *
 * The reference (ecrl/graphchem 2.3.3) is pure Python on PyTorch-Geometric and has no FFI of its
 * own; the boundary it exposes to user code is the Python class API.  Every entry
 * point below names the reference interface (file:line under /root/reference) whose work it
 * replaces.  The Python host (graphchem_b200/_lib.py) binds these with ctypes; INTEGRATION.md
 * shows the stub a graphchem maintainer would add.
 *
 * Conventions: plain pointers and sizes only (no torch types); every function returns GCB_OK or a
 * negative error code and never throws; all device buffers are caller-allocated, 16-byte
 * aligned; no hidden device allocation, no hidden host synchronisation; all kernels are deterministic
 * (no floating-point atomics).  `stream` is a cudaStream_t passed as void* (0 = legacy default
 * stream).  State the library keeps (per host thread, created lazily): one internal non-blocking
 * side stream + events on which gcb_backward overlaps its weight-gradient GEMMs (it forks from and
 * rejoins `stream`, so a CUDA-graph capture of `stream` captures it too), the opt-in launch counter /
 * CUDA-event timing facility (gcb_timing_*), and per-kernel "shared-memory attribute set" flags.
 * Behaviour switches (GCB_* environment variables) are read-only A/B knobs for profiling.
 */
#ifndef GRAPHCHEM_B200_H
#define GRAPHCHEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCB_ABI_VERSION 3

/* ---- error codes ------------------------------------------------------------------------- */
#define GCB_OK 0
#define GCB_ERR_INVALID_ARG (-1) /* reference raises ValueError (PyG _check_input, bad shapes)   */
#define GCB_ERR_INDEX (-2)       /* reference raises IndexError (edge_index/token out of range) */
#define GCB_ERR_UNSUPPORTED (-3) /* configuration outside what the kernels implement            */
#define GCB_ERR_CUDA (-4)        /* a CUDA runtime call failed; see gcb_last_cuda_error()       */
#define GCB_ERR_WORKSPACE (-5)   /* caller-provided buffer too small                            */

/* ---- enums ------------------------------------------------------------------------------- */
/* act_fn of graphchem/nn/gcn.py:42,75 -- the five callables tests/test_gcn.py:107-113 exercise */
enum gcb_act { GCB_ACT_SOFTPLUS = 0, GCB_ACT_SIGMOID = 1, GCB_ACT_RELU = 2, GCB_ACT_TANH = 3, GCB_ACT_LEAKY_RELU = 4 };
/* aggr of GeneralConv, graphchem/nn/gcn.py:41,84-86 */
enum gcb_aggr { GCB_AGGR_ADD = 0, GCB_AGGR_MEAN = 1, GCB_AGGR_MAX = 2 };
/* graph pooling, graphchem/nn/gcn.py:181 (the reference calls global_add_pool; mean = PyG global_mean_pool:
   segment sum / max(count, 1), BASELINE north_star item 4) */
enum gcb_pool { GCB_POOL_ADD = 0, GCB_POOL_MEAN = 1 };
/* storage type of activations (accumulation is always fp32) */
enum gcb_dtype { GCB_F32 = 0, GCB_BF16 = 1 };

/* Constructor arguments of MoleculeGCN, graphchem/nn/gcn.py:35-42. */
typedef struct gcb_model_cfg {
  int32_t atom_vocab;  /* atom_vocab_size                                  */
  int32_t bond_vocab;  /* bond_vocab_size                                  */
  int32_t out_dim;     /* output_dim (ignored when n_readout == 0)          */
  int32_t D;           /* embedding_dim                                    */
  int32_t L;           /* n_messages                                       */
  int32_t n_readout;   /* n_readout                                        */
  int32_t R;           /* readout_dim                                      */
  int32_t act;         /* enum gcb_act                                     */
  int32_t aggr;        /* enum gcb_aggr                                    */
  int32_t dtype;       /* enum gcb_dtype                                   */
  float p_dropout;     /* p_dropout (applied only when flags & GCB_TRAINING) */
  int32_t pool;        /* enum gcb_pool: graph pooling of gcn.py:181 (0 = the reference's global_add_pool) */
} gcb_model_cfg;

/* ---- packed batch ------------------------------------------------------------------------ */
/*
 * A PyG-style Batch (x, edge_index, edge_attr, batch; examples/predict_cn.ipynb:212,245 and
 * graphchem/data/structs.py:61-66) pre-packed into ONE int32 buffer: tokens, the edge list, a
 * destination-sorted and a source-sorted CSR (stable, i.e. edge-id order inside a row, which is
 * the reference's scatter order), graph segments and token buckets.  GCB_PACK_HEADER_INTS header
 * ints are followed by the sections; offsets (in ints, from the buffer start) live in the header.
 */
#define GCB_PACK_MAGIC 0x47434232 /* "GCB2" */
#define GCB_PACK_HEADER_INTS 48
#define GCB_TILE_ROWS 128
enum gcb_pack_field {
  GCB_H_MAGIC = 0, GCB_H_N, GCB_H_E, GCB_H_G, GCB_H_M, GCB_H_AV, GCB_H_BV, GCB_H_TOTAL_INTS,
  GCB_H_X_TOK, GCB_H_E_TOK, GCB_H_SRC, GCB_H_DST,
  GCB_H_IN_PTR, GCB_H_IN_EID, GCB_H_IN_SRC,
  GCB_H_OUT_PTR, GCB_H_OUT_EID, GCB_H_OUT_DST,
  GCB_H_GRAPH_PTR, GCB_H_GRAPH_NODE,
  GCB_H_ATOK_PTR, GCB_H_ATOK_NODE, GCB_H_BTOK_PTR, GCB_H_BTOK_ROW,
  GCB_H_NODE_GRAPH, /* [N] graph id of every node (the PyG `batch` vector as int32) */
  /* ELL-4 views of the CSRs: first four entries of every row, -1 padded, one int4 per row */
  GCB_H_IN_SRC4, GCB_H_IN_EID4, GCB_H_OUT_DST4, GCB_H_OUT_INPOS4,
  GCB_H_OUT_INPOS, /* [E] for every src-CSR entry: position of the same edge in the dst-CSR */
  /* Closed row tiles for the fused GEMM + message kernels: tile t covers rows [tile_ptr[t],
   * tile_ptr[t+1]), at most GCB_TILE_ROWS of them, and no edge leaves a tile (molecules are never
   * split).  GCB_H_NTILES = 0 when some connected component spans more than GCB_TILE_ROWS rows
   * (the unfused kernels are used then).  Written by the packer, data dependent. */
  GCB_H_TILE_PTR, GCB_H_NTILES,
  GCB_H_FLAGS,    /* GCB_PACK_* bits */
  GCB_H_BOND_POS, /* GCB_PACK_EACH only: [E] bond-matrix row that holds edge e, or -1 (constant row) */
  /* The four ELL-4 views are the LAST sections of the buffer and a pure function of the CSRs in front of them:
   * ints [0, header[GCB_H_COMPACT_INTS]) are enough to move over PCIe, gcb_pack_expand_ell() derives the rest on
   * the device (37 % fewer H2D bytes at the benchmark shape). */
  GCB_H_COMPACT_INTS,
  GCB_H_FIELDS
};

/* GCB_PACK_EACH ("each molecule alone"): the batch is laid out so that every molecule sees the bond
 * rows it would see if it were the only graph in the call.  The reference indexes the bond matrix by
 * ATOM id (gcn.py:163), so a batched forward differs from per-molecule forwards; the
 * notebooks predict one molecule per call (examples/predict_cn.ipynb:258-263, 333-334).  With this
 * flag bond row r holds the molecule's OWN edge number (r - first atom of the molecule), so one
 * batched launch reproduces the per-molecule results.  Inference only. */
#define GCB_PACK_EACH 1

/* Number of int32 elements a packed batch needs; fills header[GCB_PACK_HEADER_INTS] (host memory)
 * with sizes and section offsets.  M = number of live bond rows = min(N,E) when n_messages >= 1,
 * E when n_messages == 0. */
int64_t gcb_pack_ints(int32_t N, int32_t E, int32_t G, int32_t M, int32_t atom_vocab, int32_t bond_vocab,
                      int32_t* header);
/* Same with GCB_PACK_* flags (GCB_PACK_EACH adds the [E] bond-position section). */
int64_t gcb_pack_ints_flags(int32_t N, int32_t E, int32_t G, int32_t M, int32_t atom_vocab, int32_t bond_vocab,
                            int32_t flags, int32_t* header);

/* Host packer (replaces nothing in the reference: it is the "prepacked into pinned CSR buffers"
 * step of the north star; the index semantics restate PyG's scatter order).
 * x, edge_attr: int32 or int64 ([N], [E]); edge_index: int64 or int32 [2,E] row-major;
 * batch: int64/int32 [N] or NULL (one graph, gcn.py:181 batch=None).  idx_bytes = 4 or 8 for all.
 * check_bond_rows != 0 enforces max(edge_index) < E (the EdgeConv-on-bond-matrix precondition,
 * gcn.py:163).  Returns GCB_ERR_INDEX as the reference would raise IndexError. */
int gcb_pack_host(const void* x, const void* edge_attr, const void* edge_index, const void* batch,
                  int32_t idx_bytes, int32_t N, int32_t E, int32_t G, int32_t M, int32_t atom_vocab,
                  int32_t bond_vocab, int32_t check_bond_rows, int32_t* packed, int64_t packed_ints);

/* Host collate + pack (replaces torch_geometric.loader.DataLoader's Collater ->
 * Batch.from_data_list, examples/predict_cn.ipynb:212,245): concatenates n_graphs
 * molecules (per-graph pointer arrays, MoleculeEncoder.encode layout of features.py:263-321:
 * int32 atoms, int32 bonds, int64 [2,e] connectivity) with cumulative node offsets and packs the
 * result.  y_out (float [n_graphs*n_targets]) receives the concatenated targets. */
int gcb_collate_pack_host(const int32_t* const* atoms, const int32_t* n_atoms, const int32_t* const* bonds,
                          const int64_t* const* conn, const int32_t* n_edges, const float* const* y,
                          int32_t n_targets, int32_t n_graphs, int32_t n_messages, int32_t atom_vocab,
                          int32_t bond_vocab, int32_t* packed, int64_t packed_ints, float* y_out);

/* Collate + pack in GCB_PACK_EACH layout (batched prediction that equals one forward per molecule:
 * the `for g in dataset: model(g)` loops of examples/predict_cn.ipynb:258-263, 333-334 in one call).
 * Same arguments as gcb_collate_pack_host; size the buffer with gcb_pack_ints_flags(GCB_PACK_EACH).
 * Needs n_messages >= 1 and sum(n_edges) >= sum(n_atoms); returns GCB_ERR_INDEX when a molecule's
 * connectivity addresses a bond row it does not have (the reference raises IndexError for it alone)
 * and GCB_ERR_UNSUPPORTED when the batch has fewer edges than atoms. */
int gcb_collate_pack_each_host(const int32_t* const* atoms, const int32_t* n_atoms, const int32_t* const* bonds,
                               const int64_t* const* conn, const int32_t* n_edges, const float* const* y,
                               int32_t n_targets, int32_t n_graphs, int32_t n_messages, int32_t atom_vocab,
                               int32_t bond_vocab, int32_t* packed, int64_t packed_ints, float* y_out);

/* Derives the ELL-4 views (GCB_H_IN_SRC4 .. GCB_H_OUT_INPOS4) of a packed batch ON THE DEVICE from its CSR sections:
 * the host copies only the first header[GCB_H_COMPACT_INTS] ints of a host-packed batch into a full-size device
 * buffer and calls this on the copy's stream.  Bit-identical to the views the host packers write
 * (tests/test_pack_device.py).  Replaces nothing in the reference (PyG's Collater ships index tensors as they are,
 * examples/predict_cn.ipynb:245); it exists for the end-to-end path's PCIe budget. */
int gcb_pack_expand_ell(int32_t* packed_dev, const int32_t* packed_header_host, void* stream);

/* ---- device-side packer (SURVEY.md 8 d config 5 "packed on device"; 8 f1) ------------------- */
/* The host gathers a batch into a RAW buffer (0.7 kB per molecule instead of 3.9 kB): header, molecule
 * atom / edge offsets [G+1], atom tokens [N], bond tokens [E], src / dst [E] with batch-global atom ids,
 * closed-tile table; the device derives every other section.  The packed batch is bit-identical to
 * gcb_collate_pack_host on the same molecules (same collate semantics: Collater -> Batch.from_data_list,
 * examples/predict_cn.ipynb:212,245; same scatter order). */
#define GCB_RAW_MAGIC 0x47435231 /* "GCR1" */
#define GCB_RAW_HEADER_INTS 16
/* int32 elements of the raw buffer of a batch with N atoms, E directed edges, G molecules. */
int64_t gcb_raw_ints(int32_t N, int32_t E, int32_t G);
/* Host half: validate (same GCB_ERR_INDEX cases as gcb_collate_pack_host), gather into raw (host, normally
 * pinned), concatenate targets into y_out, and fill packed_header[GCB_PACK_HEADER_INTS] (host) with the
 * layout of the packed batch including its tile count.  GCB_ERR_UNSUPPORTED when an edge leaves its
 * molecule (use gcb_collate_pack_host for such input). */
int gcb_collate_raw_host(const int32_t* const* atoms, const int32_t* n_atoms, const int32_t* const* bonds,
                         const int64_t* const* conn, const int32_t* n_edges, const float* const* y,
                         int32_t n_targets, int32_t n_graphs, int32_t n_messages, int32_t atom_vocab,
                         int32_t bond_vocab, int32_t* raw, int64_t raw_ints, float* y_out, int32_t* packed_header);
/* int32 elements of device scratch gcb_pack_device needs for the batch described by packed_header. */
int64_t gcb_pack_device_scratch_ints(const int32_t* packed_header);
/* Device half: raw_dev (device copy of raw) -> packed_dev (device, packed_header[GCB_H_TOTAL_INTS] ints).
 * raw_header_host = the first GCB_RAW_HEADER_INTS ints of raw, packed_header_host from gcb_collate_raw_host
 * (both host memory).  Enqueues its kernels on `stream`; no allocation, no synchronisation. */
int gcb_pack_device(const int32_t* raw_dev, const int32_t* raw_header_host, const int32_t* packed_header_host,
                    int32_t* packed_dev, int64_t packed_ints, int32_t* scratch_dev, int64_t scratch_ints, void* stream);
/* Test hook (like gcb_debug_gemm_*): the per-molecule / per-chunk routines of gcb_pack_device's kernels run
 * serially on the CPU over HOST buffers, so the index logic is pinned without a GPU.  force_scratch = 1 takes
 * the large-molecule path (scratch instead of thread-local arrays) for every molecule, 2 the phases of the
 * experimental warp-per-molecule kernel (GCB_PACK_WARP=1), lane by lane.  Not a product path. */
int gcb_debug_pack_emulate_host(const int32_t* raw, const int32_t* packed_header, int32_t* packed, int64_t packed_ints,
                                int32_t* scratch, int64_t scratch_ints, int32_t force_scratch);

/* ---- parameters -------------------------------------------------------------------------- */
/* Flat fp32 parameter vector in the reference's state_dict order (gcn.py:78-114):
 * emb_atom.weight, emb_bond.weight, atom_conv.lin_msg.{weight,bias}, atom_conv.lin_edge.{weight,
 * bias}, bond_conv.nn.0.{weight,bias}, readout.k.0.{weight,bias} (k = 0..n_readout).
 * offsets[] receives n_tensors+1 element offsets; returns n_tensors (8 + 2*(n_readout+1) or 8). */
#define GCB_MAX_PARAM_TENSORS 64
int gcb_param_layout(const gcb_model_cfg* cfg, int64_t* offsets);

/* Derived (re-laid-out, dtype-converted) copies of the conv weights the kernels consume.  Must be
 * refreshed after every parameter update. */
int64_t gcb_derived_bytes(const gcb_model_cfg* cfg);
int gcb_prepare_weights(const gcb_model_cfg* cfg, const float* params, void* derived, void* stream);

/* ---- forward / backward ------------------------------------------------------------------ */
#define GCB_TRAINING 1    /* nn.Module.training: dropout active (gcn.py:167-169,176-178,194-196) */
#define GCB_SAVE_FOR_BWD 2 /* keep per-layer activations in the workspace for gcb_backward        */

int64_t gcb_workspace_bytes(const gcb_model_cfg* cfg, int32_t N, int32_t E, int32_t G, int32_t flags);

/* MoleculeGCN.forward, graphchem/nn/gcn.py:120-202 (embedding :152-157, EdgeConv :163-164,
 * GeneralConv :172-173, dropout :167-178, global_add_pool :181, readout :184-199).
 * packed: device packed batch; packed_header_host: host copy of its first GCB_PACK_HEADER_INTS ints,
 * taken AFTER packing (the packer writes the closed-tile count there; a header straight from
 * gcb_pack_ints selects the untiled kernels).  out: float [G, out_dim] (or [G, D] when n_readout == 0);
 * out_atom: dtype [N, D]; out_bond: dtype [E, D] (either may be NULL to skip materialising). */
int gcb_forward(const gcb_model_cfg* cfg, const int32_t* packed, const int32_t* packed_header_host,
                const float* params, const void* derived, void* workspace, int64_t workspace_bytes,
                int32_t flags, uint64_t seed, float* out, void* out_atom, void* out_bond, void* stream);

/* loss.backward() through the forward above (examples/predict_cn.ipynb:250).  Needs the workspace
 * of a forward run with GCB_SAVE_FOR_BWD and the same flags/seed.  d_out: float [G, out_dim];
 * d_out_atom (dtype [N,D]) and d_out_bond (dtype [E,D]) may be NULL.  grad: flat fp32 in the
 * gcb_param_layout order, overwritten (accumulate == 0) or accumulated into (accumulate != 0). */
int gcb_backward(const gcb_model_cfg* cfg, const int32_t* packed, const int32_t* packed_header_host,
                 const float* params, const void* derived, void* workspace, int64_t workspace_bytes,
                 int32_t flags, uint64_t seed, const float* d_out, const void* d_out_atom,
                 const void* d_out_bond, float* grad, int32_t accumulate, void* stream);

/* F.mse_loss(pred, y) mean-reduced and its gradient (examples/predict_cn.ipynb:249): loss[0] =
 * mean((pred-y)^2); d_pred = 2 (pred-y) / n * grad_scale. */
int gcb_mse_loss(const float* pred, const float* y, int64_t n, float* loss, float* d_pred, float grad_scale,
                 void* stream);

/* torch.optim.Adam.step (examples/predict_cn.ipynb:230-232,251): betas, eps, no weight decay,
 * bias-corrected, one fused pass over the flat buffers.  hyper (device, float[2]) = {lr, grad_scale}
 * so that the per-epoch lr rewrite (predict_cn.ipynb:241-242) survives CUDA-graph replay;
 * step_count (device int32[1]) is incremented by the kernel. */
int gcb_adam_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  const float* hyper, int32_t* step_count, float beta1, float beta2, float eps,
                  void* stream);

/* The notebook training loop over n_batches pre-packed device batches in one call (examples/predict_cn.ipynb:243-252:
 * for batch in loader: model(batch) -> F.mse_loss -> backward -> optimizer.step): per batch b exactly what gcb_forward,
 * gcb_mse_loss (loss into losses[b]), gcb_backward, gcb_adam_step and gcb_prepare_weights enqueue on `stream`, in that
 * order, with dropout seed seed0 + b.  packed / packed_header_host / y: host arrays of n_batches pointers (device packed
 * batch, its host header, device targets [G, out]); pred / d_pred: float scratch for the largest batch's [G, out];
 * workspace sized for the largest batch with GCB_TRAINING | GCB_SAVE_FOR_BWD; losses: device float[n_batches]. */
int gcb_train_steps(const gcb_model_cfg* cfg, int32_t n_batches, const int32_t* const* packed,
                    const int32_t* const* packed_header_host, const float* const* y, float* params, void* derived,
                    void* workspace, int64_t workspace_bytes, uint64_t seed0, float* pred, float* d_pred, float* grad,
                    float* exp_avg, float* exp_avg_sq, const float* hyper, int32_t* step_count, float beta1, float beta2,
                    float eps, float* losses, void* stream);

/* ---- bench / debug facility (thread-local) ------------------------------------------------- */
/* Number of kernels this thread has launched through the library since load. */
int64_t gcb_launch_count(void);
/* enable != 0: bracket every kernel launch with CUDA events on its stream (not capturable). */
int gcb_timing_enable(int32_t enable);
/* Synchronise the recorded events, write a JSON array [{"name","launches","ms"}] aggregated by
 * kernel name into buf (NUL-terminated, truncated to buflen), clear the table; returns the length. */
int64_t gcb_timing_report(char* buf, int64_t buflen);

/* Debug: C[rows,N] = epilogue(A[rows,K] W[N,K]^T) in bf16 through the same dispatcher the model
 * uses (tcgen05 kernel when the shape is supported and GCB_DISABLE_TC != 1). */
int gcb_debug_gemm_nt_bf16(const void* A, const void* W, void* C, int32_t rows, int32_t N, int32_t K,
                           const float* bias, const int32_t* rowptr, int32_t aggr, const void* res, int32_t act,
                           void* stream);

/* Debug: dW[Nn,Kk] = A[rows,Nn]^T B[rows,Kk], db[Nn] = sum_r w(r) A[r,:] (bf16 in, fp32 out);
 * partial: >= 148*(Nn*Kk+Nn) floats of scratch. */
int gcb_debug_gemm_tn_bf16(const void* A, const void* B, int32_t rows, int32_t Nn, int32_t Kk, float* partial,
                           float* dW, float* db, const int32_t* rowptr, int32_t aggr, void* stream);

/* ---- misc -------------------------------------------------------------------------------- */
int gcb_abi_version(void);
const char* gcb_last_cuda_error(void);
const char* gcb_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHCHEM_B200_H */
