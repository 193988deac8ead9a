// Host-side packer: PyG-style Batch -> one int32 buffer with tokens, edge list, destination- and
// source-sorted CSR, graph segments and token buckets.  Pure integer work; must match
// oracle/csr_oracle.py (stable argsort) bit for bit.
//
// Reference semantics restated here (not code): PyG scatters messages in edge-id order over
// edge_index[1], collate concatenates graphs with cumulative node offsets
// (call site examples/predict_cn.ipynb:212,245), MoleculeEncoder.encode emits
// int32 tokens and an int64 [2,e] connectivity (graphchem/preprocessing/features.py:294-319).
#include <cstring>
#include <vector>

#include "common.cuh"

namespace gcb {

static inline int64_t pad4(int64_t n) { return (n + 3) & ~int64_t(3); }

// Section table of a packed batch: (header field, element count), in buffer order.
struct Sec { int field; int64_t n; };
static int section_table(int32_t N, int32_t E, int32_t G, int32_t av, int32_t bv, int32_t flags, Sec* t) {
  int k = 0;
  auto put = [&](int field, int64_t n) { t[k].field = field; t[k].n = n; ++k; };
  put(GCB_H_X_TOK, N); put(GCB_H_E_TOK, E); put(GCB_H_SRC, E); put(GCB_H_DST, E);
  put(GCB_H_IN_PTR, (int64_t)N + 1); put(GCB_H_IN_EID, E); put(GCB_H_IN_SRC, E);
  put(GCB_H_OUT_PTR, (int64_t)N + 1); put(GCB_H_OUT_EID, E); put(GCB_H_OUT_DST, E);
  put(GCB_H_GRAPH_PTR, (int64_t)G + 1); put(GCB_H_GRAPH_NODE, N);
  put(GCB_H_ATOK_PTR, (int64_t)av + 1); put(GCB_H_ATOK_NODE, N);
  put(GCB_H_BTOK_PTR, (int64_t)bv + 1); put(GCB_H_BTOK_ROW, E);
  put(GCB_H_NODE_GRAPH, N);
  put(GCB_H_IN_SRC4, 4 * (int64_t)N); put(GCB_H_IN_EID4, 4 * (int64_t)N); put(GCB_H_OUT_DST4, 4 * (int64_t)N);
  put(GCB_H_OUT_INPOS4, 4 * (int64_t)N);
  put(GCB_H_OUT_INPOS, E);
  put(GCB_H_TILE_PTR, (int64_t)N + 1);
  if (flags & GCB_PACK_EACH) put(GCB_H_BOND_POS, E);
  return k;
}

static int64_t layout(int32_t N, int32_t E, int32_t G, int32_t M, int32_t av, int32_t bv, int32_t flags, int32_t* h) {
  int64_t off = GCB_PACK_HEADER_INTS;
  if (h) {
    std::memset(h, 0, sizeof(int32_t) * GCB_PACK_HEADER_INTS);
    h[GCB_H_MAGIC] = GCB_PACK_MAGIC;
    h[GCB_H_N] = N; h[GCB_H_E] = E; h[GCB_H_G] = G; h[GCB_H_M] = M; h[GCB_H_AV] = av; h[GCB_H_BV] = bv;
    h[GCB_H_FLAGS] = flags;
  }
  Sec t[GCB_H_FIELDS];
  const int k = section_table(N, E, G, av, bv, flags, t);
  for (int i = 0; i < k; ++i) {
    if (h) h[t[i].field] = (int32_t)off;
    off += pad4(t[i].n);
  }
  if (off > INT32_MAX) return -1;
  if (h) h[GCB_H_TOTAL_INTS] = (int32_t)off;
  return off;
}

// Every byte of a packed batch is defined (two packs of the same input compare equal) without clearing
// the whole buffer first: only the alignment pads and the tails of the sections the packer fills
// partially are zeroed.
static void zero_unwritten(int32_t* packed, const int32_t* h, int64_t e_tok_written) {
  const int32_t N = h[GCB_H_N], E = h[GCB_H_E], G = h[GCB_H_G], M = h[GCB_H_M];
  Sec t[GCB_H_FIELDS];
  const int k = section_table(N, E, G, h[GCB_H_AV], h[GCB_H_BV], h[GCB_H_FLAGS], t);
  for (int i = 0; i < k; ++i)
    for (int64_t j = t[i].n; j < pad4(t[i].n); ++j) packed[h[t[i].field] + j] = 0;
  auto tail = [&](int field, int64_t written, int64_t n) {
    if (written < n) std::memset(packed + h[field] + written, 0, sizeof(int32_t) * (size_t)(n - written));
  };
  tail(GCB_H_BTOK_ROW, M, E);
  tail(GCB_H_TILE_PTR, (int64_t)packed[GCB_H_NTILES] + 1, (int64_t)N + 1);
  tail(GCB_H_E_TOK, e_tok_written, E);
}

// Stable counting sort of ids 0..n-1 by key: ptr[nb+1], perm[n].
static void bucket(const int32_t* key, int64_t n, int32_t nb, int32_t* ptr, int32_t* perm) {
  std::memset(ptr, 0, sizeof(int32_t) * ((size_t)nb + 1));
  for (int64_t i = 0; i < n; ++i) ptr[key[i] + 1]++;
  for (int32_t b = 0; b < nb; ++b) ptr[b + 1] += ptr[b];
  std::vector<int32_t> cur(ptr, ptr + nb);
  for (int64_t i = 0; i < n; ++i) perm[cur[key[i]]++] = (int32_t)i;
}

static inline int64_t read_idx(const void* p, int32_t idx_bytes, int64_t i) {
  return idx_bytes == 8 ? ((const int64_t*)p)[i] : (int64_t)((const int32_t*)p)[i];
}

// Fill every derived section from x_tok/e_tok/src/dst (already written) and a batch vector.
static int finish_pack(int32_t* packed, const int32_t* h, const int32_t* batch_i32 /*[N] or null*/) {
  const int32_t N = h[GCB_H_N], E = h[GCB_H_E], G = h[GCB_H_G], M = h[GCB_H_M];
  const int32_t av = h[GCB_H_AV], bv = h[GCB_H_BV];
  int32_t* x_tok = packed + h[GCB_H_X_TOK];
  int32_t* e_tok = packed + h[GCB_H_E_TOK];
  int32_t* src = packed + h[GCB_H_SRC];
  int32_t* dst = packed + h[GCB_H_DST];
  int32_t* in_ptr = packed + h[GCB_H_IN_PTR];
  int32_t* in_eid = packed + h[GCB_H_IN_EID];
  int32_t* in_src = packed + h[GCB_H_IN_SRC];
  int32_t* out_ptr = packed + h[GCB_H_OUT_PTR];
  int32_t* out_eid = packed + h[GCB_H_OUT_EID];
  int32_t* out_dst = packed + h[GCB_H_OUT_DST];
  bucket(dst, E, N, in_ptr, in_eid);
  for (int64_t k = 0; k < E; ++k) in_src[k] = src[in_eid[k]];
  bucket(src, E, N, out_ptr, out_eid);
  for (int64_t k = 0; k < E; ++k) out_dst[k] = dst[out_eid[k]];
  {  // ELL-4 views (first four entries of each CSR row, -1 padded)
    int32_t* s4 = packed + h[GCB_H_IN_SRC4];
    int32_t* e4 = packed + h[GCB_H_IN_EID4];
    int32_t* d4 = packed + h[GCB_H_OUT_DST4];
    int32_t* o4 = packed + h[GCB_H_OUT_INPOS4];
    int32_t* out_inpos = packed + h[GCB_H_OUT_INPOS];
    {
      std::vector<int32_t> pos_of_eid((size_t)E);
      for (int64_t k = 0; k < E; ++k) pos_of_eid[(size_t)in_eid[k]] = (int32_t)k;
      for (int64_t k = 0; k < E; ++k) out_inpos[k] = pos_of_eid[(size_t)out_eid[k]];
    }
    for (int64_t i = 0; i < N; ++i) {
      for (int u = 0; u < 4; ++u) {
        const int64_t ki = (int64_t)in_ptr[i] + u, ko = (int64_t)out_ptr[i] + u;
        s4[4 * i + u] = ki < in_ptr[i + 1] ? in_src[ki] : -1;
        e4[4 * i + u] = ki < in_ptr[i + 1] ? in_eid[ki] : -1;
        d4[4 * i + u] = ko < out_ptr[i + 1] ? out_dst[ko] : -1;
        o4[4 * i + u] = ko < out_ptr[i + 1] ? out_inpos[ko] : -1;
      }
    }
  }
  int32_t* graph_ptr = packed + h[GCB_H_GRAPH_PTR];
  int32_t* graph_node = packed + h[GCB_H_GRAPH_NODE];
  int32_t* node_graph = packed + h[GCB_H_NODE_GRAPH];
  if (batch_i32) {
    bucket(batch_i32, N, G, graph_ptr, graph_node);
    std::memcpy(node_graph, batch_i32, sizeof(int32_t) * (size_t)N);
  } else {
    graph_ptr[0] = 0;
    for (int32_t g = 1; g <= G; ++g) graph_ptr[g] = N;
    for (int32_t i = 0; i < N; ++i) graph_node[i] = i;
    std::memset(node_graph, 0, sizeof(int32_t) * (size_t)N);
  }
  bucket(x_tok, N, av, packed + h[GCB_H_ATOK_PTR], packed + h[GCB_H_ATOK_NODE]);
  bucket(e_tok, M, bv, packed + h[GCB_H_BTOK_PTR], packed + h[GCB_H_BTOK_ROW]);
  {  // closed row tiles: cut positions are the rows no edge spans; greedy fill up to GCB_TILE_ROWS
    int32_t* tile_ptr = packed + h[GCB_H_TILE_PTR];
    std::vector<int32_t> span((size_t)N + 2, 0);  // span[p] > 0 after the prefix sum: some edge has lo < p <= hi
    for (int64_t e = 0; e < E; ++e) {
      const int32_t lo = src[e] < dst[e] ? src[e] : dst[e], hi = src[e] < dst[e] ? dst[e] : src[e];
      if (lo < hi) { span[(size_t)lo + 1]++; span[(size_t)hi + 1]--; }
    }
    int32_t nt = 0, start = 0, last_cut = 0, cover = 0;
    bool ok = true;
    tile_ptr[0] = 0;
    for (int32_t p = 1; p <= N && ok; ++p) {
      cover += span[(size_t)p];
      const bool cut = p == N || cover == 0;
      if (!cut) continue;
      if (p - start > GCB_TILE_ROWS) {  // close the tile at the previous cut
        if (last_cut == start) { ok = false; break; }
        tile_ptr[++nt] = last_cut;
        start = last_cut;
        if (p - start > GCB_TILE_ROWS) { ok = false; break; }
      }
      last_cut = p;
    }
    if (ok && N > 0) tile_ptr[++nt] = N;
    packed[GCB_H_NTILES] = ok ? nt : 0;
  }
  return GCB_OK;
}

// Collated molecules: every edge stays inside its molecule, atoms and edges of a molecule are contiguous, and
// graph ids increase -- so all index structures except the token buckets are per-molecule problems whose
// working set (tens of atoms / edges) lives in L1.  One sequential pass over the molecules instead of a dozen
// passes over the whole batch; the result is identical to finish_pack() (stable order inside every row).
// src / dst / tokens are already written; n_atoms / n_edges describe the molecules in batch order.
static int finish_pack_local(int32_t* packed, const int32_t* h, const int32_t* n_atoms, const int32_t* n_edges) {
  const int32_t N = h[GCB_H_N], E = h[GCB_H_E], G = h[GCB_H_G], M = h[GCB_H_M];
  const int32_t av = h[GCB_H_AV], bv = h[GCB_H_BV];
  const int32_t* src = packed + h[GCB_H_SRC];
  const int32_t* dst = packed + h[GCB_H_DST];
  int32_t* in_ptr = packed + h[GCB_H_IN_PTR];
  int32_t* in_eid = packed + h[GCB_H_IN_EID];
  int32_t* in_src = packed + h[GCB_H_IN_SRC];
  int32_t* out_ptr = packed + h[GCB_H_OUT_PTR];
  int32_t* out_eid = packed + h[GCB_H_OUT_EID];
  int32_t* out_dst = packed + h[GCB_H_OUT_DST];
  int32_t* s4 = packed + h[GCB_H_IN_SRC4];
  int32_t* e4 = packed + h[GCB_H_IN_EID4];
  int32_t* d4 = packed + h[GCB_H_OUT_DST4];
  int32_t* o4 = packed + h[GCB_H_OUT_INPOS4];
  int32_t* out_inpos = packed + h[GCB_H_OUT_INPOS];
  int32_t* graph_ptr = packed + h[GCB_H_GRAPH_PTR];
  int32_t* graph_node = packed + h[GCB_H_GRAPH_NODE];
  int32_t* node_graph = packed + h[GCB_H_NODE_GRAPH];
  int32_t* tile_ptr = packed + h[GCB_H_TILE_PTR];
  int32_t max_a = 0, max_e = 0;
  for (int32_t g = 0; g < G; ++g) { max_a = n_atoms[g] > max_a ? n_atoms[g] : max_a; max_e = n_edges[g] > max_e ? n_edges[g] : max_e; }
  std::vector<int32_t> pin((size_t)max_a + 2), pout((size_t)max_a + 2), cin((size_t)max_a + 1), cout((size_t)max_a + 1),
      pos((size_t)max_e + 1), span((size_t)max_a + 2);
  // closed row tiles: greedy over the cut positions, in order (same walk as finish_pack)
  int32_t nt = 0, start = 0, last_cut = 0;
  bool tiles_ok = true;
  tile_ptr[0] = 0;
  auto cut_at = [&](int32_t p) {
    if (!tiles_ok) return;
    if (p - start > GCB_TILE_ROWS) {
      if (last_cut == start) { tiles_ok = false; return; }
      tile_ptr[++nt] = last_cut;
      start = last_cut;
      if (p - start > GCB_TILE_ROWS) { tiles_ok = false; return; }
    }
    last_cut = p;
  };
  int32_t a0 = 0, e0 = 0;
  for (int32_t g = 0; g < G; ++g) {
    const int32_t na = n_atoms[g], ne = n_edges[g];
    graph_ptr[g] = a0;
    for (int32_t i = 0; i <= na + 1; ++i) { pin[i] = 0; pout[i] = 0; span[i] = 0; }
    for (int32_t e = 0; e < ne; ++e) {
      const int32_t sl = src[e0 + e] - a0, dl = dst[e0 + e] - a0;
      pin[dl + 1]++;
      pout[sl + 1]++;
      const int32_t lo = sl < dl ? sl : dl, hi = sl < dl ? dl : sl;
      if (lo < hi) { span[lo + 1]++; span[hi + 1]--; }
    }
    for (int32_t i = 0; i < na; ++i) { pin[i + 1] += pin[i]; pout[i + 1] += pout[i]; cin[i] = pin[i]; cout[i] = pout[i]; }
    for (int32_t i = 0; i < na; ++i) {
      in_ptr[a0 + i] = e0 + pin[i];
      out_ptr[a0 + i] = e0 + pout[i];
      graph_node[a0 + i] = a0 + i;
      node_graph[a0 + i] = g;
    }
    for (int32_t e = 0; e < ne; ++e) {  // stable: edge-id order inside every row
      const int32_t k = cin[dst[e0 + e] - a0]++;
      in_eid[e0 + k] = e0 + e;
      in_src[e0 + k] = src[e0 + e];
      pos[e] = k;
    }
    for (int32_t e = 0; e < ne; ++e) {
      const int32_t k = cout[src[e0 + e] - a0]++;
      out_eid[e0 + k] = e0 + e;
      out_dst[e0 + k] = dst[e0 + e];
      out_inpos[e0 + k] = e0 + pos[e];
    }
    for (int32_t i = 0; i < na; ++i) {  // ELL-4 views
      const int64_t a = (int64_t)(a0 + i) * 4;
      for (int u = 0; u < 4; ++u) {
        const int32_t ki = pin[i] + u, ko = pout[i] + u;
        const bool hi_ = ki < pin[i + 1], ho_ = ko < pout[i + 1];
        s4[a + u] = hi_ ? in_src[e0 + ki] : -1;
        e4[a + u] = hi_ ? in_eid[e0 + ki] : -1;
        d4[a + u] = ho_ ? out_dst[e0 + ko] : -1;
        o4[a + u] = ho_ ? out_inpos[e0 + ko] : -1;
      }
    }
    int32_t cover = 0;  // cuts inside the molecule (it may be disconnected) and at its end
    for (int32_t p = 1; p <= na; ++p) {
      cover += span[p];
      if (p == na || cover == 0) cut_at(a0 + p);
    }
    a0 += na;
    e0 += ne;
  }
  graph_ptr[G] = N;
  in_ptr[N] = E;
  out_ptr[N] = E;
  if (tiles_ok && N > 0 && last_cut != N) tiles_ok = false;  // (cannot happen: the last molecule ends at N)
  if (tiles_ok && N > 0) tile_ptr[++nt] = N;
  packed[GCB_H_NTILES] = tiles_ok ? nt : 0;
  bucket(packed + h[GCB_H_X_TOK], N, av, packed + h[GCB_H_ATOK_PTR], packed + h[GCB_H_ATOK_NODE]);
  bucket(packed + h[GCB_H_E_TOK], M, bv, packed + h[GCB_H_BTOK_PTR], packed + h[GCB_H_BTOK_ROW]);
  return GCB_OK;
}

}  // namespace gcb

using namespace gcb;

extern "C" int64_t gcb_pack_ints(int32_t N, int32_t E, int32_t G, int32_t M, int32_t atom_vocab,
                                 int32_t bond_vocab, int32_t* header) {
  if (N < 0 || E < 0 || G < 0 || M < 0 || M > E || atom_vocab < 0 || bond_vocab < 0) return GCB_ERR_INVALID_ARG;
  return layout(N, E, G, M, atom_vocab, bond_vocab, 0, header);
}

extern "C" int64_t gcb_pack_ints_flags(int32_t N, int32_t E, int32_t G, int32_t M, int32_t atom_vocab,
                                       int32_t bond_vocab, int32_t flags, int32_t* header) {
  if (N < 0 || E < 0 || G < 0 || M < 0 || M > E || atom_vocab < 0 || bond_vocab < 0 || (flags & ~GCB_PACK_EACH))
    return GCB_ERR_INVALID_ARG;
  return layout(N, E, G, M, atom_vocab, bond_vocab, flags, header);
}

extern "C" int gcb_pack_host(const void* x, const void* edge_attr, const void* edge_index, const void* batch,
                             int32_t idx_bytes, int32_t N, int32_t E, int32_t G, int32_t M, int32_t atom_vocab,
                             int32_t bond_vocab, int32_t check_bond_rows, int32_t* packed, int64_t packed_ints) {
  if (idx_bytes != 4 && idx_bytes != 8) return GCB_ERR_INVALID_ARG;
  if (!packed || (N > 0 && !x) || (E > 0 && (!edge_attr || !edge_index))) return GCB_ERR_INVALID_ARG;
  int32_t h[GCB_PACK_HEADER_INTS];
  int64_t total = gcb_pack_ints(N, E, G, M, atom_vocab, bond_vocab, h);
  if (total < 0) return GCB_ERR_INVALID_ARG;
  if (packed_ints < total) return GCB_ERR_WORKSPACE;
  std::memcpy(packed, h, sizeof(h));
  int32_t* x_tok = packed + h[GCB_H_X_TOK];
  int32_t* e_tok = packed + h[GCB_H_E_TOK];
  int32_t* src = packed + h[GCB_H_SRC];
  int32_t* dst = packed + h[GCB_H_DST];
  for (int64_t i = 0; i < N; ++i) {
    int64_t t = read_idx(x, idx_bytes, i);
    if (t < 0 || t >= atom_vocab) return GCB_ERR_INDEX;  // nn.Embedding: index out of range
    x_tok[i] = (int32_t)t;
  }
  for (int64_t e = 0; e < E; ++e) {
    int64_t t = read_idx(edge_attr, idx_bytes, e);
    if (t < 0 || t >= bond_vocab) return GCB_ERR_INDEX;
    e_tok[e] = (int32_t)t;
  }
  const int64_t lim = check_bond_rows ? (N < E ? N : E) : N;
  for (int64_t e = 0; e < E; ++e) {
    int64_t s = read_idx(edge_index, idx_bytes, e), d = read_idx(edge_index, idx_bytes, (int64_t)E + e);
    if (s < 0 || d < 0 || s >= lim || d >= lim) return GCB_ERR_INDEX;  // PyG _lift IndexError
    src[e] = (int32_t)s;
    dst[e] = (int32_t)d;
  }
  std::vector<int32_t> b32;
  if (batch) {
    b32.resize((size_t)N);
    for (int64_t i = 0; i < N; ++i) {
      int64_t g = read_idx(batch, idx_bytes, i);
      if (g < 0 || g >= G) return GCB_ERR_INDEX;
      b32[(size_t)i] = (int32_t)g;
    }
  }
  const int rc = finish_pack(packed, h, batch ? b32.data() : nullptr);
  if (rc == GCB_OK) zero_unwritten(packed, h, E);
  return rc;
}

extern "C" int gcb_collate_pack_host(const int32_t* const* atoms, const int32_t* n_atoms,
                                     const int32_t* const* bonds, const int64_t* const* conn,
                                     const int32_t* n_edges, const float* const* y, int32_t n_targets,
                                     int32_t n_graphs, int32_t n_messages, int32_t atom_vocab,
                                     int32_t bond_vocab, int32_t* packed, int64_t packed_ints, float* y_out) {
  if (n_graphs < 0 || !packed || (n_graphs > 0 && (!atoms || !n_atoms || !bonds || !conn || !n_edges)))
    return GCB_ERR_INVALID_ARG;
  int64_t N64 = 0, E64 = 0;
  for (int32_t g = 0; g < n_graphs; ++g) {
    if (n_atoms[g] < 0 || n_edges[g] < 0) return GCB_ERR_INVALID_ARG;
    N64 += n_atoms[g];
    E64 += n_edges[g];
  }
  if (N64 > INT32_MAX || E64 > INT32_MAX) return GCB_ERR_INVALID_ARG;
  const int32_t N = (int32_t)N64, E = (int32_t)E64, G = n_graphs;
  const int32_t M = n_messages >= 1 ? (N < E ? N : E) : E;
  int32_t h[GCB_PACK_HEADER_INTS];
  int64_t total = gcb_pack_ints(N, E, G, M, atom_vocab, bond_vocab, h);
  if (total < 0) return GCB_ERR_INVALID_ARG;
  if (packed_ints < total) return GCB_ERR_WORKSPACE;
  std::memcpy(packed, h, sizeof(h));
  int32_t* x_tok = packed + h[GCB_H_X_TOK];
  int32_t* e_tok = packed + h[GCB_H_E_TOK];
  int32_t* src = packed + h[GCB_H_SRC];
  int32_t* dst = packed + h[GCB_H_DST];
  const int64_t lim = n_messages >= 1 ? M : N;
  int64_t node_off = 0, edge_off = 0;
  bool local = true;  // every edge stays inside its molecule (always, for encoder output)
  for (int32_t g = 0; g < G; ++g) {
    const int32_t na = n_atoms[g], ne = n_edges[g];
    for (int32_t i = 0; i < na; ++i) {
      int32_t t = atoms[g][i];
      if (t < 0 || t >= atom_vocab) return GCB_ERR_INDEX;
      x_tok[node_off + i] = t;
    }
    for (int32_t e = 0; e < ne; ++e) {
      int32_t t = bonds[g][e];
      if (t < 0 || t >= bond_vocab) return GCB_ERR_INDEX;
      e_tok[edge_off + e] = t;
      int64_t s = conn[g][e] + node_off, d = conn[g][(int64_t)ne + e] + node_off;  // [2,e] row-major
      if (conn[g][e] < 0 || conn[g][(int64_t)ne + e] < 0 || s >= lim || d >= lim) return GCB_ERR_INDEX;
      local = local && conn[g][e] < na && conn[g][(int64_t)ne + e] < na;
      src[edge_off + e] = (int32_t)s;
      dst[edge_off + e] = (int32_t)d;
    }
    if (y_out) {
      for (int32_t t = 0; t < n_targets; ++t) y_out[(int64_t)g * n_targets + t] = y && y[g] ? y[g][t] : 0.f;
    }
    node_off += na;
    edge_off += ne;
  }
  int rc;
  if (local) {
    rc = finish_pack_local(packed, h, n_atoms, n_edges);
  } else {  // an edge into another molecule's atoms (what PyG's collate would also produce): the general builder
    std::vector<int32_t> b32((size_t)N);
    int64_t off = 0;
    for (int32_t g = 0; g < G; ++g) { for (int32_t i = 0; i < n_atoms[g]; ++i) b32[(size_t)(off + i)] = g; off += n_atoms[g]; }
    rc = finish_pack(packed, h, b32.data());
  }
  if (rc == GCB_OK) zero_unwritten(packed, h, E);
  return rc;
}

// "Each molecule alone": bond row r of the packed batch = the molecule's own edge (r - its first atom).
extern "C" int gcb_collate_pack_each_host(const int32_t* const* atoms, const int32_t* n_atoms,
                                          const int32_t* const* bonds, const int64_t* const* conn,
                                          const int32_t* n_edges, const float* const* y, int32_t n_targets,
                                          int32_t n_graphs, int32_t n_messages, int32_t atom_vocab,
                                          int32_t bond_vocab, int32_t* packed, int64_t packed_ints, float* y_out) {
  if (n_graphs < 0 || !packed || (n_graphs > 0 && (!atoms || !n_atoms || !bonds || !conn || !n_edges)))
    return GCB_ERR_INVALID_ARG;
  if (n_messages < 1) return GCB_ERR_INVALID_ARG;  // without EdgeConv on the bond matrix batched == per molecule already
  int64_t N64 = 0, E64 = 0;
  for (int32_t g = 0; g < n_graphs; ++g) {
    if (n_atoms[g] < 0 || n_edges[g] < 0) return GCB_ERR_INVALID_ARG;
    N64 += n_atoms[g];
    E64 += n_edges[g];
  }
  if (N64 > INT32_MAX || E64 > INT32_MAX) return GCB_ERR_INVALID_ARG;
  if (E64 < N64) return GCB_ERR_UNSUPPORTED;  // the bond matrix must offer one row per atom
  const int32_t N = (int32_t)N64, E = (int32_t)E64, G = n_graphs, M = N;
  int32_t h[GCB_PACK_HEADER_INTS];
  int64_t total = gcb_pack_ints_flags(N, E, G, M, atom_vocab, bond_vocab, GCB_PACK_EACH, h);
  if (total < 0) return GCB_ERR_INVALID_ARG;
  if (packed_ints < total) return GCB_ERR_WORKSPACE;
  std::memcpy(packed, h, sizeof(h));
  int32_t* x_tok = packed + h[GCB_H_X_TOK];
  int32_t* e_tok = packed + h[GCB_H_E_TOK];
  int32_t* src = packed + h[GCB_H_SRC];
  int32_t* dst = packed + h[GCB_H_DST];
  int32_t* bond_pos = packed + h[GCB_H_BOND_POS];
  std::memset(e_tok, 0, sizeof(int32_t) * (size_t)N);  // rows of atoms beyond a molecule's last edge keep token 0
  std::vector<int32_t> b32((size_t)N);
  int64_t node_off = 0, edge_off = 0;
  for (int32_t g = 0; g < G; ++g) {
    const int32_t na = n_atoms[g], ne = n_edges[g];
    for (int32_t i = 0; i < na; ++i) {
      int32_t t = atoms[g][i];
      if (t < 0 || t >= atom_vocab) return GCB_ERR_INDEX;
      x_tok[node_off + i] = t;
      b32[(size_t)(node_off + i)] = g;
      // bond row of atom i = the molecule's edge i (rows past its last edge are never addressed: see the check below)
      if (i < ne) {
        int32_t bt = bonds[g][i];
        if (bt < 0 || bt >= bond_vocab) return GCB_ERR_INDEX;
        e_tok[node_off + i] = bt;
      }
    }
    const int32_t lim = na < ne ? na : ne;  // live bond rows of this molecule alone: min(n_atoms, n_edges)
    for (int32_t e = 0; e < ne; ++e) {
      int32_t t = bonds[g][e];
      if (t < 0 || t >= bond_vocab) return GCB_ERR_INDEX;
      const int64_t s = conn[g][e], d = conn[g][(int64_t)ne + e];
      if (s < 0 || d < 0 || s >= lim || d >= lim) return GCB_ERR_INDEX;  // PyG _lift IndexError for this molecule alone
      src[edge_off + e] = (int32_t)(s + node_off);
      dst[edge_off + e] = (int32_t)(d + node_off);
      bond_pos[edge_off + e] = e < na ? (int32_t)(node_off + e) : -1;
    }
    if (y_out) {
      for (int32_t t = 0; t < n_targets; ++t) y_out[(int64_t)g * n_targets + t] = y && y[g] ? y[g][t] : 0.f;
    }
    node_off += na;
    edge_off += ne;
  }
  int rc = finish_pack(packed, h, b32.data());
  if (rc != GCB_OK) return rc;
  // the atom step reads the updated bond rows by EDGE id: send every CSR entry to the row that holds its edge
  int32_t* in_eid = packed + h[GCB_H_IN_EID];
  int32_t* e4 = packed + h[GCB_H_IN_EID4];
  for (int64_t k = 0; k < E; ++k) in_eid[k] = bond_pos[in_eid[k]] >= 0 ? bond_pos[in_eid[k]] : M;
  for (int64_t k = 0; k < 4 * (int64_t)N; ++k)
    if (e4[k] >= 0) e4[k] = bond_pos[e4[k]] >= 0 ? bond_pos[e4[k]] : M;
  zero_unwritten(packed, h, N);
  return GCB_OK;
}
