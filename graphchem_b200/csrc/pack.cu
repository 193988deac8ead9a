// Host-side packer: PyG-style Batch -> one int32 buffer with tokens, edge list, destination- and
// source-sorted CSR, graph segments and token buckets.  Pure integer work; must match
// oracle/csr_oracle.py (stable argsort) bit for bit.
//
// Reference semantics restated here (not code): PyG scatters messages in edge-id order over
// edge_index[1], collate concatenates graphs with cumulative node offsets
// (call site examples/predict_cn.ipynb:212,245), MoleculeEncoder.encode emits
// int32 tokens and an int64 [2,e] connectivity (graphchem/preprocessing/features.py:294-319).
#include <cstdlib>
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
  put(GCB_H_OUT_INPOS, E);
  put(GCB_H_TILE_PTR, (int64_t)N + 1);
  if (flags & GCB_PACK_EACH) put(GCB_H_BOND_POS, E);
  // derived tail (GCB_H_COMPACT_INTS = offset of the first of them): a function of the CSRs above, see gcb_pack_expand_ell
  put(GCB_H_IN_SRC4, 4 * (int64_t)N); put(GCB_H_IN_EID4, 4 * (int64_t)N); put(GCB_H_OUT_DST4, 4 * (int64_t)N);
  put(GCB_H_OUT_INPOS4, 4 * (int64_t)N);
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
  if (h) { h[GCB_H_TOTAL_INTS] = (int32_t)off; h[GCB_H_COMPACT_INTS] = h[GCB_H_IN_SRC4]; }
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

// Greedy closed-tile walk over cut positions given in increasing order (rows no edge spans): a tile is closed at
// the last cut that keeps it within GCB_TILE_ROWS rows.  Shared by the collate packer and the raw collate of the
// device packer, so both produce the same tile_ptr.
struct TileWalker {
  int32_t* tile_ptr;
  int32_t nt = 0, start = 0, last_cut = 0;
  bool ok = true;
  explicit TileWalker(int32_t* tp) : tile_ptr(tp) { tile_ptr[0] = 0; }
  void cut_at(int32_t p) {
    if (!ok) return;
    if (p - start > GCB_TILE_ROWS) {
      if (last_cut == start) { ok = false; return; }
      tile_ptr[++nt] = last_cut;
      start = last_cut;
      if (p - start > GCB_TILE_ROWS) { ok = false; return; }
    }
    last_cut = p;
  }
  int32_t finish(int32_t N) {  // number of tiles, 0 = no closed tiling (a component wider than a tile)
    if (ok && N > 0 && last_cut != N) ok = false;  // (cannot happen: the last molecule ends at N)
    if (ok && N > 0) tile_ptr[++nt] = N;
    return ok ? nt : 0;
  }
};

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
  TileWalker tw(tile_ptr);  // closed row tiles: greedy over the cut positions, in order (same walk as finish_pack)
  auto cut_at = [&](int32_t p) { tw.cut_at(p); };
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
  packed[GCB_H_NTILES] = tw.finish(N);
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

// =====================================================================================================
// Device-side packer (SURVEY.md 8 d config 5: batches "packed on device"; 8 f1).  The host only gathers the
// molecules of a batch into a RAW buffer -- molecule offsets, tokens, edge list with batch-global atom ids and
// the closed-tile table (0.7 kB per molecule instead of 3.9 kB) -- and the device derives every other
// section of the packed batch.  The result is bit-identical to gcb_collate_pack_host on the same molecules.
//
// The per-molecule and per-chunk routines are __host__ __device__: gcb_debug_pack_emulate_host runs the code of
// the kernels serially on the CPU (a test hook like gcb_debug_gemm_*, not a product path), so the index logic is
// pinned against the host packer without a GPU.  Only the scan over the chunk histograms differs on the device
// (block-parallel instead of a serial walk); the -m gpu tests pin that one.
// =====================================================================================================
namespace gcb {

enum {
  RAW_MAGIC = 0, RAW_N, RAW_E, RAW_G, RAW_M, RAW_AV, RAW_BV, RAW_NTILES, RAW_ATOM_OFF, RAW_EDGE_OFF, RAW_X_TOK,
  RAW_E_TOK, RAW_SRC, RAW_DST, RAW_TILE_PTR, RAW_TOTAL_INTS
};
static_assert(RAW_TOTAL_INTS < GCB_RAW_HEADER_INTS, "raw header too small");

static inline int64_t raw_tile_cap(int32_t N) { return (int64_t)N / (GCB_TILE_ROWS / 2) + 3; }  // > tiles + 1 of any greedy walk

static int64_t raw_layout(int32_t N, int32_t E, int32_t G, int32_t* rh) {
  int64_t off = GCB_RAW_HEADER_INTS;
  const int64_t n[7] = {(int64_t)G + 1, (int64_t)G + 1, N, E, E, E, raw_tile_cap(N)};
  for (int i = 0; i < 7; ++i) {
    if (rh) rh[RAW_ATOM_OFF + i] = (int32_t)off;
    off += pad4(n[i]);
  }
  if (off > INT32_MAX) return -1;
  if (rh) rh[RAW_TOTAL_INTS] = (int32_t)off;
  return off;
}

struct PackPtrs {
  const int32_t *src, *dst;
  int32_t *in_ptr, *in_eid, *in_src, *out_ptr, *out_eid, *out_dst, *s4, *e4, *d4, *o4, *out_inpos, *graph_ptr,
      *graph_node, *node_graph;
};

static PackPtrs pack_ptrs(int32_t* packed, const int32_t* h) {
  PackPtrs p;
  p.src = packed + h[GCB_H_SRC]; p.dst = packed + h[GCB_H_DST];
  p.in_ptr = packed + h[GCB_H_IN_PTR]; p.in_eid = packed + h[GCB_H_IN_EID]; p.in_src = packed + h[GCB_H_IN_SRC];
  p.out_ptr = packed + h[GCB_H_OUT_PTR]; p.out_eid = packed + h[GCB_H_OUT_EID]; p.out_dst = packed + h[GCB_H_OUT_DST];
  p.s4 = packed + h[GCB_H_IN_SRC4]; p.e4 = packed + h[GCB_H_IN_EID4]; p.d4 = packed + h[GCB_H_OUT_DST4];
  p.o4 = packed + h[GCB_H_OUT_INPOS4]; p.out_inpos = packed + h[GCB_H_OUT_INPOS];
  p.graph_ptr = packed + h[GCB_H_GRAPH_PTR]; p.graph_node = packed + h[GCB_H_GRAPH_NODE];
  p.node_graph = packed + h[GCB_H_NODE_GRAPH];
  return p;
}

// One molecule (atoms [a0, a0+na), edges [e0, e0+ne), every edge inside the molecule): both CSRs in stable
// edge-id order, the cross index, the ELL-4 views, graph segment entries.  pin / pout: na + 1 ints, pos: ne ints.
__host__ __device__ inline void pack_one_molecule(const PackPtrs& p, int32_t g, int32_t a0, int32_t na, int32_t e0,
                                                  int32_t ne, int32_t* pin, int32_t* pout, int32_t* pos) {
  p.graph_ptr[g] = a0;
  for (int32_t i = 0; i <= na; ++i) { pin[i] = 0; pout[i] = 0; }
  for (int32_t e = 0; e < ne; ++e) {  // counts of row i at index i + 1
    pin[p.dst[e0 + e] - a0 + 1]++;
    pout[p.src[e0 + e] - a0 + 1]++;
  }
  for (int32_t i = 0; i < na; ++i) { pin[i + 1] += pin[i]; pout[i + 1] += pout[i]; }  // pin[i] = first entry of row i
  for (int32_t i = 0; i < na; ++i) {
    p.in_ptr[a0 + i] = e0 + pin[i];
    p.out_ptr[a0 + i] = e0 + pout[i];
    p.graph_node[a0 + i] = a0 + i;
    p.node_graph[a0 + i] = g;
  }
  for (int32_t e = 0; e < ne; ++e) {  // pin[] turns into the row cursors: afterwards pin[i] = end of row i
    const int32_t k = pin[p.dst[e0 + e] - a0]++;
    p.in_eid[e0 + k] = e0 + e;
    p.in_src[e0 + k] = p.src[e0 + e];
    pos[e] = k;
  }
  for (int32_t e = 0; e < ne; ++e) {
    const int32_t k = pout[p.src[e0 + e] - a0]++;
    p.out_eid[e0 + k] = e0 + e;
    p.out_dst[e0 + k] = p.dst[e0 + e];
    p.out_inpos[e0 + k] = e0 + pos[e];
  }
  for (int32_t i = 0; i < na; ++i) {  // ELL-4 views: first four entries of every row, -1 padded
    const int32_t bi = i ? pin[i - 1] : 0, ei = pin[i], bo = i ? pout[i - 1] : 0, eo = pout[i];
    const int64_t a = (int64_t)(a0 + i) * 4;
    for (int u = 0; u < 4; ++u) {
      const int32_t ki = bi + u, ko = bo + u;
      p.s4[a + u] = ki < ei ? p.in_src[e0 + ki] : -1;
      p.e4[a + u] = ki < ei ? p.in_eid[e0 + ki] : -1;
      p.d4[a + u] = ko < eo ? p.out_dst[e0 + ko] : -1;
      p.o4[a + u] = ko < eo ? p.out_inpos[e0 + ko] : -1;
    }
  }
}

// The same molecule packed by the 32 lanes of a warp, as a sequence of phases that only communicate through memory:
// the kernel runs phase, __syncwarp(), phase, ...; the emulation hook runs every phase for lanes 0..31 in turn, which is
// equivalent.  w: per-warp work area of kWarpWork ints (pin | pout | pos | local src | local dst); the molecule must
// satisfy na <= kPackMaxA and ne <= kPackMaxE.  Results are identical to pack_one_molecule (the default device packer).
constexpr int kPackMaxA = 64, kPackMaxE = 160;  // molecules up to this size use thread-local / per-warp arrays, larger ones scratch
constexpr int kWarpWork = 2 * (kPackMaxA + 1) + 3 * kPackMaxE;
constexpr int kWarpPhases = 6;
__host__ __device__ inline void warp_count_inc(int32_t* p) {
#ifdef __CUDA_ARCH__
  atomicAdd(p, 1);
#else
  ++*p;
#endif
}
__host__ __device__ inline void pack_warp_phase(int phase, int lane, const PackPtrs& p, int32_t g, int32_t a0, int32_t na,
                                                int32_t e0, int32_t ne, int32_t* w) {
  int32_t* pin = w;
  int32_t* pout = w + (kPackMaxA + 1);
  int32_t* pos = w + 2 * (kPackMaxA + 1);
  int32_t* ls = pos + kPackMaxE;  // molecule-local source / destination atom of every edge
  int32_t* ld = ls + kPackMaxE;
  switch (phase) {
    case 0:  // clear the counters, fetch the edge list
      for (int32_t i = lane; i <= na; i += 32) { pin[i] = 0; pout[i] = 0; }
      for (int32_t e = lane; e < ne; e += 32) { ls[e] = p.src[e0 + e] - a0; ld[e] = p.dst[e0 + e] - a0; }
      if (lane == 0) p.graph_ptr[g] = a0;
      break;
    case 1:  // degree counts of row i at index i + 1 (integer atomics: the result does not depend on their order)
      for (int32_t e = lane; e < ne; e += 32) { warp_count_inc(&pin[ld[e] + 1]); warp_count_inc(&pout[ls[e] + 1]); }
      break;
    case 2:  // prefix sums: pin[i] = first entry of row i (two lanes, one array each)
      if (lane < 2) {
        int32_t* c = lane ? pout : pin;
        for (int32_t i = 0; i < na; ++i) c[i + 1] += c[i];
      }
      break;
    case 3:  // row pointers, graph segments
      for (int32_t i = lane; i < na; i += 32) {
        p.in_ptr[a0 + i] = e0 + pin[i];
        p.out_ptr[a0 + i] = e0 + pout[i];
        p.graph_node[a0 + i] = a0 + i;
        p.node_graph[a0 + i] = g;
      }
      break;
    case 4:  // stable placement: edge-id order inside every row, so one lane walks each CSR; cursors end at the row ends
      if (lane == 0) {
        for (int32_t e = 0; e < ne; ++e) {
          const int32_t k = pin[ld[e]]++;
          p.in_eid[e0 + k] = e0 + e;
          p.in_src[e0 + k] = a0 + ls[e];
          pos[e] = k;
        }
      } else if (lane == 1) {
        for (int32_t e = 0; e < ne; ++e) {
          const int32_t k = pout[ls[e]]++;
          p.out_eid[e0 + k] = e0 + e;
          p.out_dst[e0 + k] = a0 + ld[e];
        }
      }
      break;
    case 5:  // cross index and ELL-4 views
      for (int32_t k = lane; k < ne; k += 32) p.out_inpos[e0 + k] = e0 + pos[p.out_eid[e0 + k] - e0];
      break;
    default: break;
  }
}
// Last phase, separate because it reads out_inpos written by other lanes in phase 5.
__host__ __device__ inline void pack_warp_ell(int lane, const PackPtrs& p, int32_t a0, int32_t na, int32_t e0, const int32_t* w) {
  const int32_t* pin = w;
  const int32_t* pout = w + (kPackMaxA + 1);
  for (int32_t i = lane; i < na; i += 32) {
    const int32_t bi = i ? pin[i - 1] : 0, ei = pin[i], bo = i ? pout[i - 1] : 0, eo = pout[i];
    const int64_t a = (int64_t)(a0 + i) * 4;
    for (int u = 0; u < 4; ++u) {
      const int32_t ki = bi + u, ko = bo + u;
      p.s4[a + u] = ki < ei ? p.in_src[e0 + ki] : -1;
      p.e4[a + u] = ki < ei ? p.in_eid[e0 + ki] : -1;
      p.d4[a + u] = ko < eo ? p.out_dst[e0 + ko] : -1;
      p.o4[a + u] = ko < eo ? p.out_inpos[e0 + ko] : -1;
    }
  }
}

// Stable counting sort of ids 0..n-1 by key in three passes over chunks of kBucketChunk consecutive ids:
// (A) per-chunk histogram, (B) per-bucket exclusive scan over the chunks + bucket pointers, (C) placement in id order.
constexpr int kBucketChunk = 64;
__host__ __device__ inline void bucket_hist_chunk(const int32_t* key, int64_t n, int32_t nb, int64_t c, int32_t* hist) {
  int32_t* h = hist + c * nb;
  for (int32_t b = 0; b < nb; ++b) h[b] = 0;
  const int64_t i1 = (c + 1) * kBucketChunk < n ? (c + 1) * kBucketChunk : n;
  for (int64_t i = c * kBucketChunk; i < i1; ++i) h[key[i]]++;
}
__host__ __device__ inline void bucket_scan_bucket(int32_t nb, int64_t chunks, int32_t b, int32_t* hist, int32_t* ptr) {
  int32_t run = 0;
  for (int64_t c = 0; c < chunks; ++c) {
    const int32_t t = hist[c * nb + b];
    hist[c * nb + b] = run;
    run += t;
  }
  ptr[b + 1] = run;  // bucket size; bucket_ptr_scan turns the sizes into offsets
}
__host__ __device__ inline void bucket_ptr_scan(int32_t nb, int32_t* ptr) {
  ptr[0] = 0;
  for (int32_t b = 0; b < nb; ++b) ptr[b + 1] += ptr[b];
}
__host__ __device__ inline void bucket_place_chunk(const int32_t* key, int64_t n, int32_t nb, int64_t c, int32_t* hist,
                                                   const int32_t* ptr, int32_t* perm) {
  int32_t* h = hist + c * nb;
  const int64_t i1 = (c + 1) * kBucketChunk < n ? (c + 1) * kBucketChunk : n;
  for (int64_t i = c * kBucketChunk; i < i1; ++i) {
    const int32_t k = key[i];
    perm[ptr[k] + h[k]++] = (int32_t)i;
  }
}
static inline int64_t bucket_chunks(int64_t n) { return (n + kBucketChunk - 1) / kBucketChunk; }

// Copies (raw -> packed) and zero fills that complete a packed batch around the derived sections.
constexpr int kMaxSegs = 48;
struct PackSegs {
  int32_t header[GCB_PACK_HEADER_INTS];
  int32_t n;
  int32_t dst_off[kMaxSegs], src_off[kMaxSegs] /* -1: zero fill */, count[kMaxSegs];
  int32_t in_ptr_end, out_ptr_end, graph_ptr_end;  // positions of in_ptr[N], out_ptr[N], graph_ptr[G]
};
static int build_segs(const int32_t* h, const int32_t* rh, PackSegs* sg) {
  const int32_t N = h[GCB_H_N], E = h[GCB_H_E], G = h[GCB_H_G], M = h[GCB_H_M], nt = h[GCB_H_NTILES];
  std::memcpy(sg->header, h, sizeof(sg->header));
  sg->n = 0;
  auto put = [&](int64_t dst, int64_t src, int64_t count) {
    if (count <= 0) return true;
    if (sg->n >= kMaxSegs) return false;
    sg->dst_off[sg->n] = (int32_t)dst; sg->src_off[sg->n] = (int32_t)src; sg->count[sg->n] = (int32_t)count;
    ++sg->n;
    return true;
  };
  bool ok = put(h[GCB_H_X_TOK], rh[RAW_X_TOK], N) && put(h[GCB_H_E_TOK], rh[RAW_E_TOK], E) &&
            put(h[GCB_H_SRC], rh[RAW_SRC], E) && put(h[GCB_H_DST], rh[RAW_DST], E) &&
            put(h[GCB_H_TILE_PTR], rh[RAW_TILE_PTR], (int64_t)nt + 1);
  Sec t[GCB_H_FIELDS];  // same fills as zero_unwritten()
  const int k = section_table(N, E, G, h[GCB_H_AV], h[GCB_H_BV], h[GCB_H_FLAGS], t);
  for (int i = 0; i < k; ++i) ok = ok && put(h[t[i].field] + t[i].n, -1, pad4(t[i].n) - t[i].n);
  ok = ok && put((int64_t)h[GCB_H_BTOK_ROW] + M, -1, (int64_t)E - M);
  ok = ok && put((int64_t)h[GCB_H_TILE_PTR] + nt + 1, -1, (int64_t)N + 1 - (nt + 1));
  sg->in_ptr_end = h[GCB_H_IN_PTR] + N;
  sg->out_ptr_end = h[GCB_H_OUT_PTR] + N;
  sg->graph_ptr_end = h[GCB_H_GRAPH_PTR] + G;
  return ok ? GCB_OK : GCB_ERR_INVALID_ARG;
}
__host__ __device__ inline void seg_element(const PackSegs& sg, int s, int64_t i, const int32_t* raw, int32_t* packed) {
  packed[(int64_t)sg.dst_off[s] + i] = sg.src_off[s] >= 0 ? raw[(int64_t)sg.src_off[s] + i] : 0;
}
__host__ __device__ inline void seg_fixed(const PackSegs& sg, int32_t* packed) {  // header and the three end sentinels
  for (int i = 0; i < GCB_PACK_HEADER_INTS; ++i) packed[i] = sg.header[i];
  packed[sg.in_ptr_end] = sg.header[GCB_H_E];
  packed[sg.out_ptr_end] = sg.header[GCB_H_E];
  packed[sg.graph_ptr_end] = sg.header[GCB_H_N];
}

// ---- kernels: one thread per unit of the routines above -------------------------------------------------
__global__ void pack_segments_kernel(const __grid_constant__ PackSegs sg, const int32_t* __restrict__ raw, int32_t* __restrict__ packed) {
  const int s = blockIdx.y;
  if (s == sg.n) {
    if (blockIdx.x == 0 && threadIdx.x == 0) seg_fixed(sg, packed);
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < sg.count[s]; i += (int64_t)gridDim.x * blockDim.x)
    seg_element(sg, s, i, raw, packed);
}
__global__ void pack_molecules_kernel(const PackPtrs p, const int32_t* __restrict__ atom_off, const int32_t* __restrict__ edge_off,
                                      int32_t G, int32_t* scr_pin, int32_t* scr_pout, int32_t* scr_pos) {
  const int32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const int32_t a0 = atom_off[g], na = atom_off[g + 1] - a0, e0 = edge_off[g], ne = edge_off[g + 1] - e0;
  if (na <= kPackMaxA && ne <= kPackMaxE) {
    int32_t pin[kPackMaxA + 1], pout[kPackMaxA + 1], pos[kPackMaxE];
    pack_one_molecule(p, g, a0, na, e0, ne, pin, pout, pos);
  } else {
    pack_one_molecule(p, g, a0, na, e0, ne, scr_pin + a0 + g, scr_pout + a0 + g, scr_pos + e0);
  }
}
// Warp per molecule (default; GCB_PACK_WARP=0 selects the thread-per-molecule kernel above): the phases of pack_warp_phase with a __syncwarp() between them, the
// work area in shared memory; a molecule beyond the per-warp limits is packed by lane 0 alone out of global scratch.
constexpr int kPackWarps = 8;
__global__ void __launch_bounds__(kPackWarps * 32)
pack_molecules_warp_kernel(const PackPtrs p, const int32_t* __restrict__ atom_off, const int32_t* __restrict__ edge_off,
                           int32_t G, int32_t* scr_pin, int32_t* scr_pout, int32_t* scr_pos) {
  __shared__ int32_t work[kPackWarps][kWarpWork];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int32_t g = blockIdx.x * kPackWarps + wid;
  if (g >= G) return;  // whole warps leave together
  const int32_t a0 = atom_off[g], na = atom_off[g + 1] - a0, e0 = edge_off[g], ne = edge_off[g + 1] - e0;
  if (na > kPackMaxA || ne > kPackMaxE) {
    if (lane == 0) pack_one_molecule(p, g, a0, na, e0, ne, scr_pin + a0 + g, scr_pout + a0 + g, scr_pos + e0);
    return;
  }
  int32_t* w = work[wid];
  for (int phase = 0; phase < kWarpPhases; ++phase) {
    pack_warp_phase(phase, lane, p, g, a0, na, e0, ne, w);
    __syncwarp();
  }
  pack_warp_ell(lane, p, a0, na, e0, w);
}
__global__ void bucket_hist_kernel(const int32_t* __restrict__ key, int64_t n, int32_t nb, int64_t chunks, int32_t* hist) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < chunks) bucket_hist_chunk(key, n, nb, c, hist);
}
// Pass B on the device: one block per bucket scans that bucket's column of the chunk histograms (a serial walk
// over thousands of chunks is pure memory latency -- 400 us alone and milliseconds next to a running train step),
// then one block turns the bucket sizes into offsets.  Same results as bucket_scan_bucket / bucket_ptr_scan.
constexpr int kScanThreads = 256;
__global__ void __launch_bounds__(kScanThreads) bucket_scan_kernel(int32_t nb, int64_t chunks, int32_t* hist, int32_t* ptr) {
  __shared__ int32_t part[kScanThreads];
  const int32_t b = blockIdx.x;
  const int t = threadIdx.x;
  const int64_t per = (chunks + kScanThreads - 1) / kScanThreads;
  const int64_t c0 = t * per < chunks ? t * per : chunks, c1 = c0 + per < chunks ? c0 + per : chunks;
  int32_t sum = 0;
  for (int64_t c = c0; c < c1; ++c) sum += hist[c * nb + b];
  part[t] = sum;
  __syncthreads();
  for (int off = 1; off < kScanThreads; off <<= 1) {  // inclusive scan of the per-thread sums
    const int32_t v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int32_t run = part[t] - sum;  // chunks before this thread's segment
  for (int64_t c = c0; c < c1; ++c) {
    const int32_t v = hist[c * nb + b];
    hist[c * nb + b] = run;
    run += v;
  }
  if (t == kScanThreads - 1) ptr[b + 1] = part[t];  // bucket size
}
__global__ void __launch_bounds__(kScanThreads) bucket_ptr_kernel(int32_t nb, int32_t* ptr) {  // ONE block
  __shared__ int32_t sm[1024];
  __shared__ int32_t carry;
  const int t = threadIdx.x;
  if (t == 0) { carry = 0; ptr[0] = 0; }
  __syncthreads();
  for (int32_t base = 0; base < nb; base += 1024) {
    const int32_t n = nb - base < 1024 ? nb - base : 1024;
    for (int32_t i = t; i < n; i += kScanThreads) sm[i] = ptr[base + 1 + i];
    __syncthreads();
    if (t == 0) {
      int32_t run = carry;
      for (int32_t i = 0; i < n; ++i) { run += sm[i]; sm[i] = run; }
      carry = run;
    }
    __syncthreads();
    for (int32_t i = t; i < n; i += kScanThreads) ptr[base + 1 + i] = sm[i];
    __syncthreads();
  }
}
__global__ void bucket_place_kernel(const int32_t* __restrict__ key, int64_t n, int32_t nb, int64_t chunks, int32_t* hist,
                                    const int32_t* __restrict__ ptr, int32_t* perm) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < chunks) bucket_place_chunk(key, n, nb, c, hist, ptr, perm);
}

static int bucket_device(const int32_t* key, int64_t n, int32_t nb, int32_t* ptr, int32_t* perm, int32_t* hist, cudaStream_t s) {
  const int64_t chunks = bucket_chunks(n);
  const int bt = 64;
  if (chunks > 0)
    GCB_KERNEL("pack_bucket_hist_kernel", s, bucket_hist_kernel<<<(unsigned)((chunks + bt - 1) / bt), bt, 0, s>>>(key, n, nb, chunks, hist));
  if (nb > 0) GCB_KERNEL("pack_bucket_scan_kernel", s, bucket_scan_kernel<<<(unsigned)nb, kScanThreads, 0, s>>>(nb, chunks, hist, ptr));
  GCB_KERNEL("pack_bucket_ptr_kernel", s, bucket_ptr_kernel<<<1, kScanThreads, 0, s>>>(nb, ptr));
  if (chunks > 0)
    GCB_KERNEL("pack_bucket_place_kernel", s, bucket_place_kernel<<<(unsigned)((chunks + bt - 1) / bt), bt, 0, s>>>(key, n, nb, chunks, hist, ptr, perm));
  return GCB_OK;
}

// With the warp-per-molecule kernel (default): the atom and the bond token sort share their launches (blockIdx.y / the upper half of
// the scan grid selects the job), 4 launches instead of 8; the per-chunk / per-bucket routines are the same.
struct BucketJob { const int32_t* key; int64_t n; int32_t nb; int64_t chunks; int32_t *ptr, *perm, *hist; };
__global__ void bucket_hist2_kernel(const BucketJob j0, const BucketJob j1) {
  const BucketJob& j = blockIdx.y ? j1 : j0;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < j.chunks) bucket_hist_chunk(j.key, j.n, j.nb, c, j.hist);
}
__global__ void __launch_bounds__(kScanThreads) bucket_scan2_kernel(const BucketJob j0, const BucketJob j1) {  // j0.nb + j1.nb blocks
  __shared__ int32_t part[kScanThreads];
  const bool second = (int32_t)blockIdx.x >= j0.nb;
  const BucketJob& j = second ? j1 : j0;
  const int32_t b = second ? (int32_t)blockIdx.x - j0.nb : (int32_t)blockIdx.x;
  const int t = threadIdx.x;
  const int64_t per = (j.chunks + kScanThreads - 1) / kScanThreads;
  const int64_t c0 = t * per < j.chunks ? t * per : j.chunks, c1 = c0 + per < j.chunks ? c0 + per : j.chunks;
  int32_t sum = 0;
  for (int64_t c = c0; c < c1; ++c) sum += j.hist[c * j.nb + b];
  part[t] = sum;
  __syncthreads();
  for (int off = 1; off < kScanThreads; off <<= 1) {
    const int32_t v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int32_t run = part[t] - sum;
  for (int64_t c = c0; c < c1; ++c) {
    const int32_t v = j.hist[c * j.nb + b];
    j.hist[c * j.nb + b] = run;
    run += v;
  }
  if (t == kScanThreads - 1) j.ptr[b + 1] = part[t];
}
__global__ void __launch_bounds__(kScanThreads) bucket_ptr2_kernel(const BucketJob j0, const BucketJob j1) {  // TWO blocks
  __shared__ int32_t sm[1024];
  __shared__ int32_t carry;
  const BucketJob& j = blockIdx.x ? j1 : j0;
  const int t = threadIdx.x;
  if (t == 0) { carry = 0; j.ptr[0] = 0; }
  __syncthreads();
  for (int32_t base = 0; base < j.nb; base += 1024) {
    const int32_t n = j.nb - base < 1024 ? j.nb - base : 1024;
    for (int32_t i = t; i < n; i += kScanThreads) sm[i] = j.ptr[base + 1 + i];
    __syncthreads();
    if (t == 0) {
      int32_t run = carry;
      for (int32_t i = 0; i < n; ++i) { run += sm[i]; sm[i] = run; }
      carry = run;
    }
    __syncthreads();
    for (int32_t i = t; i < n; i += kScanThreads) j.ptr[base + 1 + i] = sm[i];
    __syncthreads();
  }
}
__global__ void bucket_place2_kernel(const BucketJob j0, const BucketJob j1) {
  const BucketJob& j = blockIdx.y ? j1 : j0;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < j.chunks) bucket_place_chunk(j.key, j.n, j.nb, c, j.hist, j.ptr, j.perm);
}
static int bucket_device2(const BucketJob& j0, const BucketJob& j1, cudaStream_t s) {
  const int bt = 64;
  const int64_t cmax = j0.chunks > j1.chunks ? j0.chunks : j1.chunks;
  const dim3 grid((unsigned)((cmax + bt - 1) / bt), 2);
  if (cmax > 0) GCB_KERNEL("pack_bucket_hist2_kernel", s, bucket_hist2_kernel<<<grid, bt, 0, s>>>(j0, j1));
  if (j0.nb + j1.nb > 0)
    GCB_KERNEL("pack_bucket_scan2_kernel", s, bucket_scan2_kernel<<<(unsigned)(j0.nb + j1.nb), kScanThreads, 0, s>>>(j0, j1));
  GCB_KERNEL("pack_bucket_ptr2_kernel", s, bucket_ptr2_kernel<<<2, kScanThreads, 0, s>>>(j0, j1));
  if (cmax > 0) GCB_KERNEL("pack_bucket_place2_kernel", s, bucket_place2_kernel<<<grid, bt, 0, s>>>(j0, j1));
  return GCB_OK;
}

struct Scratch { int64_t pin, pout, pos, hist_a, hist_b, total; };
static Scratch scratch_layout(int32_t N, int32_t E, int32_t G, int32_t M, int32_t av, int32_t bv) {
  Scratch sc;
  int64_t off = 0;
  sc.pin = off; off += pad4((int64_t)N + G);
  sc.pout = off; off += pad4((int64_t)N + G);
  sc.pos = off; off += pad4(E);
  sc.hist_a = off; off += pad4(bucket_chunks(N) * av);
  sc.hist_b = off; off += pad4(bucket_chunks(M) * bv);
  sc.total = off > 0 ? off : 4;
  return sc;
}

static int check_raw_against_header(const int32_t* rh, const int32_t* h) {
  if (rh[RAW_MAGIC] != GCB_RAW_MAGIC || h[GCB_H_MAGIC] != GCB_PACK_MAGIC) return GCB_ERR_INVALID_ARG;
  if (rh[RAW_N] != h[GCB_H_N] || rh[RAW_E] != h[GCB_H_E] || rh[RAW_G] != h[GCB_H_G] || rh[RAW_M] != h[GCB_H_M] ||
      rh[RAW_AV] != h[GCB_H_AV] || rh[RAW_BV] != h[GCB_H_BV] || rh[RAW_NTILES] != h[GCB_H_NTILES] || h[GCB_H_FLAGS] != 0)
    return GCB_ERR_INVALID_ARG;
  return GCB_OK;
}

}  // namespace gcb

extern "C" int64_t gcb_raw_ints(int32_t N, int32_t E, int32_t G) {
  if (N < 0 || E < 0 || G < 0) return GCB_ERR_INVALID_ARG;
  const int64_t t = raw_layout(N, E, G, nullptr);
  return t < 0 ? GCB_ERR_INVALID_ARG : t;
}

extern "C" int gcb_collate_raw_host(const int32_t* const* atoms, const int32_t* n_atoms, const int32_t* const* bonds,
                                    const int64_t* const* conn, const int32_t* n_edges, const float* const* y,
                                    int32_t n_targets, int32_t n_graphs, int32_t n_messages, int32_t atom_vocab,
                                    int32_t bond_vocab, int32_t* raw, int64_t raw_ints, float* y_out, int32_t* packed_header) {
  if (n_graphs < 0 || !raw || !packed_header || (n_graphs > 0 && (!atoms || !n_atoms || !bonds || !conn || !n_edges)))
    return GCB_ERR_INVALID_ARG;
  int64_t N64 = 0, E64 = 0;
  int32_t max_a = 0;
  for (int32_t g = 0; g < n_graphs; ++g) {
    if (n_atoms[g] < 0 || n_edges[g] < 0) return GCB_ERR_INVALID_ARG;
    N64 += n_atoms[g];
    E64 += n_edges[g];
    max_a = n_atoms[g] > max_a ? n_atoms[g] : max_a;
  }
  if (N64 > INT32_MAX || E64 > INT32_MAX) return GCB_ERR_INVALID_ARG;
  const int32_t N = (int32_t)N64, E = (int32_t)E64, G = n_graphs;
  const int32_t M = n_messages >= 1 ? (N < E ? N : E) : E;
  if (layout(N, E, G, M, atom_vocab, bond_vocab, 0, packed_header) < 0) return GCB_ERR_INVALID_ARG;
  int32_t rh[GCB_RAW_HEADER_INTS];
  std::memset(rh, 0, sizeof(rh));
  const int64_t total = raw_layout(N, E, G, rh);
  if (total < 0) return GCB_ERR_INVALID_ARG;
  if (raw_ints < total) return GCB_ERR_WORKSPACE;
  int32_t* atom_off = raw + rh[RAW_ATOM_OFF];
  int32_t* edge_off = raw + rh[RAW_EDGE_OFF];
  int32_t* x_tok = raw + rh[RAW_X_TOK];
  int32_t* e_tok = raw + rh[RAW_E_TOK];
  int32_t* src = raw + rh[RAW_SRC];
  int32_t* dst = raw + rh[RAW_DST];
  int32_t* tile_ptr = raw + rh[RAW_TILE_PTR];
  const int64_t lim = n_messages >= 1 ? M : N;
  std::vector<int32_t> span((size_t)max_a + 2);
  std::vector<int32_t> tiles((size_t)N + 2);  // the walk may write up to N + 1 entries before it gives up
  TileWalker tw(tiles.data());
  int32_t node_off = 0, edge_off_ = 0;
  for (int32_t g = 0; g < G; ++g) {
    const int32_t na = n_atoms[g], ne = n_edges[g];
    atom_off[g] = node_off;
    edge_off[g] = edge_off_;
    for (int32_t i = 0; i < na; ++i) {
      const int32_t t = atoms[g][i];
      if (t < 0 || t >= atom_vocab) return GCB_ERR_INDEX;
      x_tok[node_off + i] = t;
    }
    for (int32_t i = 0; i <= na + 1; ++i) span[i] = 0;
    for (int32_t e = 0; e < ne; ++e) {
      const int32_t t = bonds[g][e];
      if (t < 0 || t >= bond_vocab) return GCB_ERR_INDEX;
      e_tok[edge_off_ + e] = t;
      const int64_t sl = conn[g][e], dl = conn[g][(int64_t)ne + e];  // [2,e] row-major
      if (sl < 0 || dl < 0 || sl + node_off >= lim || dl + node_off >= lim) return GCB_ERR_INDEX;  // PyG _lift IndexError
      if (sl >= na || dl >= na) return GCB_ERR_UNSUPPORTED;  // an edge into another molecule: the host packer's general path
      src[edge_off_ + e] = (int32_t)sl + node_off;
      dst[edge_off_ + e] = (int32_t)dl + node_off;
      const int32_t lo = (int32_t)(sl < dl ? sl : dl), hi = (int32_t)(sl < dl ? dl : sl);
      if (lo < hi) { span[lo + 1]++; span[hi + 1]--; }
    }
    int32_t cover = 0;  // cuts inside the molecule (it may be disconnected) and at its end, as in finish_pack_local
    for (int32_t p = 1; p <= na; ++p) {
      cover += span[p];
      if (p == na || cover == 0) tw.cut_at(node_off + p);
    }
    if (y_out) {
      for (int32_t t = 0; t < n_targets; ++t) y_out[(int64_t)g * n_targets + t] = y && y[g] ? y[g][t] : 0.f;
    }
    node_off += na;
    edge_off_ += ne;
  }
  atom_off[G] = N;
  edge_off[G] = E;
  const int32_t nt = tw.finish(N);
  if ((int64_t)nt + 1 > raw_tile_cap(N)) return GCB_ERR_INVALID_ARG;  // (cannot happen: two consecutive tiles span > GCB_TILE_ROWS rows)
  std::memcpy(tile_ptr, tiles.data(), sizeof(int32_t) * ((size_t)nt + 1));
  for (int64_t j = (int64_t)nt + 1; j < pad4(raw_tile_cap(N)); ++j) tile_ptr[j] = 0;
  {  // define every byte: alignment pads of the other sections
    const int64_t n[6] = {(int64_t)G + 1, (int64_t)G + 1, N, E, E, E};
    for (int i = 0; i < 6; ++i)
      for (int64_t j = n[i]; j < pad4(n[i]); ++j) raw[rh[RAW_ATOM_OFF + i] + j] = 0;
  }
  packed_header[GCB_H_NTILES] = nt;
  rh[RAW_MAGIC] = GCB_RAW_MAGIC;
  rh[RAW_N] = N; rh[RAW_E] = E; rh[RAW_G] = G; rh[RAW_M] = M; rh[RAW_AV] = atom_vocab; rh[RAW_BV] = bond_vocab;
  rh[RAW_NTILES] = nt;
  std::memcpy(raw, rh, sizeof(rh));
  return GCB_OK;
}

extern "C" int64_t gcb_pack_device_scratch_ints(const int32_t* packed_header) {
  if (!packed_header || packed_header[GCB_H_MAGIC] != GCB_PACK_MAGIC) return GCB_ERR_INVALID_ARG;
  const int32_t* h = packed_header;
  return scratch_layout(h[GCB_H_N], h[GCB_H_E], h[GCB_H_G], h[GCB_H_M], h[GCB_H_AV], h[GCB_H_BV]).total;
}

// ---- ELL-4 views on the device (gcb_pack_expand_ell) ------------------------------------------------------
// One thread per (row, view): the first four entries of the row's CSR range, -1 padded -- what finish_pack and
// collate_pack write on the host, so a batch that crossed PCIe without its derived tail equals one that carried it.
namespace gcb {
__global__ void expand_ell_kernel(const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_src, const int32_t* __restrict__ in_eid,
                                  const int32_t* __restrict__ out_ptr, const int32_t* __restrict__ out_dst,
                                  const int32_t* __restrict__ out_inpos, int4* __restrict__ s4, int4* __restrict__ e4,
                                  int4* __restrict__ d4, int4* __restrict__ o4, int32_t N) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 4 * (int64_t)N) return;
  const int32_t i = (int32_t)(t >> 2), view = (int32_t)(t & 3);
  const int32_t* ptr = view < 2 ? in_ptr : out_ptr;
  const int32_t* val = view == 0 ? in_src : view == 1 ? in_eid : view == 2 ? out_dst : out_inpos;
  int4* dst = view == 0 ? s4 : view == 1 ? e4 : view == 2 ? d4 : o4;
  const int32_t k0 = __ldg(ptr + i), k1 = __ldg(ptr + i + 1);
  int4 v;
  v.x = k0 + 0 < k1 ? __ldg(val + k0 + 0) : -1;
  v.y = k0 + 1 < k1 ? __ldg(val + k0 + 1) : -1;
  v.z = k0 + 2 < k1 ? __ldg(val + k0 + 2) : -1;
  v.w = k0 + 3 < k1 ? __ldg(val + k0 + 3) : -1;
  dst[i] = v;
}
}  // namespace gcb

extern "C" int gcb_pack_expand_ell(int32_t* packed_dev, const int32_t* h, void* stream) {
  using namespace gcb;
  if (!packed_dev || !h || h[GCB_H_MAGIC] != GCB_PACK_MAGIC) return GCB_ERR_INVALID_ARG;
  const int32_t N = h[GCB_H_N];
  if (N <= 0) return GCB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t threads = 4 * (int64_t)N;
  GCB_KERNEL("expand_ell_kernel", s, expand_ell_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(
      packed_dev + h[GCB_H_IN_PTR], packed_dev + h[GCB_H_IN_SRC], packed_dev + h[GCB_H_IN_EID], packed_dev + h[GCB_H_OUT_PTR],
      packed_dev + h[GCB_H_OUT_DST], packed_dev + h[GCB_H_OUT_INPOS], reinterpret_cast<int4*>(packed_dev + h[GCB_H_IN_SRC4]),
      reinterpret_cast<int4*>(packed_dev + h[GCB_H_IN_EID4]), reinterpret_cast<int4*>(packed_dev + h[GCB_H_OUT_DST4]),
      reinterpret_cast<int4*>(packed_dev + h[GCB_H_OUT_INPOS4]), N));
  return GCB_OK;
}

extern "C" int gcb_pack_device(const int32_t* raw_dev, const int32_t* raw_header_host, const int32_t* packed_header_host,
                               int32_t* packed_dev, int64_t packed_ints, int32_t* scratch_dev, int64_t scratch_ints,
                               void* stream) {
  if (!raw_dev || !raw_header_host || !packed_header_host || !packed_dev || !scratch_dev) return GCB_ERR_INVALID_ARG;
  const int32_t* h = packed_header_host;
  const int32_t* rh = raw_header_host;
  GCB_TRY(check_raw_against_header(rh, h));
  const int32_t N = h[GCB_H_N], E = h[GCB_H_E], G = h[GCB_H_G], M = h[GCB_H_M], av = h[GCB_H_AV], bv = h[GCB_H_BV];
  const Scratch sc = scratch_layout(N, E, G, M, av, bv);
  if (packed_ints < h[GCB_H_TOTAL_INTS] || scratch_ints < sc.total) return GCB_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  PackSegs sg;
  GCB_TRY(build_segs(h, rh, &sg));
  int32_t maxc = 1;
  for (int i = 0; i < sg.n; ++i) maxc = sg.count[i] > maxc ? sg.count[i] : maxc;
  int gx = (maxc + 256 * 4 - 1) / (256 * 4);
  gx = gx < 1 ? 1 : (gx > 4 * kNumSMs ? 4 * kNumSMs : gx);
  GCB_KERNEL("pack_segments_kernel", s, pack_segments_kernel<<<dim3((unsigned)gx, (unsigned)sg.n + 1), 256, 0, s>>>(sg, raw_dev, packed_dev));
  // warp-per-molecule kernel + token sorts that share their launches: 0.169 vs 0.269 ms per 8192-molecule batch on B200
  // (profiles/r2_device_pack_time.json); GCB_PACK_WARP=0 selects the thread-per-molecule kernels
  static const bool warp_path = [] { const char* e = getenv("GCB_PACK_WARP"); return !(e && e[0] == '0'); }();
  if (G > 0) {
    const PackPtrs p = pack_ptrs(packed_dev, h);
    if (warp_path) {
      GCB_KERNEL("pack_molecules_warp_kernel", s,
                 pack_molecules_warp_kernel<<<(unsigned)((G + kPackWarps - 1) / kPackWarps), kPackWarps * 32, 0, s>>>(
                     p, raw_dev + rh[RAW_ATOM_OFF], raw_dev + rh[RAW_EDGE_OFF], G, scratch_dev + sc.pin, scratch_dev + sc.pout,
                     scratch_dev + sc.pos));
    } else {
      GCB_KERNEL("pack_molecules_kernel", s,
                 pack_molecules_kernel<<<(unsigned)((G + 31) / 32), 32, 0, s>>>(p, raw_dev + rh[RAW_ATOM_OFF], raw_dev + rh[RAW_EDGE_OFF], G,
                                                                               scratch_dev + sc.pin, scratch_dev + sc.pout, scratch_dev + sc.pos));
    }
  }
  // token buckets read the packed token sections (written by the segment copies, same stream)
  if (warp_path) {
    const BucketJob ja{packed_dev + h[GCB_H_X_TOK], N, av, bucket_chunks(N), packed_dev + h[GCB_H_ATOK_PTR],
                       packed_dev + h[GCB_H_ATOK_NODE], scratch_dev + sc.hist_a};
    const BucketJob jb{packed_dev + h[GCB_H_E_TOK], M, bv, bucket_chunks(M), packed_dev + h[GCB_H_BTOK_PTR],
                       packed_dev + h[GCB_H_BTOK_ROW], scratch_dev + sc.hist_b};
    GCB_TRY(bucket_device2(ja, jb, s));
    return GCB_OK;
  }
  GCB_TRY(bucket_device(packed_dev + h[GCB_H_X_TOK], N, av, packed_dev + h[GCB_H_ATOK_PTR], packed_dev + h[GCB_H_ATOK_NODE],
                        scratch_dev + sc.hist_a, s));
  GCB_TRY(bucket_device(packed_dev + h[GCB_H_E_TOK], M, bv, packed_dev + h[GCB_H_BTOK_PTR], packed_dev + h[GCB_H_BTOK_ROW],
                        scratch_dev + sc.hist_b, s));
  return GCB_OK;
}

// Test hook: the routines of the kernels above, run serially on the CPU over host buffers (raw, packed and scratch
// in host memory).  Pins the device packer's index logic against gcb_collate_pack_host without a GPU.
extern "C" int gcb_debug_pack_emulate_host(const int32_t* raw, const int32_t* packed_header, int32_t* packed,
                                           int64_t packed_ints, int32_t* scratch, int64_t scratch_ints, int32_t force_scratch) {
  if (!raw || !packed_header || !packed || !scratch) return GCB_ERR_INVALID_ARG;
  const int32_t* h = packed_header;
  const int32_t* rh = raw;
  GCB_TRY(check_raw_against_header(rh, h));
  const int32_t N = h[GCB_H_N], E = h[GCB_H_E], G = h[GCB_H_G], M = h[GCB_H_M], av = h[GCB_H_AV], bv = h[GCB_H_BV];
  const Scratch sc = scratch_layout(N, E, G, M, av, bv);
  if (packed_ints < h[GCB_H_TOTAL_INTS] || scratch_ints < sc.total) return GCB_ERR_WORKSPACE;
  PackSegs sg;
  GCB_TRY(build_segs(h, rh, &sg));
  for (int s = 0; s < sg.n; ++s)
    for (int64_t i = 0; i < sg.count[s]; ++i) seg_element(sg, s, i, raw, packed);
  seg_fixed(sg, packed);
  const PackPtrs p = pack_ptrs(packed, h);
  const int32_t* atom_off = raw + rh[RAW_ATOM_OFF];
  const int32_t* edge_off = raw + rh[RAW_EDGE_OFF];
  for (int32_t g = 0; g < G; ++g) {
    const int32_t a0 = atom_off[g], na = atom_off[g + 1] - a0, e0 = edge_off[g], ne = edge_off[g + 1] - e0;
    if (force_scratch == 2 && na <= kPackMaxA && ne <= kPackMaxE) {  // the warp-per-molecule phases, lane by lane
      int32_t w[kWarpWork];
      for (int i = 0; i < kWarpWork; ++i) w[i] = -9;
      for (int phase = 0; phase < kWarpPhases; ++phase)
        for (int lane = 0; lane < 32; ++lane) pack_warp_phase(phase, lane, p, g, a0, na, e0, ne, w);
      for (int lane = 0; lane < 32; ++lane) pack_warp_ell(lane, p, a0, na, e0, w);
    } else if (!force_scratch && na <= kPackMaxA && ne <= kPackMaxE) {
      int32_t pin[kPackMaxA + 1], pout[kPackMaxA + 1], pos[kPackMaxE];
      pack_one_molecule(p, g, a0, na, e0, ne, pin, pout, pos);
    } else {
      pack_one_molecule(p, g, a0, na, e0, ne, scratch + sc.pin + a0 + g, scratch + sc.pout + a0 + g, scratch + sc.pos + e0);
    }
  }
  auto bucket_emulated = [&](const int32_t* key, int64_t n, int32_t nb, int32_t* ptr, int32_t* perm, int32_t* hist) {
    const int64_t chunks = bucket_chunks(n);
    for (int64_t c = 0; c < chunks; ++c) bucket_hist_chunk(key, n, nb, c, hist);
    for (int32_t b = 0; b < nb; ++b) bucket_scan_bucket(nb, chunks, b, hist, ptr);
    bucket_ptr_scan(nb, ptr);
    for (int64_t c = 0; c < chunks; ++c) bucket_place_chunk(key, n, nb, c, hist, ptr, perm);
  };
  bucket_emulated(packed + h[GCB_H_X_TOK], N, av, packed + h[GCB_H_ATOK_PTR], packed + h[GCB_H_ATOK_NODE], scratch + sc.hist_a);
  bucket_emulated(packed + h[GCB_H_E_TOK], M, bv, packed + h[GCB_H_BTOK_PTR], packed + h[GCB_H_BTOK_ROW], scratch + sc.hist_b);
  return GCB_OK;
}
