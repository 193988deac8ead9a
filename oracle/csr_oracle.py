"""TEST INFRASTRUCTURE ONLY -- integer restatement (numpy) of the index structures the CUDA path
consumes.  Everything here must match the product's packer BIT FOR BIT.

The reference has no CSR: PyG scatters in edge-id order (`scatter_add_` / `scatter_reduce_` over
`edge_index[1]`).  A destination-sorted CSR that keeps edge-id order inside each
row is the *stable* argsort of `edge_index[1]`; the source-sorted CSR (used by the backward pass)
is the stable argsort of `edge_index[0]`; graph segments are the stable argsort of `batch`
(`global_add_pool`, gcn.py:181); token buckets are the stable argsort of the token ids
(`embedding_dense_backward` sums rows of equal token in row order).
"""
from __future__ import annotations

import numpy as np


def stable_buckets(keys: np.ndarray, n_buckets: int):
    """Return (ptr[n_buckets+1], perm) with perm = argsort(keys, stable)."""
    keys = np.asarray(keys, dtype=np.int64)
    perm = np.argsort(keys, kind="stable").astype(np.int32)
    counts = np.bincount(keys, minlength=n_buckets)[:n_buckets] if keys.size else np.zeros(n_buckets, np.int64)
    ptr = np.zeros(n_buckets + 1, dtype=np.int32)
    ptr[1:] = np.cumsum(counts)
    return ptr, perm


def closed_tiles(src, dst, n_rows, tile_rows=128):
    """Greedy partition of rows [0, n_rows) into tiles of at most `tile_rows` rows that no edge leaves
    (a cut at p is legal iff no edge has min(src,dst) < p <= max(src,dst)).  Returns tile_ptr, or [0]
    when some connected run of rows is wider than a tile."""
    lo, hi = np.minimum(src, dst).astype(np.int64), np.maximum(src, dst).astype(np.int64)
    span = np.zeros(n_rows + 2, dtype=np.int64)
    np.add.at(span, lo[lo < hi] + 1, 1)
    np.add.at(span, hi[lo < hi] + 1, -1)
    cover = np.cumsum(span)
    cuts = [p for p in range(1, n_rows + 1) if p == n_rows or cover[p] == 0]
    ptr, start, last = [0], 0, 0
    for p in cuts:
        if p - start > tile_rows:
            if last == start:
                return np.zeros(1, dtype=np.int32)
            ptr.append(last)
            start = last
            if p - start > tile_rows:
                return np.zeros(1, dtype=np.int32)
        last = p
    if n_rows > 0:
        ptr.append(n_rows)
    return np.asarray(ptr, dtype=np.int32)


def build_index(x_tok, e_tok, edge_index, batch, n_graphs, atom_vocab, bond_vocab, live_rows):
    """All index arrays of a packed batch, as a dict of int32 numpy arrays."""
    x_tok = np.asarray(x_tok).astype(np.int32)
    e_tok = np.asarray(e_tok).astype(np.int32)
    src = np.asarray(edge_index[0]).astype(np.int32)
    dst = np.asarray(edge_index[1]).astype(np.int32)
    N, E = x_tok.shape[0], e_tok.shape[0]
    in_ptr, in_eid = stable_buckets(dst, N)
    out_ptr, out_eid = stable_buckets(src, N)
    if batch is None:
        batch = np.zeros(N, dtype=np.int64)
    graph_ptr, graph_node = stable_buckets(np.asarray(batch), n_graphs)
    atok_ptr, atok_node = stable_buckets(x_tok, atom_vocab)
    btok_ptr, btok_row = stable_buckets(e_tok[:live_rows], bond_vocab)
    def ell4(ptr, vals):
        out = np.full((N, 4), -1, dtype=np.int32)
        for u in range(4):
            k = ptr[:-1].astype(np.int64) + u
            ok = k < ptr[1:]
            out[ok, u] = vals[k[ok]]
        return out.reshape(-1)
    pos_of_eid = np.empty(E, dtype=np.int32)
    pos_of_eid[in_eid] = np.arange(E, dtype=np.int32)
    out_inpos = pos_of_eid[out_eid] if E else out_eid
    in_src_sorted = src[in_eid] if E else in_eid
    out_dst_sorted = dst[out_eid] if E else out_eid
    return dict(
        in_src4=ell4(in_ptr, in_src_sorted), in_eid4=ell4(in_ptr, in_eid), out_dst4=ell4(out_ptr, out_dst_sorted), out_inpos4=ell4(out_ptr, out_inpos), out_inpos=out_inpos,
        x_tok=x_tok, e_tok=e_tok, src=src, dst=dst,
        in_ptr=in_ptr, in_eid=in_eid, in_src=src[in_eid] if E else in_eid,
        out_ptr=out_ptr, out_eid=out_eid, out_dst=dst[out_eid] if E else out_eid,
        graph_ptr=graph_ptr, graph_node=graph_node,
        atok_ptr=atok_ptr, atok_node=atok_node, btok_ptr=btok_ptr, btok_row=btok_row,
        node_graph=np.asarray(batch).astype(np.int32),
        tile_ptr=closed_tiles(src, dst, N),
    )


def each_layout(n_atoms, n_edges, bond_tok_per_graph):
    """GCB_PACK_EACH restated: molecule m with node offset a_m and edge offset e_m keeps, alone, the bond
    matrix rows 0..n_edges[m]-1 indexed by its local atom ids (reference gcn.py:163).  In the batch the
    bond row of global atom a_m + i is the molecule's edge i, so
        bond_pos[e_m + k] = a_m + k if k < n_atoms[m] else -1      (row that holds edge e_m + k)
        e_tok_each[a_m + i] = bond token of edge e_m + i if i < n_edges[m] else 0
    Returns (bond_pos [E], e_tok_each [N])."""
    n_atoms, n_edges = np.asarray(n_atoms), np.asarray(n_edges)
    a_off = np.concatenate([[0], np.cumsum(n_atoms)])
    N, E = int(a_off[-1]), int(n_edges.sum())
    bond_pos = np.full(E, -1, dtype=np.int32)
    e_tok = np.zeros(N, dtype=np.int32)
    e = 0
    for m, (na, ne) in enumerate(zip(n_atoms, n_edges)):
        for k in range(ne):
            if k < na:
                bond_pos[e + k] = a_off[m] + k
                e_tok[a_off[m] + k] = bond_tok_per_graph[m][k]
        e += ne
    return bond_pos, e_tok
