"""Synthetic ZINC-scale molecule graphs in MoleculeEncoder layout (BASELINE.json configs 3-5;
SURVEY.md 8d): every molecule is a random spanning tree plus ring-closure bonds, edges are emitted
source-atom-major with both directions of a bond sharing one token -- the order
`MoleculeEncoder.encode` produces (reference graphchem/preprocessing/features.py:299-318).

  fixed shape   : 23 atoms, 22 tree bonds + 3 ring closures -> 50 directed edges
  variable shape: n_atoms ~ clip(round(N(23, 4.5)), 6, 38), rings ~ Binomial(4, 0.6)
  tokens        : atoms uniform in [2, atom_vocab), bonds uniform in [2, bond_vocab); targets N(0,1)
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .data.packed import MoleculeStore


def _gen_class(rng: np.random.Generator, G: int, n_atoms: int, n_rings: int, atom_vocab: int, bond_vocab: int):
    """G molecules with identical (n_atoms, n_rings): returns atoms [G,n], bonds [G,2b], conn [G,2,2b]."""
    n, nb = n_atoms, n_atoms - 1 + n_rings
    a = np.zeros((G, nb), dtype=np.int64)
    b = np.zeros((G, nb), dtype=np.int64)
    adj = np.zeros((G, n, n), dtype=bool)
    rows = np.arange(G)
    for i in range(1, n):  # spanning tree: atom i bonds to a random earlier atom
        par = rng.integers(0, i, size=G)
        a[:, i - 1], b[:, i - 1] = par, i
        adj[rows, par, i] = adj[rows, i, par] = True
    for r in range(n_rings):  # ring closures between distinct, non-adjacent atoms
        todo = rows.copy()
        k = n - 1 + r
        while todo.size:
            u = rng.integers(0, n, size=todo.size)
            v = rng.integers(0, n, size=todo.size)
            ok = (u != v) & ~adj[todo, u, v]
            lo, hi = np.minimum(u, v), np.maximum(u, v)
            a[todo[ok], k], b[todo[ok], k] = lo[ok], hi[ok]
            adj[todo[ok], lo[ok], hi[ok]] = adj[todo[ok], hi[ok], lo[ok]] = True
            todo = todo[~ok]
    btok = rng.integers(2, bond_vocab, size=(G, nb))
    bidx = np.broadcast_to(np.arange(nb), (G, nb))
    src = np.concatenate([a, b], axis=1)
    dst = np.concatenate([b, a], axis=1)
    tok = np.concatenate([btok, btok], axis=1)
    key = src * (2 * nb) + np.concatenate([bidx, bidx], axis=1)  # atom-major, then bond creation order
    order = np.argsort(key, axis=1, kind="stable")
    src = np.take_along_axis(src, order, 1)
    dst = np.take_along_axis(dst, order, 1)
    tok = np.take_along_axis(tok, order, 1)
    atoms = rng.integers(2, atom_vocab, size=(G, n))
    conn = np.stack([src, dst], axis=1)  # [G, 2, 2nb]
    return atoms.astype(np.int32), tok.astype(np.int32), conn.astype(np.int64)


def make_molecules(n_molecules: int, seed: int = 1234, atom_vocab: int = 64, bond_vocab: int = 32,
                   variable: bool = False) -> MoleculeStore:
    """Generate `n_molecules` synthetic molecules as a flat host `MoleculeStore`."""
    rng = np.random.default_rng(seed)
    if not variable:
        # flat arrays, generated in chunks (the [G, 23, 23] adjacency scratch of a 10 M-molecule shard would not fit)
        chunk = 1 << 18
        parts = [_gen_class(rng, min(chunk, n_molecules - lo), 23, 3, atom_vocab, bond_vocab)
                 for lo in range(0, n_molecules, chunk)]
        targets = torch.from_numpy(rng.standard_normal((n_molecules, 1)).astype(np.float32))
        cat = lambda k: (np.concatenate([p[k] for p in parts]) if len(parts) > 1 else parts[0][k]) if parts else None  # noqa: E731
        if n_molecules == 0:
            return MoleculeStore([], [], [], targets)
        return MoleculeStore.from_flat(torch.from_numpy(cat(0).reshape(-1)), torch.from_numpy(cat(1).reshape(-1)),
                                       torch.from_numpy(cat(2).reshape(-1)),
                                       np.full(n_molecules, 23, dtype=np.int32), np.full(n_molecules, 50, dtype=np.int32),
                                       targets)
    else:
        na = np.clip(np.rint(rng.normal(23, 4.5, size=n_molecules)), 6, 38).astype(np.int64)
        nr = rng.binomial(4, 0.6, size=n_molecules)
        atoms_l, bonds_l, conn_l = [None] * n_molecules, [None] * n_molecules, [None] * n_molecules
        for key in np.unique(na * 8 + nr):
            sel = np.nonzero(na * 8 + nr == key)[0]
            atoms, bonds, conn = _gen_class(rng, sel.size, int(key // 8), int(key % 8), atom_vocab, bond_vocab)
            for j, m in enumerate(sel):
                atoms_l[m], bonds_l[m], conn_l[m] = (torch.from_numpy(atoms[j]), torch.from_numpy(bonds[j]),
                                                     torch.from_numpy(conn[j]))
    targets = torch.from_numpy(rng.standard_normal((n_molecules, 1)).astype(np.float32))
    return MoleculeStore(atoms_l, bonds_l, conn_l, targets)


def fixed_shape_counts(n_molecules: int) -> Tuple[int, int]:
    """(N, E) of a fixed-shape batch."""
    return 23 * n_molecules, 50 * n_molecules
