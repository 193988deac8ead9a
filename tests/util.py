"""Shared helpers for the parity tests: random PyG-style batches, oracle/CUDA model pairs."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from oracle.gcn_oracle import OracleMoleculeGCN
from oracle.pyg_restated import Data as OData, collate as ocollate


def random_molecule(rng: np.random.Generator, n_atoms: int, n_rings: int, av: int, bv: int):
    """One molecule in MoleculeEncoder layout (atom-major directed edges, shared bond token)."""
    a, b = [], []
    adj = set()
    for i in range(1, n_atoms):
        p = int(rng.integers(0, i))
        a.append(p); b.append(i); adj.add((p, i))
    tries = 0
    while n_rings > 0 and tries < 100 and n_atoms > 2:
        u, v = sorted(int(t) for t in rng.integers(0, n_atoms, size=2))
        tries += 1
        if u != v and (u, v) not in adj:
            a.append(u); b.append(v); adj.add((u, v)); n_rings -= 1
    nb = len(a)
    btok = rng.integers(2, bv, size=nb)
    edges = []
    for k in range(nb):
        edges.append((a[k], k, b[k], btok[k]))
        edges.append((b[k], k, a[k], btok[k]))
    edges.sort(key=lambda t: (t[0], t[1]))
    x = torch.from_numpy(rng.integers(2, av, size=n_atoms).astype(np.int32))
    ea = torch.tensor([int(e[3]) for e in edges], dtype=torch.int32)
    ei = torch.tensor([[e[0] for e in edges], [e[2] for e in edges]], dtype=torch.int64).reshape(2, len(edges))
    y = torch.from_numpy(rng.standard_normal((1, 1)).astype(np.float32))
    return OData(x=x, edge_index=ei, edge_attr=ea, y=y)


def random_batch(seed: int, n_graphs: int, av: int = 12, bv: int = 7, min_atoms: int = 2, max_atoms: int = 12):
    rng = np.random.default_rng(seed)
    graphs = [random_molecule(rng, int(rng.integers(min_atoms, max_atoms + 1)), int(rng.integers(0, 3)), av, bv)
              for _ in range(n_graphs)]
    return ocollate(graphs), graphs


def random_graph_batch(seed: int, N: int, E: int, av: int, bv: int, G: int = 1):
    """Arbitrary (non-molecular) graph as the reference's own tests use (tests/test_gcn.py:51-56):
    random edge_index with duplicate edges and isolated nodes; requires E >= N."""
    g = torch.Generator().manual_seed(seed)
    d = OData(x=torch.randint(0, av, (N,), generator=g), edge_index=torch.randint(0, N, (2, E), generator=g),
              edge_attr=torch.randint(0, bv, (E,), generator=g))
    d.batch = torch.sort(torch.randint(0, G, (N,), generator=g)).values if G > 1 else torch.zeros(N, dtype=torch.long)
    d.y = torch.randn(G, 1, generator=g)
    d.num_graphs = G
    return d


def make_pair(av, bv, out_dim, seed=0, dtype=torch.float32, device="cuda", **kw):
    """(oracle on CPU in fp32, CUDA model) sharing one state_dict."""
    from graphchem_b200.nn import MoleculeGCN
    torch.manual_seed(seed)
    oracle = OracleMoleculeGCN(av, bv, out_dim, **kw)
    model = MoleculeGCN(av, bv, out_dim, compute_dtype=dtype, **kw)
    model.load_state_dict(oracle.state_dict())
    model = model.to(device)
    return oracle, model


def oracle_fp64(oracle: OracleMoleculeGCN) -> OracleMoleculeGCN:
    """The same restatement evaluated in float64: its rounding noise (~1e-16) is far below every
    tolerance, so |CUDA - oracle64| is the CUDA path's OWN error, and |oracle32 - oracle64| is the rounding noise of
    the reference's fp32 arithmetic.  Comparing against it makes the parity tests independent of how torch's CPU
    kernels happen to block their fp32 sums on a given box / thread count."""
    import copy
    o = copy.deepcopy(oracle).double()
    o.train(oracle.training)
    return o


def data_fp64(data):
    """Shallow copy of a batch with float64 targets (integer tensors are shared)."""
    import copy
    d = copy.copy(data)
    if getattr(data, "y", None) is not None:
        d.y = data.y.double()
    return d


def nerr(a: torch.Tensor, b: torch.Tensor) -> float:
    """Norm-wise relative error max|a-b| / max|b| (b = oracle)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if a.shape != b.shape:
        return float("inf")
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def load_fuel_set(name: str):
    """(smiles, targets [n,1] fp32, train_idx, test_idx, notebook vocab sizes) of one bundled property data set
    from the committed fixture tests/golden/fuel_sets.json (made by tests/golden/make_fuel_sets.py)."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fuel_sets.json")
    with open(path) as f:
        d = json.load(f)[name]
    y = torch.tensor(d["target"], dtype=torch.float32).reshape(-1, 1)
    return d["smiles"], y, d["train_idx"], d["test_idx"], tuple(d["notebook_vocab_sizes"])


def encoded_fuel_set(name: str):
    """The notebook pipeline on one data set with the built-in perception (predict_cn.ipynb cells 0-1):
    encoder on the training split, MoleculeGraph lists for train and test."""
    from graphchem_b200.data import MoleculeGraph
    from graphchem_b200.preprocessing import MoleculeEncoder
    smiles, y, tr, te, _ = load_fuel_set(name)
    enc = MoleculeEncoder([smiles[i] for i in tr], backend="builtin")
    mk = lambda idx: [MoleculeGraph(*enc.encode(smiles[i]), y[i]) for i in idx]  # noqa: E731
    return enc, mk(tr), mk(te)


def assert_adam_states_close(got: torch.Tensor, want: torch.Tensor, lr: float, steps: int, name: str = "") -> None:
    """Parameters after a few Adam steps, two implementations.  Adam's first update is lr * sign(g): an entry whose
    gradient is zero to within the rounding noise of the two backward passes can move by +lr in one and -lr in the
    other, so a few entries may differ by up to 2 * lr * steps; everything else must agree to a fraction of lr."""
    d = (got.detach().double().cpu() - want.detach().double().cpu()).abs()
    assert d.shape == want.shape, name
    if d.numel() == 0:
        return
    assert float(d.max()) <= 2.02 * lr * steps + 1e-6, (name, float(d.max()))
    outliers = int((d > 0.05 * lr).sum())
    assert outliers <= max(1, int(2e-3 * d.numel() + 0.5)), (name, outliers, d.numel())
