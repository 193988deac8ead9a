"""TEST INFRASTRUCTURE ONLY -- CPU restatement of graphchem.nn.MoleculeGCN (reference
`graphchem/nn/gcn.py:10-202`) on top of `oracle/pyg_restated.py`.

PARITY UNPINNED for floating-point values (see pyg_restated.py header): the reference's tests
pin shapes only.  The integer side (token numbering, edge order, collate offsets) is pinned by
the notebook golden encodings committed under tests/golden/.

The forward below keeps the reference's LITERAL op order (index_select / cat / addmm /
scatter_reduce_(amax, include_self=False) / scatter_add_), so its rounding is the reference's
rounding, not that of the re-associated CUDA kernels.  Gradients come from torch autograd,
exactly as `loss.backward()` in `examples/predict_cn.ipynb:250`.

`state_dict()` keys and shapes equal the reference's (`gcn.py:78-114`), so weights are shared
with the CUDA model through `load_state_dict`.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .pyg_restated import EdgeConv, GeneralConv, global_add_pool, global_mean_pool


class OracleMoleculeGCN(nn.Module):
    # gcn.py:35-118
    def __init__(self, atom_vocab_size: int, bond_vocab_size: int, output_dim: Optional[int],
                 embedding_dim: int = 128, n_messages: int = 2, n_readout: int = 2,
                 readout_dim: int = 64, p_dropout: float = 0.0, aggr: str = "add",
                 act_fn: Callable = F.softplus, *, pool: str = "add"):
        super().__init__()
        self._pool = pool                    # not in the reference (always global_add_pool); checks the product's `pool="mean"`
        self._p_dropout = p_dropout          # gcn.py:73
        self._n_messages = n_messages        # gcn.py:74
        self.act_fn = act_fn                 # gcn.py:75
        self.emb_atom = nn.Embedding(atom_vocab_size, embedding_dim)           # gcn.py:78
        self.emb_bond = nn.Embedding(bond_vocab_size, embedding_dim)           # gcn.py:81
        self.atom_conv = GeneralConv(embedding_dim, embedding_dim, embedding_dim, aggr=aggr)  # :84-86
        self.bond_conv = EdgeConv(nn.Sequential(nn.Linear(2 * embedding_dim, embedding_dim)))  # :89-91
        if n_readout > 0:                    # gcn.py:94-114
            layers = [nn.Sequential(nn.Linear(embedding_dim, readout_dim))]
            layers += [nn.Sequential(nn.Linear(readout_dim, readout_dim)) for _ in range(n_readout - 1)]
            layers += [nn.Sequential(nn.Linear(readout_dim, output_dim))]
            self.readout = nn.ModuleList(layers)
        else:                                # gcn.py:116-118
            self.readout = None

    # gcn.py:120-202
    def forward(self, data, return_layers: bool = False) -> Tuple[torch.Tensor, ...]:
        x, edge_attr, edge_index, batch = data.x, data.edge_attr, data.edge_index, data.batch  # :144
        if data.num_node_features == 0:      # gcn.py:148-149 (dead branch in practice)
            x = torch.ones(data.num_nodes, 1)
        out_atom = self.act_fn(self.emb_atom(x))              # gcn.py:152-153
        out_bond = self.act_fn(self.emb_bond(edge_attr))      # gcn.py:156-157
        layers = [(out_atom, out_bond)]
        for _ in range(self._n_messages):                     # gcn.py:160
            out_bond = self.act_fn(self.bond_conv(out_bond, edge_index))               # :163-164
            out_bond = F.dropout(out_bond, p=self._p_dropout, training=self.training)  # :167-169
            out_atom = self.act_fn(self.atom_conv(out_atom, edge_index, out_bond))     # :172-173
            out_atom = F.dropout(out_atom, p=self._p_dropout, training=self.training)  # :176-178
            layers.append((out_atom, out_bond))
        out = (global_mean_pool if self._pool == "mean" else global_add_pool)(out_atom, batch)  # gcn.py:181
        if self.readout is not None:                          # gcn.py:184
            for layer in self.readout[:-1]:                   # gcn.py:187-196
                out = self.act_fn(layer(out))
                out = F.dropout(out, p=self._p_dropout, training=self.training)
            out = self.readout[-1](out)                       # gcn.py:199
        if return_layers:
            return out, out_atom, out_bond, layers
        return out, out_atom, out_bond                        # gcn.py:202


def train_step(model: OracleMoleculeGCN, optimizer: torch.optim.Optimizer, batch) -> torch.Tensor:
    """One step of the notebook loop, `examples/predict_cn.ipynb:247-251`."""
    optimizer.zero_grad()
    pred, _, _ = model(batch)
    loss = F.mse_loss(pred, batch.y)
    loss.backward()
    optimizer.step()
    return loss.detach()
