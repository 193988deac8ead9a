"""Generates tests/golden/gcn_small.pt: a seeded 24-molecule batch (embedding_dim 64), a seeded MoleculeGCN state_dict and what
the CPU oracle (oracle/gcn_oracle.py, the restated reference) computes for it in fp32 -- forward outputs,
MSE loss, every parameter gradient, and the parameters after three Adam steps.

The real reference cannot run here (torch_geometric / rdkit are not installable offline), so
these vectors pin the ORACLE against drift (torch upgrades, refactors), and the CUDA path against a committed
artefact instead of only against a live oracle run.  Regenerate with:  python tests/golden/make_gcn_golden.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F

from oracle.gcn_oracle import OracleMoleculeGCN, train_step
from tests.util import random_batch

AV, BV, KW = 12, 7, dict(embedding_dim=64, n_messages=3, n_readout=2, readout_dim=32)  # small: the file is committed
# second file (round 2): the non-default aggregation and pooling -- GeneralConv(aggr="max"), global_mean_pool
VARIANTS = {"gcn_small.pt": {}, "gcn_small_max_meanpool.pt": dict(aggr="max", pool="mean")}


def build(extra=None):
    extra = dict(extra or {})
    data, _ = random_batch(seed=2024, n_graphs=24, av=AV, bv=BV, max_atoms=14)
    torch.manual_seed(2024)
    model = OracleMoleculeGCN(AV, BV, 2, **KW, **extra)
    state = {k: v.clone() for k, v in model.state_dict().items()}
    model.train()
    out, out_atom, out_bond = model(data)
    g = torch.Generator().manual_seed(7)
    y = torch.randn(out.shape, generator=g)
    loss = F.mse_loss(out, y)
    model.zero_grad()
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    data.y = y
    losses = [float(train_step(model, opt, data)) for _ in range(3)]
    after = {k: v.clone() for k, v in model.state_dict().items()}
    return dict(x=data.x, edge_index=data.edge_index, edge_attr=data.edge_attr, batch=data.batch, y=y, state=state,
                out=out.detach(), out_atom=out_atom.detach(), out_bond=out_bond.detach(), loss=float(loss.detach()), grads=grads,
                adam_losses=losses, state_after_3_adam_steps=after, dims=dict(av=AV, bv=BV, out=2, **KW), extra=extra)


if __name__ == "__main__":
    only = sys.argv[1:]  # file names to (re)generate; default: those that do not exist yet (committed files stay as they are)
    for name, extra in VARIANTS.items():
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), name)
        if (only and name not in only) or (not only and os.path.exists(path)):
            continue
        torch.save(build(extra), path)
        print("wrote", path, os.path.getsize(path), "bytes")
