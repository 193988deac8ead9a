"""Small-batch regime of BASELINE.json configs[0]/[1]: the notebook training loop of the reference
(examples/predict_cn.ipynb:238-272: DataLoader(batch_size=16), model(batch), F.mse_loss, backward,
Adam.step) on a CN-sized synthetic data set (368 molecules, ~13 atoms each, fp32, D=128), run

  (a) verbatim against graphchem_b200.nn.MoleculeGCN (autograd.Function + torch.optim.Adam, CPU batches in),
  (b) through graphchem_b200.train.TrainStep on pre-packed device batches (CUDA graph per batch),
  (c) on the CPU oracle (the reference restated), all host threads.

Prints one JSON line with molecules/s of each.  RDKit is absent, so the molecules are synthetic; the
loop, model dims and batch size are the notebook's."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.nn.functional as F

from graphchem_b200.data import DataLoader, MoleculeDataset, MoleculeGraph, MoleculeStore
from graphchem_b200.nn import MoleculeGCN
from graphchem_b200.train import TrainStep
from oracle.gcn_oracle import OracleMoleculeGCN
from oracle.pyg_restated import collate
from tests.util import random_batch

AV, BV, B, EPOCHS = 42, 29, 16, 3
_, graphs = random_batch(seed=1, n_graphs=368, av=AV, bv=BV, min_atoms=5, max_atoms=22)
ds = MoleculeDataset([MoleculeGraph(g.x, g.edge_attr, g.edge_index, g.y) for g in graphs])
kw = dict(embedding_dim=128, n_messages=2, n_readout=2, readout_dim=64)
res = {"molecules": len(ds), "batch": B, "epochs_timed": EPOCHS}

# (a) notebook loop, verbatim
torch.manual_seed(0)
model = MoleculeGCN(AV, BV, 1, **kw).cuda()
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
loader = DataLoader(ds, batch_size=B, shuffle=True)


def epoch_a():
    tot = 0.0
    for batch in loader:
        opt.zero_grad()
        pred, _, _ = model(batch)
        loss = F.mse_loss(pred, batch.y)
        loss.backward()
        opt.step()
        tot += loss.detach().item()
    return tot


epoch_a()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(EPOCHS):
    epoch_a()
torch.cuda.synchronize()
res["dropin_autograd_mol_per_s"] = EPOCHS * len(ds) / (time.perf_counter() - t0)

# (b) TrainStep on pre-packed device batches
store = MoleculeStore.from_graphs(list(ds))
packed = [store.collate_pack(np.arange(i, min(i + B, len(ds))), 2, AV, BV).to("cuda") for i in range(0, len(ds), B)]
step = TrainStep(model, lr=1e-3, capture_after=1)
for pb in packed:
    step(pb)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(EPOCHS):
    for pb in packed:
        step(pb)
torch.cuda.synchronize()
res["trainstep_graph_mol_per_s"] = EPOCHS * len(ds) / (time.perf_counter() - t0)

# (c) CPU oracle
torch.set_num_threads(os.cpu_count() or 1)
oracle = OracleMoleculeGCN(AV, BV, 1, **kw)
oopt = torch.optim.Adam(oracle.parameters(), lr=1e-3)
batches = [collate(graphs[i:i + B]) for i in range(0, len(graphs), B)]


def epoch_c():
    for b in batches:
        oopt.zero_grad()
        loss = F.mse_loss(oracle(b)[0], b.y)
        loss.backward()
        oopt.step()


epoch_c()
t0 = time.perf_counter()
epoch_c()
res["cpu_oracle_mol_per_s"] = len(ds) / (time.perf_counter() - t0)
res["cpu_threads"] = os.cpu_count()
print(json.dumps(res))
