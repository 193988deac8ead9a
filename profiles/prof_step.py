"""Profiling driver: N un-graphed train steps of the bench workload (batch 8192, D=128, L=4, bf16) so that ncu
sees one kernel launch per line; the LAST step runs between cudaProfilerStart / Stop, so
`ncu --profile-from-start off ...` captures exactly one warm step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import CFGS

CFG = CFGS["train"]
from graphchem_b200.nn import MoleculeGCN
from graphchem_b200.synthetic import make_molecules
from graphchem_b200.train import TrainStep

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = MoleculeGCN(CFG["atom_vocab"], CFG["bond_vocab"], 1, embedding_dim=CFG["D"], n_messages=CFG["L"],
                    n_readout=CFG["n_readout"], readout_dim=CFG["R"], compute_dtype=torch.bfloat16).to(dev)
store = make_molecules(2 * B, seed=1234)
batches = [store.collate_pack(np.arange(i * B, (i + 1) * B), CFG["L"], 64, 32).to(dev) for i in range(2)]
step = TrainStep(model, use_cuda_graph=False)
for i in range(steps):
    if i == steps - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    loss = step(batches[i % 2])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
