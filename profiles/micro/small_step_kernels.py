"""Kernel list of ONE 16-molecule fp32 train step (the notebooks' batch; BASELINE configs[0]/[1] regime), from the
library's CUDA-event timing facility: name, launches, us per launch.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from graphchem_b200 import _lib
from graphchem_b200.data import MoleculeGraph, MoleculeStore
from graphchem_b200.nn import MoleculeGCN
from graphchem_b200.train import TrainStep
from tests.util import random_batch

dtype = torch.bfloat16 if len(sys.argv) > 1 and sys.argv[1] == "bf16" else torch.float32
_, graphs = random_batch(seed=1, n_graphs=64, av=42, bv=29, min_atoms=5, max_atoms=22)
store = MoleculeStore.from_graphs([MoleculeGraph(g.x, g.edge_attr, g.edge_index, g.y) for g in graphs])
model = MoleculeGCN(42, 29, 1, embedding_dim=128, n_messages=3, n_readout=3, readout_dim=128, compute_dtype=dtype).cuda()
pbs = [store.collate_pack(np.arange(16 * i, 16 * i + 16), 3, 42, 29).to("cuda") for i in range(4)]
step = TrainStep(model, use_cuda_graph=False)
lib = _lib.lib()
for pb in pbs:
    step(pb)
torch.cuda.synchronize()
l0 = lib.gcb_launch_count()
lib.gcb_timing_enable(1)
step(pbs[0])
torch.cuda.synchronize()
n = int(lib.gcb_launch_count() - l0)
buf = C.create_string_buffer(1 << 16)
lib.gcb_timing_report(buf, len(buf))
lib.gcb_timing_enable(0)
rows = json.loads(buf.value.decode())
print(json.dumps({"dtype": str(dtype), "N": pbs[0].N, "E": pbs[0].E, "launches": n,
                  "kernels": [[r["name"], r["launches"], round(1e3 * r["ms"] / r["launches"], 1)] for r in rows]}))
