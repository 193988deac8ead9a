"""Side measurements for DESIGN.md (not the bench line): train step at other batch sizes / fp32, and
forward-only (inference) throughput including BASELINE.json configs[4]'s shape (D=256, n_messages=6).
CUDA-event timed over resident pre-packed batches (4 distinct batches cycled), one JSON line per case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from graphchem_b200.nn import MoleculeGCN
from graphchem_b200.synthetic import make_molecules
from graphchem_b200.train import InferStep, TrainStep

dev = torch.device("cuda", 0)


def timed(fn, n_warm, n):
    for i in range(n_warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def case(mode, B, D, L, dtype):
    torch.manual_seed(0)
    model = MoleculeGCN(64, 32, 1, embedding_dim=D, n_messages=L, n_readout=2, readout_dim=64, compute_dtype=dtype).to(dev)
    store = make_molecules(4 * B, seed=7)
    batches = [store.collate_pack(np.arange(i * B, (i + 1) * B), L, 64, 32).to(dev) for i in range(4)]
    if mode == "train":
        step = TrainStep(model, lr=1e-3, capture_after=1)
        ms = timed(lambda i: step(batches[i % 4]), 8, 20)
    elif mode == "infer":  # CUDA-graphed forward on resident batches
        step = InferStep(model)
        ms = timed(lambda i: step(batches[i % 4]), 8, 20)
    else:
        model.eval()
        with torch.no_grad():
            ms = timed(lambda i: model(batches[i % 4]), 4, 12)
    print(json.dumps({"mode": mode, "batch": B, "D": D, "n_messages": L, "dtype": str(dtype).split(".")[-1],
                      "ms": round(ms, 4), "molecules_per_s": round(B / ms * 1e3)}), flush=True)


for args in [("train", 4096, 128, 4, torch.bfloat16), ("train", 16384, 128, 4, torch.bfloat16),
             ("train", 8192, 128, 4, torch.float32), ("forward", 8192, 128, 4, torch.bfloat16),
             ("forward", 8192, 256, 6, torch.bfloat16), ("forward", 32768, 128, 4, torch.bfloat16),
             ("infer", 8192, 128, 4, torch.bfloat16), ("infer", 32768, 128, 4, torch.bfloat16),
             ("infer", 8192, 256, 6, torch.bfloat16), ("infer", 32768, 256, 6, torch.bfloat16)]:
    case(*args)
