"""One-process timing of the bench workload (one B200): graphed train step (CUDA events, 8 resident batches cycled),
forward-only InferStep, and the per-kernel CUDA-event table of un-graphed steps.  The GCB_* switches are read from the
environment, so an A/B comparison is one invocation per flag set:
    GCB_NO_FUSED=1 python profiles/micro/step_time.py [--batch 8192] [--steps 30] [--D 128] [--L 4]
Prints ONE JSON line."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from graphchem_b200 import _lib
from graphchem_b200.nn import MoleculeGCN
from graphchem_b200.synthetic import make_molecules
from graphchem_b200.train import InferStep, TrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--D", type=int, default=128)
ap.add_argument("--L", type=int, default=4)
ap.add_argument("--resident", type=int, default=8)
ap.add_argument("--no-train", action="store_true")
ap.add_argument("--kernels", action="store_true", help="per-kernel CUDA-event table of un-graphed steps")
args = ap.parse_args()

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, NR = args.batch, args.resident
model = MoleculeGCN(64, 32, 1, embedding_dim=args.D, n_messages=args.L, n_readout=2, readout_dim=64,
                    compute_dtype=torch.bfloat16).to(dev)
store = make_molecules(NR * B, seed=1234)
batches = [store.collate_pack(np.arange(i * B, (i + 1) * B), args.L, 64, 32).to(dev) for i in range(NR)]
lib = _lib.lib()
out = {"flags": {k: v for k, v in os.environ.items() if k.startswith("GCB_")}, "batch": B, "D": args.D, "L": args.L}


def timed(fn, n):
    for i in range(max(NR, 10)):
        fn(batches[i % NR])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(batches[i % NR])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if not args.no_train:
    step = TrainStep(model, lr=1e-3, capture_after=1)  # resident batches repeat: capture on first sight
    ms = timed(step, args.steps)
    out["train_ms_per_step"] = round(ms, 4)
    out["train_mol_per_s"] = round(B / (ms * 1e-3))
    out["loss"] = float(step.loss.item())
infer = InferStep(model)
ms = timed(infer, args.steps)
out["infer_ms_per_step"] = round(ms, 4)
out["infer_mol_per_s"] = round(B / (ms * 1e-3))
if args.kernels and not args.no_train:
    plain = TrainStep(model, lr=1e-3, use_cuda_graph=False)
    for i in range(3):
        plain(batches[i % NR])
    torch.cuda.synchronize()
    l0 = lib.gcb_launch_count()
    plain(batches[0])
    out["launches_per_step"] = int(lib.gcb_launch_count() - l0)
    lib.gcb_timing_enable(1)
    n_prof = 6
    for i in range(n_prof):
        plain(batches[i % NR])
    torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 16)
    lib.gcb_timing_report(buf, len(buf))
    lib.gcb_timing_enable(0)
    rep = json.loads(buf.value.decode())
    out["kernels"] = {r["name"]: [round(r["launches"] / n_prof, 1), round(r["ms"] / r["launches"] * 1e3, 1)]
                      for r in sorted(rep, key=lambda r: -r["ms"])}
    out["kernels_sum_ms"] = round(sum(r["ms"] for r in rep) / n_prof, 4)
print(json.dumps(out), flush=True)
