"""Streaming train loop (one B200): molecules/s from encoded molecules through PackedLoader -> TrainStep with LONG epochs
(32 batches of 8192; bench.py's streaming legs use 8-batch epochs, whose pipeline fill dominates), host packer vs
device packer, loader alone vs loader + train step, and the host time spent inside RawBatch.pack.  JSON lines."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from graphchem_b200.data import PackedLoader
from graphchem_b200.data import packed as packed_mod
from graphchem_b200.nn.gcn import MoleculeGCN
from graphchem_b200.synthetic import make_molecules
from graphchem_b200.train import TrainStep

B, NB = 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 32
store = make_molecules(B * NB, seed=1234)
torch.manual_seed(0)
model = MoleculeGCN(64, 32, 1, embedding_dim=128, n_messages=4, compute_dtype=torch.bfloat16).cuda()
step = TrainStep(model, lr=1e-3, capture_after=1)

t_pack = [0.0, 0]
_orig = packed_mod.RawBatch.pack


def _timed(self, *a, **k):
    t0 = time.perf_counter()
    r = _orig(self, *a, **k)
    t_pack[0] += time.perf_counter() - t0
    t_pack[1] += 1
    return r


packed_mod.RawBatch.pack = _timed


def run(device_pack, workers, with_step, epochs=2):
    loader = PackedLoader(store, B, 4, 64, 32, device="cuda", workers=workers, prefetch=4, device_slots=3,
                          device_pack=device_pack)
    for pb in loader:  # warm-up epoch: pinned pool, slot ring, graph captures
        if with_step:
            step(pb)
    torch.cuda.synchronize()
    t_pack[0], t_pack[1] = 0.0, 0
    t0 = time.perf_counter()
    n = 0
    for _ in range(epochs):
        for pb in loader:
            if with_step:
                step(pb)
            n += 1
    if with_step:
        float(step.loss.item())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"device_pack": device_pack, "workers": workers, "train_step": with_step, "batches": n,
                      "batches_per_epoch": NB, "ms_per_batch": dt / n * 1e3, "M_mol_per_s": B * n / dt / 1e6,
                      "host_ms_in_pack_call": (t_pack[0] / t_pack[1] * 1e3) if t_pack[1] else None}), flush=True)


for dp, w, st in ((False, 16, False), (True, 16, False), (False, 16, True), (True, 16, True), (True, 4, True)):
    run(dp, w, st)
