"""Where does a Screener call spend its time?  (one B200)  Loader alone (raw collate -> H2D -> device pack), then the
whole call, for several worker counts; prints JSON lines."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from graphchem_b200.data import PackedLoader
from graphchem_b200.nn import MoleculeGCN
from graphchem_b200.synthetic import make_molecules
from graphchem_b200.train import Screener

B, NB = 32768, 12
dev = torch.device("cuda", 0)
model = MoleculeGCN(64, 32, 1, embedding_dim=256, n_messages=6, n_readout=2, readout_dim=64, compute_dtype=torch.bfloat16).to(dev)
store = make_molecules(NB * B, seed=4321)
for workers in (16, 8, 4):
    ld = PackedLoader(store, B, 6, 64, 32, device=dev, workers=workers, prefetch=4, device_slots=3, device_pack=True, lookahead=False)
    ts = []
    for ep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for pb in ld:
            pass
        torch.cuda.synchronize(); ts.append(round((time.perf_counter() - t0) * 1e3 / NB, 2))
    t0 = time.perf_counter(); raw = store.collate_raw(range(B), 6, 64, 32, pin=False); t_raw = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"loader_alone_ms_per_batch": ts, "workers": workers, "raw_collate_ms_1core": round(t_raw, 2)}), flush=True)
    ld.close()
    sc = Screener(model, batch_size=B, workers=workers)
    ts = []
    for ep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sc(store)
        torch.cuda.synchronize(); ts.append(round((time.perf_counter() - t0) * 1e3 / NB, 2))
    print(json.dumps({"screener_ms_per_batch": ts, "workers": workers}), flush=True)
