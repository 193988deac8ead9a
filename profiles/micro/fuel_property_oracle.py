"""CPU twin of profiles/micro/fuel_property.py: the same notebook recipe on the same real molecules, the same initial
weights (OracleMoleculeGCN under torch.manual_seed(seed)) and the same batch order (PackedLoader's permutation rule),
run on the CPU oracle (the reference restated, oracle/gcn_oracle.py) -- what the reference stack would compute on host
cores.  Prints one JSON line per data set; the GPU script's numbers go beside it (profiles/README.md).

    python profiles/micro/fuel_property_oracle.py [--sets cn] [--epochs 400] [--seed 0] [--threads 4]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from oracle.gcn_oracle import OracleMoleculeGCN, train_step
from oracle.pyg_restated import Data as OData, collate
from tests.util import encoded_fuel_set

KW = dict(embedding_dim=128, n_messages=3, n_readout=3, readout_dim=128)


def median_ae_r2(pred, y):
    pred, y = pred.double().flatten(), y.double().flatten()
    return float((pred - y).abs().median()), 1.0 - float(((pred - y) ** 2).sum() / ((y - y.mean()) ** 2).sum())


def run(name, epochs, seed):
    enc, train, test = encoded_fuel_set(name)
    av, bv = enc.vocab_sizes
    torch.manual_seed(seed)
    model = OracleMoleculeGCN(av, bv, 1, **KW).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    od = [OData(x=g.x, edge_index=g.edge_index, edge_attr=g.edge_attr, y=g.y) for g in train]
    curve = {}
    t0 = time.perf_counter()
    for epoch in range(epochs):
        for g in opt.param_groups:
            g["lr"] = max(0.0, 1e-3 - epoch * 1e-8)
        perm = np.random.default_rng(seed + epoch + 1).permutation(len(train))   # PackedLoader._batches, epoch counted from 1
        tot = 0.0
        for i in range(0, len(train), 16):
            tot += float(train_step(model, opt, collate([od[j] for j in perm[i:i + 16]])))
        if epoch % 25 == 0 or epoch == epochs - 1:
            curve[epoch] = tot / len(train)
    t_train = time.perf_counter() - t0
    model.eval()
    with torch.no_grad():
        p_test = torch.cat([model(OData(x=g.x, edge_index=g.edge_index, edge_attr=g.edge_attr))[0] for g in test])
        p_train = torch.cat([model(OData(x=g.x, edge_index=g.edge_index, edge_attr=g.edge_attr))[0] for g in train])
    mae, r2 = median_ae_r2(p_test, torch.cat([g.y for g in test]))
    mae_tr, r2_tr = median_ae_r2(p_train, torch.cat([g.y for g in train]))
    return {"set": name, "impl": "cpu oracle", "seed": seed, "epochs": epochs, "threads": torch.get_num_threads(),
            "train_s": round(t_train, 1), "train_molecules_per_s": round(epochs * len(train) / t_train, 1),
            "loss_per_molecule": curve, "test_median_ae": round(mae, 4), "test_r2": round(r2, 4),
            "train_median_ae": round(mae_tr, 4), "train_r2": round(r2_tr, 4)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", nargs="+", default=["cn"])
    ap.add_argument("--epochs", type=int, default=400)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--threads", type=int, default=4)
    a = ap.parse_args()
    torch.set_num_threads(a.threads)
    for s in a.sets:
        print(json.dumps(run(s, a.epochs, a.seed)), flush=True)
