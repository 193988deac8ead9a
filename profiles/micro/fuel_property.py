"""BASELINE.json configs[0] / configs[1] on the REAL bundled molecules: the notebook recipe of the reference
(examples/predict_cn.ipynb cells 0-5 and its YSI / RON / MON / LHV twins: sklearn split with random_state 42,
MoleculeEncoder on the training SMILES, MoleculeGCN(D=128, n_messages=3, n_readout=3, readout_dim=128), batch 16
reshuffled every epoch, Adam lr 1e-3 - epoch*1e-8, MSE, 400 epochs, per-molecule evaluation, median absolute
error and R^2 on the test split), fp32, on one B200:

  SMILES -> built-in featuriser (host, graphchem_b200.preprocessing.smiles) -> MoleculeStore -> PackedLoader
  (C packer, pinned ring, shuffle) -> TrainStep (forward + MSE + backward + Adam, C-ABI) -> predict_each.

Prints one JSON line per data set with the loss curve, the test metrics beside the notebook's printed ones
(the notebooks are unseeded, so agreement is distributional), the training throughput, and
the first-steps parity of the loss trajectory against the CPU oracle on identical batches and weights.

Initial weights and batch order are those of the CPU twin profiles/micro/fuel_property_oracle.py (same seed), so the
two loss curves start together and then drift apart the way any two fp32 runs of a chaotic optimisation do.

    python profiles/micro/fuel_property.py [--sets cn ysi ron mon lhv] [--epochs 400] [--seeds 0 1 2]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from graphchem_b200.data import MoleculeStore, PackedLoader
from graphchem_b200.nn import MoleculeGCN
from graphchem_b200.train import TrainStep
from tests.util import encoded_fuel_set

EPOCH_RUNNER = True   # --step-by-step: the streamed loader + one TrainStep call per batch instead
NOTEBOOK = {"cn": (4.83, 0.884), "ysi": (5.65, 0.980), "ron": (5.65, 0.758), "mon": (6.96, 0.611), "lhv": (0.575, 0.911)}
KW = dict(embedding_dim=128, n_messages=3, n_readout=3, readout_dim=128)


def median_ae_r2(pred: torch.Tensor, y: torch.Tensor):
    pred, y = pred.double().flatten(), y.double().flatten()
    mae = float((pred - y).abs().median())
    r2 = 1.0 - float(((pred - y) ** 2).sum() / ((y - y.mean()) ** 2).sum())
    return mae, r2


def oracle_trajectory(name, enc, train, steps=5):
    """Loss of the first `steps` Adam steps on the unshuffled batches, CUDA path vs CPU oracle, same weights."""
    import torch.nn.functional as F
    from oracle.gcn_oracle import OracleMoleculeGCN, train_step
    from oracle.pyg_restated import Data as OData, collate
    av, bv = enc.vocab_sizes
    torch.manual_seed(0)
    oracle = OracleMoleculeGCN(av, bv, 1, **KW)
    model = MoleculeGCN(av, bv, 1, **KW)
    model.load_state_dict(oracle.state_dict())
    model = model.cuda().train()
    oopt = torch.optim.Adam(oracle.parameters(), lr=1e-3)
    step = TrainStep(model, lr=1e-3, use_cuda_graph=False)
    store = MoleculeStore.from_graphs(train)
    out = []
    for k in range(steps):
        idx = np.arange(16 * k, 16 * (k + 1))
        lo = float(train_step(oracle, oopt, collate([OData(x=train[i].x, edge_index=train[i].edge_index,
                                                          edge_attr=train[i].edge_attr, y=train[i].y) for i in idx])))
        lc = float(step(store.collate_pack(idx, KW["n_messages"], av, bv).to("cuda")).item())
        out.append({"oracle": lo, "cuda": lc, "rel": abs(lc - lo) / max(abs(lo), 1e-30)})
    return out


def run(name: str, epochs: int, seed: int = 0, oracle_steps: bool = True, dtype=torch.float32):
    t0 = time.perf_counter()
    enc, train, test = encoded_fuel_set(name)
    t_feat = time.perf_counter() - t0
    av, bv = enc.vocab_sizes
    from oracle.gcn_oracle import OracleMoleculeGCN
    torch.manual_seed(seed)                      # same initial weights as the CPU twin (fuel_property_oracle.py)
    init = OracleMoleculeGCN(av, bv, 1, **KW).state_dict()
    model = MoleculeGCN(av, bv, 1, compute_dtype=dtype, **KW)
    model.load_state_dict(init)
    model = model.cuda().train()
    store = MoleculeStore.from_graphs(train)
    loader = PackedLoader(store, 16, KW["n_messages"], av, bv, device="cuda", shuffle=True, seed=seed, workers=2, prefetch=4)
    step = TrainStep(model, lr=1e-3, use_cuda_graph=False)   # every shuffled batch has its own (N, E): nothing to replay
    curve = {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for epoch in range(epochs):
        step.lr = max(0.0, 1e-3 - epoch * 1e-8)
        if EPOCH_RUNNER:
            # same batches in the same order as the loader would deliver them, packed and shipped once per epoch,
            # stepped by one library call
            losses = step.run_epoch(store.collate_pack_epoch(loader._batches(epoch + 1), KW["n_messages"], av, bv, "cuda"))
        else:
            losses = torch.stack([step(pb).reshape(()).clone() for pb in loader])
        if epoch % 25 == 0 or epoch == epochs - 1:
            curve[epoch] = float(losses.sum().item()) / len(train)   # the notebook's train_loss / len(dataset)
    torch.cuda.synchronize()
    t_train = time.perf_counter() - t0
    model.eval()
    y_test = torch.cat([g.y for g in test])
    y_train = torch.cat([g.y for g in train])
    mae, r2 = median_ae_r2(model.predict_each(test), y_test)
    mae_tr, r2_tr = median_ae_r2(model.predict_each(train), y_train)
    return {"set": name, "impl": "cuda", "dtype": str(dtype).replace("torch.", ""), "seed": seed, "molecules_train": len(train), "molecules_test": len(test), "vocab_sizes": [av, bv],
            "epochs": epochs, "featurise_s": round(t_feat, 3), "train_s": round(t_train, 3),
            "train_molecules_per_s": round(epochs * len(train) / t_train, 1), "loss_per_molecule": curve,
            "test_median_ae": round(mae, 4), "test_r2": round(r2, 4), "train_median_ae": round(mae_tr, 4),
            "train_r2": round(r2_tr, 4), "notebook_test_median_ae_r2": NOTEBOOK[name],
            "oracle_first_steps": oracle_trajectory(name, enc, train) if oracle_steps else None}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", nargs="+", default=["cn", "ysi"])
    ap.add_argument("--epochs", type=int, default=400)
    ap.add_argument("--seeds", nargs="+", type=int, default=[0])
    ap.add_argument("--step-by-step", action="store_true", help="PackedLoader + one TrainStep call per batch (default: the epoch runner)")
    ap.add_argument("--no-oracle-steps", action="store_true", help="skip the first-steps comparison with the CPU oracle")
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"], help="activation dtype (bf16: tile kernels + tcgen05 GEMMs)")
    a = ap.parse_args()
    EPOCH_RUNNER = not a.step_by_step
    for s in a.sets:
        for seed in a.seeds:
            print(json.dumps(run(s, a.epochs, seed, not a.no_oracle_steps and a.dtype == "fp32",
                                 torch.bfloat16 if a.dtype == "bf16" else torch.float32)), flush=True)
