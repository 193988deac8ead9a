"""Bundled fuel-property data sets behind the reference's loader names (graphchem/datasets/sets.py:41-110:
`load_cn`, `load_lhv`, `load_mon`, `load_ron`, `load_ysi`), each returning `(smiles: List[str],
targets: float32 Tensor [n, 1])` in the reference's row order -- what the notebooks feed to `train_test_split`
(examples/predict_cn.ipynb:32).  Host-side convenience only (SURVEY.md 2 marks the CSV loaders out of scope for
the hot path); the rows are the same data as tests/golden/fuel_sets.json, which tests/golden/make_fuel_sets.py
generated from the reference's own loader."""
import json
import os
from typing import List, Tuple

import torch

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "static", "fuel_sets.json")
_sets = None


def _load_set(prop: str) -> Tuple[List[str], torch.Tensor]:
    global _sets
    if _sets is None:
        with open(_PATH, "r") as f:
            _sets = json.load(f)
    rows = _sets[prop]
    return list(rows["smiles"]), torch.tensor(rows["target"], dtype=torch.float32).reshape(-1, 1)


def load_cn() -> Tuple[List[str], torch.Tensor]:
    """Cetane number (460 molecules)."""
    return _load_set("cn")


def load_lhv() -> Tuple[List[str], torch.Tensor]:
    """Lower heating value (388 molecules)."""
    return _load_set("lhv")


def load_mon() -> Tuple[List[str], torch.Tensor]:
    """Motor octane number (308 molecules)."""
    return _load_set("mon")


def load_ron() -> Tuple[List[str], torch.Tensor]:
    """Research octane number (308 molecules)."""
    return _load_set("ron")


def load_ysi() -> Tuple[List[str], torch.Tensor]:
    """Yield sooting index (558 molecules)."""
    return _load_set("ysi")
