"""`graphchem_b200.datasets` (reference graphchem/datasets/sets.py:41-110): same names, same return contract, same
rows as the reference's loader (via the committed fixture; directly against the reference when it is present)."""
import importlib.util
import os

import pytest
import torch

from tests.util import load_fuel_set

SETS = {"cn": 460, "ysi": 558, "ron": 308, "mon": 308, "lhv": 388}
REF = "/root/reference/graphchem/datasets/sets.py"


@pytest.mark.parametrize("name", list(SETS))
def test_loader_contract_and_rows(name):
    import graphchem.datasets as shim
    import graphchem_b200.datasets as ds
    smiles, y = getattr(ds, f"load_{name}")()
    assert getattr(shim, f"load_{name}") is getattr(ds, f"load_{name}")
    assert isinstance(smiles, list) and all(isinstance(s, str) for s in smiles) and len(smiles) == SETS[name]
    assert y.dtype == torch.float32 and tuple(y.shape) == (SETS[name], 1)          # sets.py:32-37
    g_smiles, g_y, train_idx, test_idx, _ = load_fuel_set(name)
    assert smiles == g_smiles and torch.equal(y, g_y)


def test_notebook_split_is_the_fixture_split():
    from sklearn.model_selection import train_test_split
    import graphchem_b200.datasets as ds
    smiles, y = ds.load_cn()
    _, _, train_idx, test_idx, _ = load_fuel_set("cn")
    tr, te = train_test_split(list(range(len(smiles))), test_size=0.2, random_state=42)  # predict_cn.ipynb:32
    assert tr == train_idx and te == test_idx


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", list(SETS))
def test_rows_equal_the_reference_loader(name):
    spec = importlib.util.spec_from_file_location("_ref_sets", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import graphchem_b200.datasets as ds
    r_smiles, r_y = getattr(mod, f"load_{name}")()
    smiles, y = getattr(ds, f"load_{name}")()
    assert smiles == r_smiles and torch.equal(y, r_y)
