"""Generates tests/golden/fuel_sets.json: the five bundled property data sets of the reference (SMILES + target,
file order) as its own loader returns them (`graphchem/datasets/sets.py:12-37`, imported by path from
/root/reference), the train/test split of the notebooks (`train_test_split(test_size=0.2, random_state=42)`,
examples/predict_cn.ipynb cell 0) as index lists, and the integers the notebooks print for that split:
the (atom, bond) vocabulary sizes of `MoleculeEncoder(smiles_train)` (module trees: predict_cn.ipynb:182-183,
predict_ysi.ipynb:179-180, predict_ron.ipynb:179-180, predict_lhv.ipynb:179-180).

The reference cannot travel to the GPU box and RDKit / PyG are not installable here, so these INPUTS and
notebook-printed integers are the committed fixture for SURVEY.md 8 f2 (featuriser) and BASELINE configs[0]/[1].
Regenerate (in the build container only):  python tests/golden/make_fuel_sets.py"""
import importlib.util
import json
import os

from sklearn.model_selection import train_test_split

REF = "/root/reference/graphchem/datasets/sets.py"
NOTEBOOK_VOCAB = {"cn": [42, 29], "ysi": [68, 56], "ron": [45, 32], "mon": [45, 32], "lhv": [63, 54]}
NOTEBOOK_SPLIT = {"cn": [368, 92], "ysi": [446, 112], "ron": [246, 62], "mon": [246, 62], "lhv": [310, 78]}


def main():
    spec = importlib.util.spec_from_file_location("ref_sets", REF)
    sets = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sets)
    out = {}
    for name in ("cn", "ysi", "ron", "mon", "lhv"):
        smiles, y = getattr(sets, f"load_{name}")()
        idx = list(range(len(smiles)))
        train_idx, test_idx = train_test_split(idx, test_size=0.2, random_state=42)
        s_train, _, _, _ = train_test_split(smiles, y, test_size=0.2, random_state=42)
        assert [smiles[i] for i in train_idx] == s_train          # same permutation as the notebooks' call
        assert [len(train_idx), len(test_idx)] == NOTEBOOK_SPLIT[name], (name, len(train_idx), len(test_idx))
        out[name] = {"smiles": smiles, "target": [round(float(v), 6) for v in y[:, 0]], "train_idx": train_idx,
                     "test_idx": test_idx, "notebook_vocab_sizes": NOTEBOOK_VOCAB[name]}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fuel_sets.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
