"""Host-side featurisation contract (reference `graphchem/preprocessing/features.py:10-384`).

Out of scope for the GPU work (north star: "RDKit featurisation stays on host cores"); kept so that
user code written against graphchem keeps importing.  RDKit is imported lazily -- it is not a
dependency of the CUDA path and is absent from the build image.  The OUTPUT LAYOUT is part of the
data contract: int32 atom tokens [n_atoms], int32 bond tokens [2*n_bonds] (both directions of a
bond share one token), int64 connectivity [2, 2*n_bonds] in source-atom-major order
(features.py:294-319); tokens start at 2, `unk` is 1, vocab_size = num_classes + 1 (:148-193).
"""
from __future__ import annotations

import pickle
from typing import Iterable, List, Tuple

import numpy as np
import torch


def _chem():
    try:
        from rdkit import Chem
    except ImportError as e:  # pragma: no cover - RDKit is not installed in the build image
        raise ImportError("MoleculeEncoder needs RDKit (rdkit-pypi==2022.9.5 in the reference's pins); "
                          "pre-tokenised molecules can be fed through graphchem_b200.data.MoleculeStore") from e
    return Chem


def get_ring_size(obj, max_size: int = 12) -> int:
    """Smallest ring size (< max_size) containing the atom/bond, 0 if acyclic, max_size otherwise."""
    if not obj.IsInRing():
        return 0
    return next((k for k in range(max_size) if obj.IsInRingSize(k)), max_size)


_ATOM_FIELDS = ("GetChiralTag", "GetDegree", "GetExplicitValence", "GetFormalCharge", "GetHybridization",
                "GetImplicitValence", "GetIsAromatic", "GetNoImplicit", "GetNumExplicitHs", "GetNumImplicitHs",
                "GetNumRadicalElectrons", "GetSymbol", "GetTotalDegree", "GetTotalNumHs", "GetTotalValence")


def atom_to_str(atom) -> str:
    """str() of the 16-tuple of RDKit atom properties the reference tokenises (features.py:66-83)."""
    return str(tuple(getattr(atom, f)() for f in _ATOM_FIELDS) + (get_ring_size(atom),))


def bond_to_str(bond) -> str:
    """str() of (type, conjugated, stereo, ring size, [sorted end symbols]) (features.py:112-119)."""
    ends = sorted([bond.GetBeginAtom().GetSymbol(), bond.GetEndAtom().GetSymbol()])
    return str((bond.GetBondType(), bond.GetIsConjugated(), bond.GetStereo(), get_ring_size(bond), [ends]))


class Tokenizer:
    """First-seen integer ids: 'unk' -> 1, first real token -> 2; frozen when `train` is False."""

    def __init__(self):
        self._data = {"unk": 1}
        self.num_classes = 1
        self.train = True
        self.unknown: List[str] = []

    def __call__(self, item: str) -> int:
        tok = self._data.get(item)
        if tok is not None:
            return tok
        if not self.train:
            self.unknown.append(item)
            return 1
        self.num_classes += 1
        self._data[item] = self.num_classes
        return self.num_classes

    @property
    def vocab_size(self) -> int:
        return self.num_classes + 1


class MoleculeEncoder:
    def __init__(self, smiles: List[str]):
        Chem = _chem()
        self._atom_tokenizer, self._bond_tokenizer = Tokenizer(), Tokenizer()
        mols = []
        for smi in smiles:
            mol = Chem.MolFromSmiles(smi)
            if mol is None:
                raise ValueError(f"Unable to parse SMILES: {smi}")
            mols.append(mol)
        # the reference tokenises all atoms first, then all bonds (features.py:230-245)
        for mol in mols:
            for atom in mol.GetAtoms():
                self._atom_tokenizer(atom_to_str(atom))
        for mol in mols:
            for atom in mol.GetAtoms():
                for bond in atom.GetBonds():
                    self._bond_tokenizer(bond_to_str(bond))
        self._atom_tokenizer.train = self._bond_tokenizer.train = False

    @property
    def vocab_sizes(self) -> Tuple[int, int]:
        return self._atom_tokenizer.vocab_size, self._bond_tokenizer.vocab_size

    def encode(self, smiles: str) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        mol = _chem().MolFromSmiles(smiles)
        if mol is None:
            raise ValueError(f"Unable to parse SMILES string: {smiles}")
        atoms, bonds, src, dst = [], [], [], []
        for atom in mol.GetAtoms():
            atoms.append(self._atom_tokenizer(atom_to_str(atom)))
            here = atom.GetIdx()
            for bond in atom.GetBonds():  # atom-major, RDKit bond order; row 0 = this atom
                a, b = bond.GetBeginAtomIdx(), bond.GetEndAtomIdx()
                bonds.append(self._bond_tokenizer(bond_to_str(bond)))
                src.append(here)
                dst.append(b if a == here else a)
        conn = torch.from_numpy(np.array([src, dst], dtype=np.int64).reshape(2, len(src)))
        return torch.tensor(atoms, dtype=torch.int32), torch.tensor(bonds, dtype=torch.int32), conn

    def encode_many(self, smiles: Iterable[str]):
        return [self.encode(s) for s in smiles]

    def save(self, filename: str) -> None:
        with open(filename, "wb") as f:
            pickle.dump(self, f, pickle.HIGHEST_PROTOCOL)


def load_encoder(filename: str) -> MoleculeEncoder:
    with open(filename, "rb") as f:
        return pickle.load(f)
