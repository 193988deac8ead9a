"""Host-side featurisation (reference `graphchem/preprocessing/features.py:10-384`).

Stays on host cores (north star) and is not timed as GPU work.  The molecule perception behind the
property tuples comes from RDKit when it is importable and from the built-in restatement
`graphchem_b200.preprocessing.smiles` otherwise (RDKit is absent from the build image); both yield the
same strings, hence the same first-seen token ids (pinned by tests/test_smiles.py on the reference's
test strings, the four notebook encodings and the eight notebook vocabulary sizes).  `backend=` on
`MoleculeEncoder` forces one of them.  The OUTPUT LAYOUT is part of the data contract: int32 atom tokens [n_atoms], int32 bond tokens [2*n_bonds] (both directions of a
bond share one token), int64 connectivity [2, 2*n_bonds] in source-atom-major order
(features.py:294-319); tokens start at 2, `unk` is 1, vocab_size = num_classes + 1 (:148-193).
"""
from __future__ import annotations

import pickle
from typing import Iterable, List, Tuple

import numpy as np
import torch


def _chem(backend: str = "auto"):
    """The object whose `MolFromSmiles` the encoder calls: `rdkit.Chem`, or the built-in perception."""
    if backend not in ("auto", "rdkit", "builtin"):
        raise ValueError(f"backend must be 'auto', 'rdkit' or 'builtin', not {backend!r}")
    if backend != "builtin":
        try:
            from rdkit import Chem
            return Chem
        except ImportError:
            if backend == "rdkit":
                raise ImportError("backend='rdkit' needs RDKit (rdkit-pypi==2022.9.5 in the reference's pins)")
    from . import smiles
    return smiles


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
    def __init__(self, smiles: List[str], backend: str = "auto"):
        Chem = _chem(backend)
        self._backend = backend
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
        mol = _chem(getattr(self, "_backend", "auto")).MolFromSmiles(smiles)
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
