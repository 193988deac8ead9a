from graphchem_b200.data import MoleculeGraph, MoleculeDataset  # noqa: F401
