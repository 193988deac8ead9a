from graphchem_b200.data.structs import MoleculeGraph, MoleculeDataset  # noqa: F401
