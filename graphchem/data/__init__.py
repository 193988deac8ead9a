"""Import-path shim: the reference's data contract types (graphchem/data/structs.py) from graphchem_b200.data."""
import graphchem_b200.data as _b200

__all__ = ["MoleculeGraph", "MoleculeDataset"]
MoleculeGraph, MoleculeDataset = _b200.MoleculeGraph, _b200.MoleculeDataset
