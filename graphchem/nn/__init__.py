"""Import-path shim: `graphchem.nn.MoleculeGCN` is the B200-native drop-in of graphchem_b200.nn."""
import graphchem_b200.nn as _b200

__all__ = ["MoleculeGCN"]
MoleculeGCN = _b200.MoleculeGCN
