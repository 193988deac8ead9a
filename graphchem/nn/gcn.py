"""Import-path shim for `graphchem.nn.gcn.MoleculeGCN` (see graphchem_b200/nn/gcn.py)."""
import graphchem_b200.nn.gcn as _b200

MoleculeGCN = _b200.MoleculeGCN
