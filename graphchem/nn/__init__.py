from graphchem_b200.nn import MoleculeGCN  # noqa: F401
