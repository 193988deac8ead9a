from .gcn import MoleculeGCN  # noqa: F401
