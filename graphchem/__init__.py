"""Import-path shim: `graphchem.nn.MoleculeGCN`, `graphchem.data.*`, `graphchem.preprocessing.*` resolve to
the B200-native implementation in graphchem_b200 (same class API as ecrl/graphchem 2.3.3)."""
from graphchem_b200 import __version__  # noqa: F401
