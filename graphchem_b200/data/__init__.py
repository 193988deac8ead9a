from .structs import Batch, Data, DataLoader, MoleculeDataset, MoleculeGraph  # noqa: F401
from .packed import MoleculeStore, PackedBatch, RawBatch, pack_batch, live_bond_rows  # noqa: F401
from .loader import PackedLoader  # noqa: F401
