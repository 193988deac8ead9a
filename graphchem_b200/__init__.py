"""graphchem_b200 -- B200-native (sm_100a) implementation of graphchem's MoleculeGCN hot path.

Public surface (mirrors ecrl/graphchem 2.3.3 for this path):
    graphchem_b200.nn.MoleculeGCN
    graphchem_b200.data.{MoleculeGraph, MoleculeDataset, Data, Batch, DataLoader, PackedBatch, MoleculeStore}
    graphchem_b200.preprocessing.{MoleculeEncoder, Tokenizer, load_encoder, ...}   (RDKit imported lazily)
    graphchem_b200.optim.FusedAdam, graphchem_b200.train.TrainStep
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
from . import data, nn  # noqa: F401
