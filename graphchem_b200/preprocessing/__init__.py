from .features import MoleculeEncoder, Tokenizer, load_encoder  # noqa: F401
from .features import atom_to_str, bond_to_str, get_ring_size  # noqa: F401
