from graphchem_b200.preprocessing import MoleculeEncoder, Tokenizer, load_encoder  # noqa: F401
from graphchem_b200.preprocessing import atom_to_str, bond_to_str, get_ring_size  # noqa: F401
