"""Import-path shim: encoder, tokenizer and the atom / bond string functions of graphchem/preprocessing/features.py,
served by graphchem_b200.preprocessing (RDKit when importable, the built-in perception otherwise)."""
import graphchem_b200.preprocessing as _b200

__all__ = ["MoleculeEncoder", "Tokenizer", "load_encoder", "atom_to_str", "bond_to_str", "get_ring_size"]
for _name in __all__:
    globals()[_name] = getattr(_b200, _name)
del _name
