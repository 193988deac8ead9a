"""Import-path shim for `graphchem.preprocessing.features` (see graphchem_b200/preprocessing/features.py)."""
import graphchem_b200.preprocessing.features as _b200

for _name in dir(_b200):
    if not _name.startswith("_"):
        globals()[_name] = getattr(_b200, _name)
del _name
