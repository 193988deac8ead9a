"""Import-path shim: `graphchem.datasets.load_*` resolve to graphchem_b200.datasets (same names and return
contract as ecrl/graphchem 2.3.3)."""
import graphchem_b200.datasets as _b200

__all__ = ["load_cn", "load_lhv", "load_mon", "load_ron", "load_ysi"]
for _name in __all__:
    globals()[_name] = getattr(_b200, _name)
del _name
