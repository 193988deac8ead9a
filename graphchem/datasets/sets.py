"""Import-path shim for `graphchem.datasets.sets` (see graphchem_b200/datasets/__init__.py)."""
import graphchem_b200.datasets as _b200

load_cn, load_lhv, load_mon = _b200.load_cn, _b200.load_lhv, _b200.load_mon
load_ron, load_ysi = _b200.load_ron, _b200.load_ysi
