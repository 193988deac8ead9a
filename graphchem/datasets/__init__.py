from graphchem_b200.datasets import load_cn, load_lhv, load_mon, load_ron, load_ysi  # noqa: F401
