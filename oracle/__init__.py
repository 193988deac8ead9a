"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference hot path (graphchem MoleculeGCN on
PyG 2.6.1).  Imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only;
never by the product package graphchem_b200/."""
