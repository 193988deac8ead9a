"""-m "not gpu": tokenizer contract of reference features.py:122-193 (tests/test_features.py:105-116);
the RDKit-backed encoder tests run only where RDKit is importable (not in the build image)."""
import pickle

import pytest

from graphchem_b200.preprocessing import MoleculeEncoder, Tokenizer, load_encoder


def test_tokenizer_ids_and_unknowns():
    t = Tokenizer()
    assert t("a") == 2 and t("b") == 3 and t("a") == 2
    assert t.num_classes == 3 and t.vocab_size == 4
    t.train = False
    assert t("zzz") == 1 and t.unknown == ["zzz"] and t.vocab_size == 4
    assert pickle.loads(pickle.dumps(t))("b") == 3


def test_import_paths_of_the_reference_resolve():
    from graphchem.nn import MoleculeGCN
    from graphchem.data import MoleculeDataset, MoleculeGraph
    from graphchem.preprocessing import MoleculeEncoder as E2
    import graphchem_b200.nn as nn2
    assert MoleculeGCN is nn2.MoleculeGCN and E2 is MoleculeEncoder and MoleculeGraph and MoleculeDataset


def test_encoder_needs_rdkit_or_matches_golden(tmp_path):
    try:
        import rdkit  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            MoleculeEncoder(["CC"])
        return
    import json, os
    enc = MoleculeEncoder(["C=C"])
    a, b, c = enc.encode("C=C")
    assert a.shape == (2,) and b.shape == (2,) and c.shape == (2, 2) and min(enc.vocab_sizes) > 1
    enc.save(str(tmp_path / "e.pkl"))
    assert load_encoder(str(tmp_path / "e.pkl")).vocab_sizes == enc.vocab_sizes
