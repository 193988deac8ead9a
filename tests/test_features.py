"""-m "not gpu": tokenizer contract of reference features.py:122-193 (tests/test_features.py:105-116) and the
encoder's perception back ends; the RDKit-vs-built-in comparison runs only where RDKit is importable."""
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


def test_encoder_backends(tmp_path):
    """Default: RDKit when importable, the built-in perception otherwise (same strings, tests/test_smiles.py);
    backend='rdkit' without RDKit is an ImportError, an unknown backend a ValueError."""
    try:
        import rdkit  # noqa: F401
        have = True
    except ImportError:
        have = False
    if not have:
        with pytest.raises(ImportError):
            MoleculeEncoder(["CC"], backend="rdkit")
    with pytest.raises(ValueError):
        MoleculeEncoder(["CC"], backend="openbabel")
    enc = MoleculeEncoder(["C=C", "CCO"])
    a, b, c = enc.encode("C=C")
    assert a.shape == (2,) and b.shape == (2,) and c.shape == (2, 2) and min(enc.vocab_sizes) > 1
    enc.save(str(tmp_path / "e.pkl"))
    assert load_encoder(str(tmp_path / "e.pkl")).vocab_sizes == enc.vocab_sizes
    if have:  # both perceptions must number a data set identically
        from tests.util import load_fuel_set
        smiles = load_fuel_set("ysi")[0]
        e1, e2 = MoleculeEncoder(smiles, backend="rdkit"), MoleculeEncoder(smiles, backend="builtin")
        assert e1.vocab_sizes == e2.vocab_sizes
        for s in smiles:
            for t1, t2 in zip(e1.encode(s), e2.encode(s)):
                assert (t1 == t2).all()
