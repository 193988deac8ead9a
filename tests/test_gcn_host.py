"""Host-side contract of the drop-in `MoleculeGCN` (no GPU needed): constructor signature, attributes,
`state_dict` keys / shapes against the restated reference (graphchem/nn/gcn.py:35-118; the reference's own
tests check exactly these observables, tests/test_gcn.py:22-24, 85), accepted activations, and that the
product refuses to compute without a CUDA device (no CPU fallback)."""
import inspect

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from graphchem_b200 import _lib
from graphchem_b200.nn.gcn import MoleculeGCN
from oracle.gcn_oracle import OracleMoleculeGCN


def test_constructor_signature_matches_reference():
    # gcn.py:35-42: names, order and defaults
    sig = inspect.signature(MoleculeGCN.__init__)
    ref = inspect.signature(OracleMoleculeGCN.__init__)
    names = [n for n, p in sig.parameters.items() if p.kind is not inspect.Parameter.KEYWORD_ONLY]
    assert names == [n for n, p in ref.parameters.items() if p.kind is not inspect.Parameter.KEYWORD_ONLY]
    for n in names[1:]:
        assert sig.parameters[n].default == ref.parameters[n].default


@pytest.mark.parametrize("dims", [
    dict(atom_vocab_size=64, bond_vocab_size=32, output_dim=1),                                    # defaults
    dict(atom_vocab_size=42, bond_vocab_size=29, output_dim=1, embedding_dim=128, n_messages=3,
         n_readout=3, readout_dim=128),                                                            # predict_cn.ipynb:218-227
    dict(atom_vocab_size=10, bond_vocab_size=5, output_dim=None, embedding_dim=64, n_messages=3,
         n_readout=0),                                                                             # tests/test_gcn.py:70-85
])
def test_state_dict_keys_and_shapes(dims):
    m, o = MoleculeGCN(**dims), OracleMoleculeGCN(**dims)
    sm, so = m.state_dict(), o.state_dict()
    assert list(sm) == list(so)
    assert [tuple(v.shape) for v in sm.values()] == [tuple(v.shape) for v in so.values()]
    assert (m.readout is None) == (dims.get("n_readout", 2) == 0)
    m.load_state_dict(so)  # weights travel between the two by name
    for k in so:
        assert torch.equal(m.state_dict()[k], so[k])


def test_notebook_parameter_count_and_attributes():
    m = MoleculeGCN(42, 29, 1, embedding_dim=128, n_messages=3, n_readout=3, readout_dim=128, p_dropout=0.2)
    assert sum(p.numel() for p in m.parameters()) == 124673  # predict_cn.ipynb:181-196
    assert m._p_dropout == 0.2 and m._n_messages == 3 and m.act_fn is F.softplus
    assert isinstance(m.emb_atom, nn.Embedding) and isinstance(m.emb_bond, nn.Embedding)
    assert isinstance(m.readout, nn.ModuleList) and len(m.readout) == 4
    assert m.training and not m.eval().training


@pytest.mark.parametrize("fn,aid", [
    (F.softplus, _lib.ACT_SOFTPLUS), (F.sigmoid, _lib.ACT_SIGMOID), (torch.sigmoid, _lib.ACT_SIGMOID),
    (F.relu, _lib.ACT_RELU), (F.tanh, _lib.ACT_TANH), (F.leaky_relu, _lib.ACT_LEAKY_RELU),
    (nn.Softplus(), _lib.ACT_SOFTPLUS), (nn.ReLU(), _lib.ACT_RELU), (nn.Tanh(), _lib.ACT_TANH),
    (nn.Sigmoid(), _lib.ACT_SIGMOID), (nn.LeakyReLU(), _lib.ACT_LEAKY_RELU),
])
def test_accepted_activations(fn, aid):
    m = MoleculeGCN(8, 4, 1, embedding_dim=16, act_fn=fn)
    assert m.act_fn is fn and m._act_id == aid


@pytest.mark.parametrize("fn", [F.gelu, nn.Softplus(beta=2.0), nn.LeakyReLU(0.2), lambda t: t])
def test_unfused_activation_is_refused_not_approximated(fn):
    with pytest.raises(ValueError):
        MoleculeGCN(8, 4, 1, embedding_dim=16, act_fn=fn)


def test_pool_keyword():
    assert MoleculeGCN(8, 4, 1, embedding_dim=16)._cfg().pool == 0
    assert MoleculeGCN(8, 4, 1, embedding_dim=16, pool="mean")._cfg().pool == 1
    with pytest.raises(ValueError):
        MoleculeGCN(8, 4, 1, embedding_dim=16, pool="max")


def test_aggregations():
    """gcn.py:84-86 hands `aggr` to GeneralConv: add (default), mean and max are built, anything else is refused."""
    with pytest.raises(NotImplementedError):
        MoleculeGCN(8, 4, 1, embedding_dim=16, aggr="min")
    for aggr in ("add", "mean", "max"):
        assert MoleculeGCN(8, 4, 1, embedding_dim=16, aggr=aggr).atom_conv.aggr == aggr


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour of a box without a GPU")
def test_no_cpu_fallback():
    from graphchem_b200.data.structs import Data
    m = MoleculeGCN(8, 4, 1, embedding_dim=16)
    d = Data(x=torch.tensor([1, 2], dtype=torch.int32), edge_index=torch.tensor([[0, 1], [1, 0]]),
             edge_attr=torch.tensor([1, 1], dtype=torch.int32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(d)
