"""-m "not gpu": pin the oracle with everything the reference offers for this path.

The reference's tests hold shape assertions only (tests/test_gcn.py:22-24,60-62,85,99-101,149-151);
they are replayed here against the oracle, together with the structural facts SURVEY.md 3.3
derives from gcn.py:163 (EdgeConv run on the bond matrix) and the notebook golden encodings
(tests/golden/encodings.json)."""
import json
import math
import os

import pytest
import torch
import torch.nn.functional as F

from oracle.gcn_oracle import OracleMoleculeGCN, train_step
from oracle.pyg_restated import Data, collate
from tests.util import nerr, random_batch

HERE = os.path.dirname(os.path.abspath(__file__))


def _ref_test_data(n_atoms=3, n_bonds=5, av=10, bv=5):
    return Data(x=torch.randint(0, av, (n_atoms,)), edge_index=torch.randint(0, n_atoms, (2, n_bonds)),
                edge_attr=torch.randint(0, bv, (n_bonds,)), batch=torch.tensor([0] * n_atoms, dtype=torch.long))


def test_reference_shape_tests_hold_for_the_oracle():
    m = OracleMoleculeGCN(10, 5, 2, embedding_dim=64, n_messages=3, n_readout=3, readout_dim=128, p_dropout=0.2)
    assert m._p_dropout == 0.2 and m._n_messages == 3 and m.readout is not None
    out, a, b = m(_ref_test_data())
    assert out.shape == (1, 2) and a.shape == (3, 64) and b.shape == (5, 64)
    m = OracleMoleculeGCN(10, 5, None, embedding_dim=64, n_messages=3, n_readout=0, readout_dim=128, p_dropout=0.2)
    assert m.readout is None
    out, a, b = m(_ref_test_data())
    assert out.shape == (1, 64) and a.shape == (3, 64) and b.shape == (5, 64)
    for fn in (F.softplus, F.sigmoid, F.relu, F.tanh, F.leaky_relu):
        m = OracleMoleculeGCN(10, 5, None, embedding_dim=64, n_messages=3, n_readout=0, act_fn=fn, p_dropout=0.2)
        out, a, b = m(_ref_test_data())
        assert out.shape == (1, 64) and a.shape == (3, 64) and b.shape == (5, 64)


def test_state_dict_matches_the_printed_module_tree():
    # examples/predict_cn.ipynb:181-196: Embedding(42,128), Embedding(29,128), GeneralConv(128,128),
    # EdgeConv(Linear(256->128)), readout 3 x Linear(128->128) + Linear(128->1); 124 673 parameters
    m = OracleMoleculeGCN(42, 29, 1, embedding_dim=128, n_messages=3, n_readout=3, readout_dim=128)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert shapes["emb_atom.weight"] == (42, 128) and shapes["emb_bond.weight"] == (29, 128)
    assert shapes["atom_conv.lin_msg.weight"] == (128, 128) and shapes["atom_conv.lin_edge.bias"] == (128,)
    assert shapes["bond_conv.nn.0.weight"] == (128, 256)
    assert shapes["readout.3.0.weight"] == (1, 128)
    assert not any(k.startswith("atom_conv.lin_self") for k in shapes)
    assert sum(p.numel() for p in m.parameters()) == 124673


def test_bond_matrix_quirk_of_gcn_py_163():
    data, _ = random_batch(seed=0, n_graphs=6)
    N, E = data.x.numel(), data.edge_index.size(1)
    assert E > N
    m = OracleMoleculeGCN(12, 7, 1, embedding_dim=32, n_messages=2).eval()
    out, a, b = m(data)
    assert b.shape == (E, 32)
    assert torch.allclose(b[N:], torch.full_like(b[N:], math.log(2.0)))  # softplus(0) rows
    # batched != solo prediction for the same molecule (SURVEY.md 3.3 consequence (i))
    _, graphs = random_batch(seed=0, n_graphs=6)
    solo = m(graphs[1])[0]
    assert not torch.allclose(solo, out[1:2], atol=1e-6)
    # an atom index >= E is an IndexError
    c = Data(x=torch.tensor([2]), edge_index=torch.zeros(2, 0, dtype=torch.long), edge_attr=torch.zeros(0, dtype=torch.long),
             y=torch.zeros(1, 1))
    cc = Data(x=torch.tensor([2, 2]), edge_index=torch.tensor([[0, 1], [1, 0]]), edge_attr=torch.tensor([2, 2]), y=torch.zeros(1, 1))
    with pytest.raises(IndexError):
        m(collate([c, cc]))


def test_amax_backward_splits_ties_evenly():
    src = torch.tensor([[1.0], [1.0], [0.5]], requires_grad=True)
    out = torch.zeros(2, 1).scatter_reduce(0, torch.tensor([[0], [0], [0]]), src, "amax", include_self=False)
    out.sum().backward()
    assert src.grad.flatten().tolist() == [0.5, 0.5, 0.0]
    assert float(out[1]) == 0.0  # rows without contributions stay exactly 0


def test_notebook_training_step_reduces_loss():
    data, _ = random_batch(seed=1, n_graphs=16)
    torch.manual_seed(0)
    m = OracleMoleculeGCN(12, 7, 1, embedding_dim=32, n_messages=2)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    losses = [float(train_step(m, opt, data)) for _ in range(30)]
    assert losses[-1] < losses[0]


def test_collate_golden_encodings():
    """Notebook outputs `encoding_train[0]` pin token numbering, edge order
    and dtypes; collating them pins the node offsets."""
    gold = json.load(open(os.path.join(HERE, "golden", "encodings.json")))
    graphs = []
    for name, g in gold.items():
        x = torch.tensor(g["atoms"], dtype=torch.int32)
        ea = torch.tensor(g["bonds"], dtype=torch.int32)
        ei = torch.tensor(g["connectivity"], dtype=torch.int64)
        assert ei.shape == (2, ea.numel()) and int(ei.max()) == x.numel() - 1
        assert (ei[0][1:] >= ei[0][:-1]).all(), name           # source-atom-major
        tok = {(int(s), int(d)): int(t) for s, d, t in zip(ei[0], ei[1], ea)}
        assert all(tok[(d, s)] == t for (s, d), t in tok.items()), name  # one token per bond
        assert min(g["atoms"]) >= 2 and min(g["bonds"]) >= 2   # 'unk' is 1, first real token 2
        graphs.append(Data(x=x, edge_index=ei, edge_attr=ea, y=torch.zeros(1, 1)))
    b = collate(graphs)
    assert b.x.numel() == sum(g.x.numel() for g in graphs)
    off = 0
    for k, g in enumerate(graphs):
        e0 = sum(gr.edge_attr.numel() for gr in graphs[:k])
        assert torch.equal(b.edge_index[:, e0:e0 + g.edge_attr.numel()], g.edge_index + off)
        off += g.x.numel()
    m = OracleMoleculeGCN(8, 8, 1, embedding_dim=16, n_messages=2).eval()
    assert m(b)[0].shape == (4, 1)


def test_oracle_matches_real_pyg():
    """Only runs where torch_geometric is importable (not in this image): bit-for-bit comparison
    of the restated EdgeConv / GeneralConv / global_add_pool with the real ones."""
    tg = pytest.importorskip("torch_geometric")
    from torch_geometric.nn import EdgeConv, GeneralConv, global_add_pool
    from oracle import pyg_restated as R
    data, _ = random_batch(seed=2, n_graphs=5)
    torch.manual_seed(0)
    real_e = EdgeConv(torch.nn.Sequential(torch.nn.Linear(32, 16)))
    mine_e = R.EdgeConv(torch.nn.Sequential(torch.nn.Linear(32, 16)))
    mine_e.load_state_dict(real_e.state_dict())
    xb = torch.randn(data.edge_index.size(1), 16)
    assert torch.equal(real_e(xb, data.edge_index), mine_e(xb, data.edge_index))
    real_g = GeneralConv(16, 16, 16)
    mine_g = R.GeneralConv(16, 16, 16)
    mine_g.load_state_dict(real_g.state_dict())
    xa = torch.randn(data.x.numel(), 16)
    assert torch.equal(real_g(xa, data.edge_index, xb), mine_g(xa, data.edge_index, xb))
    assert torch.equal(global_add_pool(xa, data.batch), R.global_add_pool(xa, data.batch))


@pytest.mark.parametrize("name", ["gcn_small.pt", "gcn_small_max_meanpool.pt"])
def test_oracle_reproduces_committed_golden_vectors(name):
    """tests/golden/gcn_small*.pt (made by tests/golden/make_gcn_golden.py): the oracle must keep producing the
    committed forward outputs, gradients and Adam trajectory (guards against drift of the checker itself)."""
    from tests.golden.make_gcn_golden import build
    gold = torch.load(os.path.join(HERE, "golden", name), weights_only=False)
    now = build(gold.get("extra"))
    for k in ("x", "edge_index", "edge_attr", "batch"):
        assert torch.equal(gold[k], now[k]), k
    for k in ("out", "out_atom", "out_bond"):
        assert torch.allclose(gold[k], now[k], rtol=1e-5, atol=1e-6), k
    assert abs(gold["loss"] - now["loss"]) <= 1e-5 * max(1.0, abs(gold["loss"]))
    for k, v in gold["grads"].items():
        assert torch.allclose(v, now["grads"][k], rtol=1e-4, atol=1e-6), k
    for k, v in gold["state_after_3_adam_steps"].items():
        assert torch.allclose(v, now["state_after_3_adam_steps"][k], rtol=1e-4, atol=1e-6), k


@pytest.mark.parametrize("act", [F.softplus, F.tanh, F.relu])
def test_fp32_oracle_is_within_its_noise_floor_of_the_fp64_oracle(act):
    """the float64 evaluation of the restatement separates algorithmic from rounding error.
    The fp32 restatement (the reference's own arithmetic) must sit within a few 1e-6 of it -- which is why the
    -m gpu parity tests compare the CUDA path with the fp64 evaluation at the north star's 1e-5 / 1e-4 bars."""
    from tests.util import oracle_fp64
    data, _ = random_batch(seed=11, n_graphs=8)
    torch.manual_seed(0)
    o32 = OracleMoleculeGCN(12, 7, 2, embedding_dim=128, n_messages=4, act_fn=act)
    o64 = oracle_fp64(o32)
    r32, r64 = o32(data), o64(data)
    assert all(t.dtype == torch.float64 for t in r64)
    for a, b in zip(r32, r64):
        assert nerr(a, b) <= 5e-6
    F.mse_loss(r32[0], data.y.expand_as(r32[0])).backward()
    F.mse_loss(r64[0], data.y.double().expand_as(r64[0])).backward()
    for (name, p), q in zip(o32.named_parameters(), o64.parameters()):
        assert nerr(p.grad, q.grad) <= 2e-5, name


def test_global_mean_pool_restated():
    """nn/pool/glob.py::global_mean_pool: segment sum / max(count, 1); batch=None averages all rows"""
    from oracle.pyg_restated import global_mean_pool
    x = torch.arange(12.0).view(6, 2)
    batch = torch.tensor([0, 0, 2, 2, 2, 3])
    out = global_mean_pool(x, batch, size=5)
    want = torch.stack([x[:2].mean(0), torch.zeros(2), x[2:5].mean(0), x[5], torch.zeros(2)])
    assert torch.equal(out, want)
    assert torch.equal(global_mean_pool(x, None), x.mean(0, keepdim=True))


def test_general_conv_max_restated():
    """GeneralConv(aggr="max"): per-edge messages, zero-initialised amax (rows without in-edges stay 0) + x"""
    from oracle.pyg_restated import GeneralConv
    torch.manual_seed(0)
    conv = GeneralConv(4, 4, 4, aggr="max")
    x, ea = torch.randn(3, 4), torch.randn(4, 4)
    ei = torch.tensor([[0, 1, 1, 2], [1, 0, 0, 0]])
    out = conv(x, ei, ea)
    msg = conv.lin_msg(x[ei[0]]) + conv.lin_edge(ea)
    want = torch.stack([msg[1:4].max(0).values, msg[0], torch.zeros(4)]) + x
    assert torch.allclose(out, want, atol=1e-6)
