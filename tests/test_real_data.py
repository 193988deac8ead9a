"""BASELINE.json configs[0] (CN, default dims) and configs[1] (YSI, embedding_dim=128, n_messages=3, fp32, "bit-exact
graph batching vs reference") on the REAL bundled molecules: SMILES from the committed fixture
(tests/golden/fuel_sets.json) through the built-in featuriser (tests/test_smiles.py pins it to the notebooks).

CPU part: collate and packed index structures against the restated PyG collate / stable-argsort oracle, bit for bit.
GPU part: forward, gradients and the first Adam steps of the notebook loop against the CPU oracle, fp32."""
import functools

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from graphchem_b200.data import Batch, MoleculeStore, pack_batch
from oracle.csr_oracle import build_index
from oracle.pyg_restated import Data as OData, collate as ocollate
from tests.util import assert_adam_states_close, encoded_fuel_set, make_pair, nerr, oracle_fp64

NOTEBOOK_DIMS = dict(embedding_dim=128, n_messages=3, n_readout=3, readout_dim=128)   # predict_ysi.ipynb:219-222
DEFAULT_DIMS = dict(embedding_dim=128, n_messages=2, n_readout=2, readout_dim=64)     # gcn.py:36-39


@functools.lru_cache(maxsize=None)
def _set(name):
    return encoded_fuel_set(name)


def _odata(graphs, idx):
    return ocollate([OData(x=graphs[i].x, edge_index=graphs[i].edge_index, edge_attr=graphs[i].edge_attr, y=graphs[i].y)
                     for i in idx])


@pytest.mark.parametrize("name,idx", [("ysi", range(0, 16)), ("ysi", range(432, 446)), ("ysi", range(446)),
                                      ("cn", range(16, 32)), ("cn", range(368))])
def test_real_molecules_collate_and_pack_bit_exact(name, idx):
    enc, train, _ = _set(name)
    av, bv = enc.vocab_sizes
    idx = list(idx)
    ref = _odata(train, idx)
    b = Batch.from_data_list([train[i] for i in idx])
    for k in ("x", "edge_index", "edge_attr", "y", "batch", "ptr"):
        assert torch.equal(getattr(b, k), getattr(ref, k)) and getattr(b, k).dtype == getattr(ref, k).dtype, k
    assert b.num_graphs == len(idx)
    pb = pack_batch(b, 3, av, bv, pin=False)
    assert (pb.N, pb.E, pb.G, pb.M) == (ref.x.numel(), ref.edge_index.size(1), len(idx), min(ref.x.numel(), ref.edge_index.size(1)))
    want = build_index(ref.x.numpy(), ref.edge_attr.numpy(), ref.edge_index.numpy(), ref.batch.numpy(), len(idx), av, bv, pb.M)
    for k, v in pb.sections().items():
        assert np.array_equal(v.numpy(), want[k]), k
    pc = MoleculeStore.from_graphs(train).collate_pack(np.asarray(idx), 3, av, bv, pin=False)
    assert torch.equal(pc.ints, pb.ints) and torch.equal(pc.y, pb.y)
    assert pb.n_tiles > 0                      # molecules are far smaller than a 128-row tile


def _forward_backward_parity(name, idx, dims, fwd_tol=1e-5, grad_tol=1e-4):
    from graphchem_b200.data import Data
    enc, train, _ = _set(name)
    av, bv = enc.vocab_sizes
    data = _odata(train, list(idx))
    oracle, model = make_pair(av, bv, 1, seed=3, **dims)
    dev = Data(x=data.x.cuda(), edge_index=data.edge_index.cuda(), edge_attr=data.edge_attr.cuda(), batch=data.batch.cuda(),
               y=data.y.cuda(), num_graphs=data.num_graphs)
    out, out_atom, out_bond = model(dev)
    loss = F.mse_loss(out, dev.y)
    loss.backward()
    # the float64 evaluation of the oracle: the comparison measures the CUDA path's own fp32 error (tests/util.py)
    o64 = oracle_fp64(oracle)
    o64.zero_grad()
    o, oa, ob = o64(data)
    lo = F.mse_loss(o, data.y.double())
    lo.backward()
    errs = [nerr(out, o), nerr(out_atom, oa), nerr(out_bond, ob)]
    gerr = max(nerr(pm.grad, po.grad) for po, pm in zip(o64.parameters(), model.parameters()))
    assert max(errs) <= fwd_tol, errs
    assert abs(float(loss) - float(lo)) <= 1e-5 * abs(float(lo))
    assert gerr <= grad_tol, gerr


@pytest.mark.gpu
@pytest.mark.parametrize("idx", [range(0, 16), range(430, 446), range(446)])
def test_config1_ysi_fp32_forward_and_gradients(idx):
    _forward_backward_parity("ysi", idx, NOTEBOOK_DIMS)


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [DEFAULT_DIMS, NOTEBOOK_DIMS])
def test_config0_cn_fp32_forward_and_gradients(dims):
    _forward_backward_parity("cn", range(0, 16), dims)
    _forward_backward_parity("cn", range(368), dims)


@pytest.mark.gpu
def test_config0_cn_notebook_loop_first_steps_and_per_molecule_prediction():
    """predict_cn.ipynb:247-251 for the first batches (unshuffled), fused TrainStep vs the oracle's zero_grad / forward /
    mse / backward / Adam.step; then the notebook's evaluation `[model(mol)[0] for mol in ds]` (:333-334) against
    predict_each, on weights that have moved."""
    from graphchem_b200.train import TrainStep
    from oracle.gcn_oracle import train_step
    enc, train, test = _set("cn")
    av, bv = enc.vocab_sizes
    oracle, model = make_pair(av, bv, 1, seed=5, **NOTEBOOK_DIMS)
    model.train(); oracle.train()
    opt = torch.optim.Adam(oracle.parameters(), lr=1e-3)
    step = TrainStep(model, lr=1e-3, use_cuda_graph=False)
    store = MoleculeStore.from_graphs(train)
    for k in range(4):
        idx = list(range(16 * k, 16 * k + 16))
        lo = float(train_step(oracle, opt, _odata(train, idx)))
        lc = float(step(store.collate_pack(np.asarray(idx), 3, av, bv).to("cuda")).item())
        assert abs(lc - lo) <= 2e-4 * abs(lo), (k, lc, lo)      # CN targets are 0..100: losses in the hundreds
    for name, v in model.state_dict().items():
        assert_adam_states_close(v, oracle.state_dict()[name], 1e-3, 4, name)
    model.eval(); oracle.eval()
    with torch.no_grad():
        want = torch.cat([oracle(OData(x=g.x, edge_index=g.edge_index, edge_attr=g.edge_attr))[0] for g in test[:24]])
    got = model.predict_each(test[:24])
    assert nerr(got, want) <= 5e-4
