"""-m gpu: batched per-molecule prediction (`MoleculeGCN.predict_each`, GCB_PACK_EACH) against the loop it
replaces -- `model(g)` one molecule at a time, reference examples/predict_cn.ipynb:258-263, 333-334 -- and
against the CPU oracle run one molecule at a time."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import make_pair, nerr, random_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_predict_each_equals_per_molecule_loop(dtype, tol):
    from graphchem_b200.data import MoleculeGraph
    data, graphs = random_batch(seed=4, n_graphs=150, max_atoms=14)
    mgs = [MoleculeGraph(g.x, g.edge_attr, g.edge_index, g.y) for g in graphs]
    oracle, model = make_pair(12, 7, 2, dtype=dtype, embedding_dim=128, n_messages=3, n_readout=2, readout_dim=64)
    oracle.eval(); model.eval()
    out, atoms, bonds = model.predict_each(mgs, batch_size=64, return_all=True)
    assert out.shape == (150, 2) and len(atoms) == len(bonds) == 150
    with torch.no_grad():
        batched = model(data)[0].float().cpu()
        for i in (0, 1, 7, 63, 64, 149):  # both sides of the chunk boundary
            po, ao, bo = oracle(graphs[i])
            pm, am, bm = model(mgs[i])
            assert atoms[i].shape == ao.shape and bonds[i].shape == bo.shape
            # the loop the notebooks run, on this package
            assert nerr(out[i:i + 1].float(), pm.float()) <= (1e-6 if dtype == torch.float32 else 1e-2)
            assert nerr(atoms[i].float(), am.float()) <= (1e-6 if dtype == torch.float32 else 1e-2)
            assert nerr(bonds[i].float(), bm.float()) <= (1e-6 if dtype == torch.float32 else 1e-2)
            # the reference semantics, one molecule alone
            # (a single molecule's prediction is a couple of small numbers: norm-wise error is loose in bf16)
            assert nerr(atoms[i].float(), ao) <= tol and nerr(bonds[i].float(), bo) <= tol
            assert nerr(out[i:i + 1].float(), po) <= (tol if dtype == torch.float32 else 5 * tol)
    # and the point of it all: a plain batched forward is NOT the per-molecule result
    assert nerr(batched, out.float()) > 1e-3


@pytest.mark.parametrize("kw", [dict(aggr="max"), dict(aggr="mean"), dict(pool="mean")])
def test_predict_each_other_aggregations_and_pooling(kw):
    """max / mean aggregation and mean pooling through the each-molecule layout: equal to the oracle one molecule at a time"""
    from graphchem_b200.data import MoleculeGraph
    _, graphs = random_batch(seed=6, n_graphs=40, max_atoms=12)
    mgs = [MoleculeGraph(g.x, g.edge_attr, g.edge_index, g.y) for g in graphs]
    oracle, model = make_pair(12, 7, 2, embedding_dim=64, n_messages=2, n_readout=1, readout_dim=32, **kw)
    oracle.eval(); model.eval()
    out, atoms, bonds = model.predict_each(mgs, batch_size=16, return_all=True)
    with torch.no_grad():
        for i in (0, 15, 16, 39):
            po, ao, bo = oracle(graphs[i])
            assert nerr(out[i:i + 1], po) <= 1e-5 and nerr(atoms[i], ao) <= 1e-5 and nerr(bonds[i], bo) <= 1e-5


def test_each_batches_are_inference_only():
    from graphchem_b200.data import MoleculeGraph, MoleculeStore
    _, graphs = random_batch(seed=5, n_graphs=8)
    store = MoleculeStore.from_graphs([MoleculeGraph(g.x, g.edge_attr, g.edge_index, g.y) for g in graphs])
    _, model = make_pair(12, 7, 1, embedding_dim=64, n_messages=2)
    pb = store.collate_pack(list(range(8)), 2, 12, 7, each=True).to("cuda")
    model.train()
    with pytest.raises(ValueError):
        model(pb)  # gradients requested on a GCB_PACK_EACH batch
    model.eval()
    with torch.no_grad():
        assert model(pb)[0].shape == (8, 1)


def test_predict_each_with_lone_atoms_and_unbonded_fragments():
    """'C' (one atom, no bond, tests/test_features.py:146-155 shape) and a salt-like molecule with unbonded atoms
    have fewer edges than atoms; the reference's `model(g)` handles each of them alone, so must predict_each."""
    from graphchem_b200.data import MoleculeGraph
    _, graphs = random_batch(seed=9, n_graphs=6, max_atoms=9)
    lone = MoleculeGraph(torch.tensor([3], dtype=torch.int32), torch.zeros(0, dtype=torch.int32),
                         torch.zeros(2, 0, dtype=torch.long), torch.zeros(1, 1))
    # three atoms, one bond between atoms 0 and 1 (edge ids 0, 1 < E = 2), atom 2 unbonded
    salt = MoleculeGraph(torch.tensor([2, 4, 5], dtype=torch.int32), torch.tensor([3, 3], dtype=torch.int32),
                         torch.tensor([[0, 1], [1, 0]]), torch.zeros(1, 1))
    mgs = [MoleculeGraph(g.x, g.edge_attr, g.edge_index, g.y) for g in graphs]
    mols = mgs[:2] + [lone] + mgs[2:4] + [salt, lone] + mgs[4:]
    oracle, model = make_pair(12, 7, 1, embedding_dim=64, n_messages=2)
    oracle.eval(); model.eval()
    out, atoms, bonds = model.predict_each(mols, batch_size=4, return_all=True)
    assert out.shape == (len(mols), 1)
    with torch.no_grad():
        for i, g in enumerate(mols):
            po, ao, bo = oracle(g)
            assert atoms[i].shape == ao.shape and bonds[i].shape == bo.shape, i
            assert nerr(out[i:i + 1], po) <= 1e-5 and nerr(atoms[i], ao) <= 1e-5 and nerr(bonds[i], bo) <= 1e-5, i
    # only lone atoms: the chunk that round 1 rejected with NotImplementedError
    assert model.predict_each([lone, lone, lone]).shape == (3, 1)
