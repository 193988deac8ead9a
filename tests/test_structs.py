"""-m "not gpu": data contract of reference graphchem/data/structs.py (behaviour of the
reference's tests/test_structs.py:8-104 re-checked against this package)."""
import pytest
import torch

from graphchem_b200.data import Batch, Data, DataLoader, MoleculeDataset, MoleculeGraph


def _graph(target=None):
    return MoleculeGraph(torch.randint(0, 9, (10,)), torch.randint(0, 4, (20,)), torch.randint(0, 10, (2, 20)), target)


def test_target_normalisation():
    g = _graph()
    assert g.x.shape == (10,) and g.edge_index.shape == (2, 20) and g.edge_attr.shape == (20,)
    assert g.y.shape == (1, 1) and g.y.dtype == torch.float32 and float(g.y) == 0.0
    assert torch.allclose(_graph(torch.tensor([2.5])).y, torch.tensor([[2.5]]))
    assert _graph(torch.rand(1, 3)).y.shape == (1, 3)
    assert _graph(torch.rand(3)).y.shape == (1, 3)
    with pytest.raises(ValueError):
        _graph(torch.rand(2, 3))


def test_dataset_protocol():
    assert len(MoleculeDataset([])) == 0
    ds = MoleculeDataset([_graph() for _ in range(5)])
    assert len(ds) == 5 and ds.len() == 5
    assert isinstance(ds[0], MoleculeGraph) and ds.get(4) is ds[4]
    with pytest.raises(IndexError):
        _ = ds[5]
    assert len(ds[1:3]) == 2 and len(ds[[0, 4]]) == 2


def test_data_properties_and_batch_none():
    g = _graph()
    assert g.batch is None and g.num_nodes == 10 and g.num_node_features == 1 and g.num_edges == 20
    assert Data().num_node_features == 0
    with pytest.raises(AttributeError):
        _ = g.nonexistent


def test_dataloader_collates_like_pyg():
    ds = MoleculeDataset([_graph(torch.tensor([float(i)])) for i in range(7)])
    batches = list(DataLoader(ds, batch_size=3, shuffle=False))
    assert [b.num_graphs for b in batches] == [3, 3, 1]  # drop_last=False
    b = batches[0]
    assert b.x.shape == (30,) and b.edge_index.shape == (2, 60) and b.y.shape == (3, 1)
    assert b.batch.tolist() == [0] * 10 + [1] * 10 + [2] * 10 and b.ptr.tolist() == [0, 10, 20, 30]
    assert int(b.edge_index[:, 20:40].min()) >= 10 and int(b.edge_index[:, 40:].min()) >= 20
    assert isinstance(b, Batch)
