"""-m "not gpu": the host packer / collate (C-ABI, pure integer) against the numpy oracle, bit for bit."""
import numpy as np
import pytest
import torch

from graphchem_b200.data import Batch, Data, MoleculeGraph, MoleculeStore, pack_batch
from oracle.csr_oracle import build_index
from oracle.pyg_restated import collate as ocollate
from tests.util import random_batch, random_graph_batch


def _assert_matches_oracle(pb, data, G, av, bv):
    batch = None if getattr(data, "batch", None) is None else data.batch.numpy()
    ref = build_index(data.x.numpy(), data.edge_attr.numpy(), data.edge_index.numpy(), batch, G, av, bv, pb.M)
    for k, v in pb.sections().items():
        assert v.dtype == torch.int32
        assert np.array_equal(v.numpy(), ref[k]), k


@pytest.mark.parametrize("seed,G", [(0, 1), (1, 16), (2, 64)])
def test_pack_molecule_batches_bit_exact(seed, G):
    data, _ = random_batch(seed, G)
    pb = pack_batch(data, 2, 12, 7, pin=False)
    assert (pb.N, pb.E, pb.G) == (data.x.numel(), data.edge_index.size(1), G)
    assert pb.M == min(pb.N, pb.E)
    _assert_matches_oracle(pb, data, G, 12, 7)
    pb0 = pack_batch(data, 0, 12, 7, pin=False)
    assert pb0.M == pb0.E
    _assert_matches_oracle(pb0, data, G, 12, 7)


def test_pack_random_graph_with_duplicates_and_unsorted_batch():
    data = random_graph_batch(seed=3, N=50, E=300, av=10, bv=5, G=1)
    g = torch.Generator().manual_seed(0)
    data.batch = torch.randint(0, 7, (50,), generator=g)  # unsorted on purpose
    data.num_graphs = 7
    pb = pack_batch(data, 3, 10, 5, pin=False)
    _assert_matches_oracle(pb, data, 7, 10, 5)


def test_pack_empty_edges_and_batch_none():
    d = Data(x=torch.tensor([3], dtype=torch.int32), edge_index=torch.zeros(2, 0, dtype=torch.long),
             edge_attr=torch.zeros(0, dtype=torch.int32))
    pb = pack_batch(d, 2, 12, 7, pin=False)
    assert (pb.N, pb.E, pb.G, pb.M) == (1, 0, 1, 0)
    s = pb.sections()
    assert s["in_ptr"].tolist() == [0, 0] and s["graph_ptr"].tolist() == [0, 1] and s["graph_node"].tolist() == [0]


def test_pack_errors_follow_the_reference():
    # atom index >= E with n_messages >= 1: IndexError (reference gcn.py:163 via PyG index_select)
    d = Data(x=torch.tensor([2, 2, 2]), edge_index=torch.tensor([[1, 2], [2, 1]]), edge_attr=torch.tensor([2, 2]))
    with pytest.raises(IndexError):
        pack_batch(d, 2, 12, 7, pin=False)
    pack_batch(d, 0, 12, 7, pin=False)  # n_messages == 0 never indexes the bond matrix
    # token out of range: nn.Embedding IndexError
    d = Data(x=torch.tensor([2, 12]), edge_index=torch.tensor([[0, 1], [1, 0]]), edge_attr=torch.tensor([2, 2]))
    with pytest.raises(IndexError):
        pack_batch(d, 2, 12, 7, pin=False)
    # PyG _check_input: shape / dtype of edge_index -> ValueError
    d = Data(x=torch.tensor([2, 3]), edge_index=torch.tensor([[0, 1, 1]]), edge_attr=torch.tensor([2, 2, 2]))
    with pytest.raises(ValueError):
        pack_batch(d, 2, 12, 7, pin=False)
    d = Data(x=torch.tensor([2, 3]), edge_index=torch.tensor([[0., 1.], [1., 0.]]), edge_attr=torch.tensor([2, 2]))
    with pytest.raises(ValueError):
        pack_batch(d, 2, 12, 7, pin=False)


def test_collate_matches_pyg_restatement_and_c_collate():
    data, graphs = random_batch(seed=5, n_graphs=33)
    mgs = [MoleculeGraph(g.x, g.edge_attr, g.edge_index, g.y) for g in graphs]
    b = Batch.from_data_list(mgs)
    for k in ("x", "edge_index", "edge_attr", "y", "batch", "ptr"):
        assert torch.equal(getattr(b, k), getattr(data, k)), k
        assert getattr(b, k).dtype == getattr(data, k).dtype
    assert b.num_graphs == 33 and b.x.dtype == torch.int32 and b.edge_index.dtype == torch.int64
    store = MoleculeStore.from_graphs(mgs)
    pb = store.collate_pack(np.arange(33), 2, 12, 7, pin=False)
    ref = pack_batch(b, 2, 12, 7, pin=False)
    assert torch.equal(pb.ints, ref.ints) and torch.equal(pb.y, ref.y)
    # a shuffled subset, as DataLoader(shuffle=True) would draw it
    idx = np.random.default_rng(0).permutation(33)[:16]
    pb = store.collate_pack(idx, 2, 12, 7, pin=False)
    ref = pack_batch(Batch.from_data_list([mgs[i] for i in idx]), 2, 12, 7, pin=False)
    assert torch.equal(pb.ints, ref.ints) and torch.equal(pb.y, ref.y)


def test_synthetic_generator_is_encoder_shaped():
    from graphchem_b200.synthetic import make_molecules
    store = make_molecules(64, seed=1)
    assert len(store) == 64 and (store.n_atoms == 23).all() and (store.n_edges == 50).all()
    conn = store.conn[:100].view(2, 50)
    assert (conn[0][1:] >= conn[0][:-1]).all(), "edges must be grouped by source atom (features.py:307-318)"
    fwd = {(int(s), int(d)): int(t) for s, d, t in zip(conn[0], conn[1], store.bonds[:50])}
    assert all(fwd[(d, s)] == t for (s, d), t in fwd.items()), "both directions of a bond share one token"
    assert len(fwd) == 50
    var = make_molecules(200, seed=2, variable=True)
    assert var.n_atoms.min() >= 6 and var.n_atoms.max() <= 38
    pb = var.collate_pack(np.arange(200), 4, 64, 32, pin=False)
    assert pb.G == 200 and pb.E >= pb.N


def test_closed_row_tiles():
    """tile_ptr: consecutive tiles of <= 128 rows that no edge leaves (molecules never split), matching the
    numpy oracle; a component wider than a tile leaves n_tiles == 0 (untiled kernels)."""
    from graphchem_b200 import _lib
    data, graphs = random_batch(seed=9, n_graphs=300, max_atoms=30)
    pb = pack_batch(data, 3, 12, 7, pin=False)
    tp = pb.sections()["tile_ptr"].numpy()
    assert pb.n_tiles == len(tp) - 1 > 1 and tp[0] == 0 and tp[-1] == pb.N
    assert (np.diff(tp) > 0).all() and (np.diff(tp) <= _lib.TILE_ROWS).all()
    tile_of = np.searchsorted(tp, np.arange(pb.N), side="right") - 1
    ei = data.edge_index.numpy()
    assert (tile_of[ei[0]] == tile_of[ei[1]]).all(), "an edge leaves its tile"
    # greedy: two neighbouring tiles never fit into one
    ptr = data.ptr.numpy()
    assert set(tp.tolist()) <= set(ptr.tolist()), "tiles must end on molecule boundaries"
    for a, b, c in zip(tp[:-2], tp[1:-1], tp[2:]):
        nxt = ptr[np.searchsorted(ptr, b, side="right")]  # end of the first molecule of the next tile
        assert nxt - a > _lib.TILE_ROWS
    # a 200-atom ring: one component wider than a tile
    n = 200
    s = torch.arange(n)
    ring = Data(x=torch.full((n,), 3), edge_index=torch.stack([torch.cat([s, (s + 1) % n]), torch.cat([(s + 1) % n, s])]),
                edge_attr=torch.full((2 * n,), 2))
    pr = pack_batch(ring, 2, 12, 7, pin=False)
    assert pr.n_tiles == 0
    _assert_matches_oracle(pr, ring, 1, 12, 7)
    # ... next to small molecules it still poisons the whole batch (documented fallback)
    both = ocollate([graphs[0], type(graphs[0])(x=ring.x.int(), edge_index=ring.edge_index, edge_attr=ring.edge_attr.int(),
                                                y=torch.zeros(1, 1))])
    assert pack_batch(both, 2, 12, 7, pin=False).n_tiles == 0


def test_each_molecule_alone_layout():
    """GCB_PACK_EACH (batched prediction == one forward per molecule): bond rows re-indexed per molecule,
    CSR edge ids redirected to the rows that hold them; everything else as the plain collate."""
    from graphchem_b200 import _lib
    from oracle.csr_oracle import each_layout
    data, graphs = random_batch(seed=13, n_graphs=40, max_atoms=9)
    mgs = [MoleculeGraph(g.x, g.edge_attr, g.edge_index, g.y) for g in graphs]
    store = MoleculeStore.from_graphs(mgs)
    idx = np.arange(40)
    plain = store.collate_pack(idx, 2, 12, 7, pin=False)
    each = store.collate_pack(idx, 2, 12, 7, pin=False, each=True)
    assert each.M == each.N == plain.N and int(each.header[_lib.H_FLAGS]) == _lib.PACK_EACH
    sp, se = plain.sections(), each.sections()
    bond_pos, e_tok_each = each_layout(store.n_atoms, store.n_edges, [g.edge_attr.numpy() for g in graphs])
    assert np.array_equal(se["bond_pos"].numpy(), bond_pos)
    assert np.array_equal(se["e_tok"].numpy()[:each.N], e_tok_each)
    for k in ("x_tok", "src", "dst", "in_ptr", "in_src", "out_ptr", "out_dst", "graph_ptr", "tile_ptr", "in_src4", "out_dst4",
              "out_inpos"):
        assert torch.equal(sp[k], se[k]), k
    redirect = lambda eid: np.where(eid >= 0, np.where(bond_pos[np.maximum(eid, 0)] >= 0, bond_pos[np.maximum(eid, 0)], each.M), -1)
    assert np.array_equal(se["in_eid"].numpy(), redirect(sp["in_eid"].numpy()))
    assert np.array_equal(se["in_eid4"].numpy(), redirect(sp["in_eid4"].numpy()))
    # a molecule whose connectivity addresses a bond row it does not have alone: IndexError, as the reference
    bad = MoleculeGraph(torch.tensor([2, 2, 2], dtype=torch.int32), torch.tensor([2, 2], dtype=torch.int32),
                        torch.tensor([[1, 2], [2, 1]]))
    with pytest.raises(IndexError):
        MoleculeStore.from_graphs([mgs[0], bad]).collate_pack(np.arange(2), 2, 12, 7, pin=False, each=True)


def test_packed_loader_matches_direct_packing():
    """PackedLoader (thread-pool packing, prefetch, resident cache) yields exactly the batches a direct
    collate_pack of the same index slices gives, in order; the cache replays them."""
    from graphchem_b200.data import PackedLoader
    from graphchem_b200.synthetic import make_molecules
    store = make_molecules(300, seed=3, variable=True)
    loader = PackedLoader(store, 64, 4, 64, 32, device="cpu", workers=4, prefetch=2, cache=True)
    assert len(loader) == 5
    got = list(loader)
    assert [b.G for b in got] == [64, 64, 64, 64, 44]
    for k, b in enumerate(got):
        ref = store.collate_pack(np.arange(64 * k, min(300, 64 * k + 64)), 4, 64, 32, pin=False)
        assert torch.equal(b.ints, ref.ints) and torch.equal(b.y, ref.y)
    again = list(loader)
    assert all(a is b for a, b in zip(got, again))  # resident batches are replayed, not re-packed
    sh = PackedLoader(store, 64, 4, 64, 32, device="cpu", shuffle=True, drop_last=True, seed=1, indices=np.arange(0, 300, 2))
    e1 = [b.G for b in sh]
    assert e1 == [64, 64] and len(sh) == 2


def test_packed_loader_host_side_logic_on_cpu():
    """PackedLoader with device='cpu' (no CUDA involved): batch composition, shuffle per epoch, drop_last, shards
    (`indices`), the resident cache and the each-molecule layout all equal direct collate_pack calls."""
    from graphchem_b200.data import PackedLoader
    from tests.util import encoded_fuel_set
    enc, train, _ = encoded_fuel_set("lhv")
    av, bv = enc.vocab_sizes
    store = MoleculeStore.from_graphs(train)
    n = len(store)
    loader = PackedLoader(store, 32, 3, av, bv, device="cpu", workers=3, prefetch=2)
    assert len(loader) == -(-n // 32)
    got = list(loader)
    assert sum(b.G for b in got) == n
    for k, b in enumerate(got):
        ref = store.collate_pack(np.arange(32 * k, min(n, 32 * k + 32)), 3, av, bv, pin=False)
        assert torch.equal(b.ints, ref.ints) and torch.equal(b.y, ref.y)
    # shuffle: a permutation of the data set per epoch, different between epochs, reproducible from the seed
    sh = PackedLoader(store, 32, 3, av, bv, device="cpu", shuffle=True, seed=5, drop_last=True, workers=2)
    assert len(sh) == n // 32
    e1 = torch.cat([b.y for b in sh]); e2 = torch.cat([b.y for b in sh])
    sh2 = PackedLoader(store, 32, 3, av, bv, device="cpu", shuffle=True, seed=5, drop_last=True, workers=1)
    assert torch.equal(e1, torch.cat([b.y for b in sh2])) and not torch.equal(e1, e2)
    assert e1.numel() == (n // 32) * 32
    # shard + cache: the second epoch replays the very same packed batches
    shard = PackedLoader(store, 16, 3, av, bv, device="cpu", indices=np.arange(1, n, 2), cache=True)
    first = list(shard); again = list(shard)
    assert sum(b.G for b in first) == len(range(1, n, 2)) and all(a is b for a, b in zip(first, again))
    ref = store.collate_pack(np.arange(1, n, 2)[:16], 3, av, bv, pin=False)
    assert torch.equal(first[0].ints, ref.ints)
    # each-molecule layout (predict_each)
    each = next(iter(PackedLoader(store, 8, 3, av, bv, device="cpu", each=True)))
    ref = store.collate_pack(np.arange(8), 3, av, bv, pin=False, each=True)
    assert torch.equal(each.ints, ref.ints) and each.M == each.N


def test_packed_loader_lookahead_across_epochs():
    """The persistent worker pool packs the head of epoch e+1 while epoch e drains (`lookahead`): same batches in the
    same order as without it, over several shuffled epochs; a loader left mid-epoch or re-targeted in between still
    yields exactly what a fresh loader would."""
    from graphchem_b200.data import PackedLoader
    from graphchem_b200.synthetic import make_molecules
    store = make_molecules(200, seed=7, variable=True)
    kw = dict(device="cpu", shuffle=True, seed=11, workers=3, prefetch=2)
    a = PackedLoader(store, 32, 2, 64, 32, lookahead=True, **kw)
    b = PackedLoader(store, 32, 2, 64, 32, lookahead=False, **kw)
    for _ in range(3):
        ea, eb = list(a), list(b)
        assert len(ea) == len(eb) == 7
        assert all(torch.equal(x.ints, y.ints) and torch.equal(x.y, y.y) for x, y in zip(ea, eb))
        assert a._ahead_epoch == a.epoch + 1 and 1 <= len(a._ahead) <= 5   # the next epoch is already being packed
        assert not b._ahead
    # leave an epoch early, then a full epoch: still the reference order of that epoch
    it = iter(a)
    next(it); it.close()
    next(iter(b))
    ea, eb = list(a), list(b)
    assert all(torch.equal(x.ints, y.ints) for x, y in zip(ea, eb))
    # re-target the loader between epochs: what was packed ahead is discarded, not served
    a.indices = np.arange(0, 200, 2); b.indices = np.arange(0, 200, 2)
    ea, eb = list(a), list(b)
    assert sum(x.G for x in ea) == 100 and all(torch.equal(x.ints, y.ints) for x, y in zip(ea, eb))
    a.close(); b.close()
    assert [x.G for x in a] == [x.G for x in b]   # usable again after close()
