"""-m gpu: PackedLoader on a CUDA device (recycled pinned buffers, copy stream, prefetch) delivers exactly the
batches a direct collate_pack gives, across epochs; TrainStep consumes them."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_cuda_loader_batches_are_exact_and_recycled():
    from graphchem_b200.data import PackedLoader
    from graphchem_b200.synthetic import make_molecules
    store = make_molecules(1000, seed=5, variable=True)
    loader = PackedLoader(store, 128, 4, 64, 32, device="cuda", workers=4, prefetch=2)
    for epoch in range(3):  # later epochs run on recycled pinned buffers
        n = 0
        for k, b in enumerate(loader):
            ref = store.collate_pack(np.arange(128 * k, min(1000, 128 * k + 128)), 4, 64, 32, pin=False)
            assert b.ints.is_cuda and torch.equal(b.ints.cpu(), ref.ints) and torch.equal(b.y.cpu(), ref.y)
            assert (b.header == ref.header).all()
            n += 1
        assert n == 8
    # pinned pairs: free in the pool, or held by the batches already packed ahead for the next epoch -- exactly the
    # `prefetch + workers` pairs of the first epoch, none added since
    held = [f.result() for f in loader._ahead]
    assert len(loader._pool) + len(held) == 4 + 2 and loader._pool_ready


def test_trainstep_over_loader_learns():
    from graphchem_b200.data import PackedLoader
    from graphchem_b200.nn import MoleculeGCN
    from graphchem_b200.synthetic import make_molecules
    from graphchem_b200.train import TrainStep
    store = make_molecules(512, seed=6)
    torch.manual_seed(0)
    model = MoleculeGCN(64, 32, 1, embedding_dim=128, n_messages=2, compute_dtype=torch.bfloat16).cuda()
    step = TrainStep(model, lr=1e-3, use_cuda_graph=False)
    loader = PackedLoader(store, 128, 2, 64, 32, device="cuda", workers=2, cache=True, shuffle=True)
    first = last = None
    for epoch in range(12):
        tot = 0.0
        for pb in loader:
            tot += float(step(pb).item())
        first = tot if first is None else first
        last = tot
    assert last < first


def test_device_slot_ring_lets_trainstep_replay_graphs():
    """device_slots: batches land in a fixed ring of device buffers, TrainStep captures one CUDA graph per slot
    and replays it; contents still equal a direct pack and training equals the un-graphed run."""
    from graphchem_b200.data import PackedLoader
    from graphchem_b200.nn import MoleculeGCN
    from graphchem_b200.synthetic import make_molecules
    from graphchem_b200.train import TrainStep
    store = make_molecules(1024, seed=8)  # fixed-shape molecules: every batch has the same header
    loader = PackedLoader(store, 128, 2, 64, 32, device="cuda", workers=3, prefetch=4, device_slots=3)
    seen = set()
    for k, b in enumerate(loader):
        ref = store.collate_pack(np.arange(128 * k, 128 * k + 128), 2, 64, 32, pin=False)
        assert torch.equal(b.ints.cpu(), ref.ints) and torch.equal(b.y.cpu(), ref.y)
        seen.add(b.ints.data_ptr())
    assert len(seen) == 3

    def run(graph):
        torch.manual_seed(0)
        model = MoleculeGCN(64, 32, 1, embedding_dim=128, n_messages=2, compute_dtype=torch.bfloat16).cuda()
        step = TrainStep(model, lr=1e-3, use_cuda_graph=graph)
        losses = []
        for epoch in range(3):
            for pb in loader:
                losses.append(step(pb).clone())
        torch.cuda.synchronize()
        return torch.stack(losses).cpu(), (len(step._graphs) if graph else 0)

    l_graph, n_graphs = run(True)
    l_plain, _ = run(False)
    assert n_graphs == 3
    assert torch.equal(l_graph, l_plain)


def test_inferstep_matches_module_forward():
    """InferStep (graphed forward, no saved activations) returns exactly what model.eval()(batch) returns."""
    from graphchem_b200.nn import MoleculeGCN
    from graphchem_b200.synthetic import make_molecules
    from graphchem_b200.train import InferStep
    store = make_molecules(600, seed=9, variable=True)
    torch.manual_seed(1)
    model = MoleculeGCN(64, 32, 2, embedding_dim=128, n_messages=3, compute_dtype=torch.bfloat16).cuda().eval()
    pbs = [store.collate_pack(np.arange(i, i + 200), 3, 64, 32).to("cuda") for i in (0, 200, 400)]
    step = InferStep(model, want_atom_bond=True)
    for rep in range(2):  # second round replays the graphs
        for pb in pbs:
            out, oa, ob = (t.clone() for t in step(pb))
            with torch.no_grad():
                ref = model(pb)
            assert torch.equal(out, ref[0]) and torch.equal(oa, ref[1]) and torch.equal(ob, ref[2])
    assert len(step._graphs) == 3


def test_streamed_variable_shape_batches_are_not_recycled_under_the_consumer():
    """Real CN molecules (every shuffled batch of 16 has its own N and E) streamed through the loader WITHOUT a
    slot ring: a batch is allocated on the copy stream and read on the consumer's stream, so it must stay alive
    until the consumer's kernels are done with it (record_stream).  Losses of the streamed run must equal, bit for
    bit, those of the same batches packed and stepped one at a time."""
    from graphchem_b200.data import MoleculeStore, PackedLoader
    from graphchem_b200.nn import MoleculeGCN
    from graphchem_b200.train import TrainStep
    from tests.util import encoded_fuel_set
    enc, train, _ = encoded_fuel_set("cn")
    av, bv = enc.vocab_sizes
    store = MoleculeStore.from_graphs(train)
    kw = dict(embedding_dim=128, n_messages=3, n_readout=3, readout_dim=128)

    def run(streamed):
        torch.manual_seed(0)
        model = MoleculeGCN(av, bv, 1, **kw).cuda().train()
        step = TrainStep(model, lr=1e-3, use_cuda_graph=False)
        loader = PackedLoader(store, 16, 3, av, bv, device="cuda", shuffle=True, seed=3, workers=2, prefetch=4)
        losses = []
        for epoch in range(6):
            if streamed:
                losses += [step(pb).clone() for pb in loader]
            else:
                loader.epoch += 1
                for idx in loader._batches():
                    pb = store.collate_pack(idx, 3, av, bv).to("cuda")
                    losses.append(step(pb).clone())
                    torch.cuda.synchronize()
        torch.cuda.synchronize()
        return torch.stack(losses).cpu()

    a, b = run(True), run(False)
    assert a.numel() == 6 * 23
    assert torch.isfinite(a).all()
    assert torch.equal(a, b)


@pytest.mark.parametrize("variable", [False, True])
def test_screen_streams_the_batched_forward(variable):
    """train.screen (BASELINE configs[4]: forward only, device-packed batches, graphed when the shapes repeat) returns,
    in order, what model.eval()(batch) returns on the same batches; with world=2 every rank screens its own block."""
    from graphchem_b200.nn import MoleculeGCN
    from graphchem_b200.synthetic import make_molecules
    from graphchem_b200.train import screen, screening_shard
    store = make_molecules(700, seed=9, variable=variable)
    torch.manual_seed(1)
    model = MoleculeGCN(64, 32, 2, embedding_dim=128, n_messages=3, compute_dtype=torch.bfloat16).cuda()

    def reference(idx, bs):
        model.eval()
        with torch.no_grad():
            outs = [model(store.collate_pack(idx[i:i + bs], 3, 64, 32).to("cuda"))[0].cpu() for i in range(0, len(idx), bs)]
        model.train()
        return torch.cat(outs)

    got = screen(model, store, batch_size=256, workers=2)
    assert model.training and got.shape == (700, 2) and not got.is_cuda
    assert torch.equal(got, reference(np.arange(700), 256))
    for r in range(2):
        part = screen(model, store, batch_size=128, rank=r, world=2, workers=2, device_pack=bool(r))
        assert torch.equal(part, reference(screening_shard(700, r, 2), 128))


def test_lookahead_never_allocates_pinned_buffers_mid_epoch():
    """The pinned pool is sized for the pipeline that stays `prefetch + workers` deep across epoch boundaries: after the
    first epoch no worker may find it empty (cudaHostAlloc synchronises the device and invalidates a CUDA graph capture
    running on the consumer's thread -- seen as cudaErrorStreamCaptureInvalidated in bench.py's streaming leg)."""
    import time
    from graphchem_b200.data import PackedLoader
    from graphchem_b200.synthetic import make_molecules
    store = make_molecules(3 * 256, seed=5)
    loader = PackedLoader(store, 256, 2, 64, 32, device="cuda", workers=4, prefetch=2, shuffle=True, seed=1)
    seen = set()

    def buffers():
        time.sleep(0.2)  # let the batches packed ahead finish
        held = [f.result()[1][0].data_ptr() for f in loader._ahead]
        with loader._pool_lock:
            return set(held) | {b[0].data_ptr() for b in loader._pool}

    for _ in loader:
        pass
    seen = buffers()
    assert len(seen) == loader.prefetch + loader.workers
    for _ in range(4):
        for _ in loader:
            pass
        assert buffers() == seen, "a worker allocated a new pinned buffer"
    loader.close()
