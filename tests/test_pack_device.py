"""Device-side packer (include/graphchem_b200.h "device-side packer"; SURVEY.md 8 d config 5, 8 f1).

CPU (-m "not gpu"): the host half (`gcb_collate_raw_host`) and, through the test hook
`gcb_debug_pack_emulate_host`, the very routines the kernels run (they are __host__ __device__), bit for bit
against the host packer -- which tests/test_pack.py pins to the stable-argsort oracle.
GPU (-m gpu): `gcb_pack_device` itself, bit for bit against the host packer, and a forward on top of it."""
import numpy as np
import pytest
import torch

from graphchem_b200 import _lib
from graphchem_b200.data import MoleculeStore
from graphchem_b200.synthetic import make_molecules
from tests.util import random_batch

AV, BV = 12, 7


def _mol(x, edges, bonds=None):
    """Molecule from directed edge pairs (src, dst) in the given order."""
    src = [e[0] for e in edges]
    dst = [e[1] for e in edges]
    ei = torch.tensor([src, dst], dtype=torch.int64).reshape(2, len(edges))
    ea = torch.tensor(bonds if bonds is not None else [2 + (k % (BV - 2)) for k in range(len(edges))], dtype=torch.int32)
    return torch.tensor(x, dtype=torch.int32), ea, ei


def _store(mols, targets=True):
    y = torch.arange(len(mols), dtype=torch.float32).reshape(-1, 1) if targets else None
    return MoleculeStore([m[0] for m in mols], [m[1] for m in mols], [m[2] for m in mols], y)


def _ring(n, extra=()):
    e = []
    for i in range(n):
        e += [(i, (i + 1) % n), ((i + 1) % n, i)]
    return _mol([2 + i % (AV - 2) for i in range(n)], e + list(extra))


def _cases():
    rng = np.random.default_rng(0)
    _, graphs = random_batch(5, 40)
    small = MoleculeStore.from_graphs(graphs)
    # disconnected molecules (salts), isolated atoms, a molecule without edges, duplicate edges and a self loop
    odd = _store([
        _mol([2, 3, 4, 5], [(0, 1), (1, 0), (2, 3), (3, 2)]),
        _mol([2, 3, 4], [(0, 1), (1, 0)]),
        _ring(6, extra=[(0, 3), (3, 0), (0, 3), (2, 2)]),
        _mol([5], []),
        _ring(5),
        _mol([2, 2], [(1, 0), (0, 1), (1, 0), (0, 1), (1, 0)]),
    ])
    # molecules beyond the thread-local limits (64 atoms / 160 edges) and beyond one tile (128 rows)
    big = _store([_ring(70), _ring(4), _ring(20, extra=[(i, (i + 7) % 20) for i in range(20)] * 8), _ring(3), _ring(100)])
    wide = _store([_ring(5), _ring(130), _ring(7)])
    # many isolated atoms in one molecule: more tiles than molecules
    dust = _store([_mol([2] * 300, [(0, 1), (1, 0)]), _ring(4, extra=[(0, 2)] * 300)])
    fixed = make_molecules(300, seed=3, atom_vocab=AV, bond_vocab=BV)
    variable = make_molecules(257, seed=4, atom_vocab=AV, bond_vocab=BV, variable=True)
    return dict(small=small, odd=odd, big=big, wide=wide, dust=dust, fixed=fixed, variable=variable)


CASES = _cases()


def _orders(store, seed=0):
    n = len(store)
    rng = np.random.default_rng(seed)
    return [np.arange(n), rng.permutation(n), rng.integers(0, n, size=max(1, n // 3))]


def _emulate(raw, force_scratch=0):
    total = int(raw.header[_lib.H_TOTAL_INTS])
    n_scr = _lib.check(int(_lib.lib().gcb_pack_device_scratch_ints(raw.header.ctypes.data)))
    packed = torch.full((total,), -7, dtype=torch.int32)   # poison: every int must be written
    scratch = torch.full((n_scr,), -9, dtype=torch.int32)
    rc = _lib.lib().gcb_debug_pack_emulate_host(raw.ints.data_ptr(), raw.header.ctypes.data, packed.data_ptr(), total,
                                                scratch.data_ptr(), n_scr, force_scratch)
    _lib.check(rc, "gcb_debug_pack_emulate_host")
    return packed


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("L", [2, 0])
def test_raw_collate_and_emulated_device_pack_equal_host_pack(name, L):
    store = CASES[name]
    for idx in _orders(store):
        ref = store.collate_pack(idx, L, AV, BV, pin=False)
        raw = store.collate_raw(idx, L, AV, BV, pin=False)
        assert np.array_equal(raw.header, ref.header), "layout / tile count"
        assert int(raw.raw_header[7]) == ref.n_tiles
        assert raw.nbytes < ref.nbytes
        if ref.y is not None:
            assert torch.equal(raw.y, ref.y)
        for force in (0, 1, 2):
            got = _emulate(raw, force)
            assert torch.equal(got, ref.ints), f"{name}: first difference at int {int((got != ref.ints).nonzero()[0])}"


def test_raw_collate_tiles_of_special_cases():
    assert CASES["wide"].collate_raw(np.arange(3), 2, AV, BV, pin=False).header[_lib.H_NTILES] == 0   # a 130-atom ring
    assert CASES["dust"].collate_raw(np.arange(2), 2, AV, BV, pin=False).header[_lib.H_NTILES] > 2    # more tiles than molecules
    empty = CASES["small"].collate_raw(np.arange(0), 2, AV, BV, pin=False)
    ref = CASES["small"].collate_pack(np.arange(0), 2, AV, BV, pin=False)
    assert np.array_equal(empty.header, ref.header) and torch.equal(_emulate(empty), ref.ints)


def test_raw_collate_errors_follow_the_host_packer():
    bad_tok = _store([_mol([2, AV], [(0, 1), (1, 0)])])
    with pytest.raises(IndexError):
        bad_tok.collate_raw([0], 2, AV, BV, pin=False)
    with pytest.raises(IndexError):
        bad_tok.collate_pack([0], 2, AV, BV, pin=False)
    few_edges = _store([_mol([2, 3, 4], [(2, 1), (1, 2)])])   # atom id 2 >= E = 2: the bond matrix has no such row
    with pytest.raises(IndexError):
        few_edges.collate_raw([0], 2, AV, BV, pin=False)
    few_edges.collate_raw([0], 0, AV, BV, pin=False)
    crossing = _store([_mol([2, 3], [(0, 2), (2, 0), (0, 1), (1, 0)]), _ring(4)])  # edge into the next molecule's atoms
    crossing.collate_pack([0, 1], 2, AV, BV, pin=False)  # the host packer's general path takes it
    with pytest.raises(NotImplementedError):
        crossing.collate_raw([0, 1], 2, AV, BV, pin=False)
    with pytest.raises(RuntimeError, match="runs on the GPU"):
        CASES["small"].collate_raw([0], 2, AV, BV, pin=False).pack()


# ---- GPU ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_device_pack_equals_host_pack(name):
    store = CASES[name]
    for L in (2, 0):
        for idx in _orders(store):
            ref = store.collate_pack(idx, L, AV, BV)
            raw = store.collate_raw(idx, L, AV, BV).to("cuda")
            total = int(raw.header[_lib.H_TOTAL_INTS])
            out = torch.full((total + 8,), -7, dtype=torch.int32, device="cuda")
            pb = raw.pack(out=out)
            torch.cuda.synchronize()
            assert np.array_equal(pb.header, ref.header)
            got = pb.ints.cpu()
            assert torch.equal(got, ref.ints), f"{name}: first difference at int {int((got != ref.ints).nonzero()[0])}"
            assert bool((out[total:] == -7).all()), "wrote past the packed batch"


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_compact_upload_rebuilds_the_ell_views_on_the_device(name):
    """`PackedBatch.to('cuda')` / `upload_into` move only header[H_COMPACT_INTS] ints over PCIe; gcb_pack_expand_ell derives
    the four ELL-4 views from the CSRs on the device.  The device buffer must equal the host-packed one bit for bit
    (plain and each-molecule-alone layouts, L = 0 included), and nothing may be written past it."""
    store = CASES[name]
    for L, each in ((2, False), (0, False), (2, True)):
        for idx in _orders(store):
            try:
                ref = store.collate_pack(idx, L, AV, BV, each=each) if each else store.collate_pack(idx, L, AV, BV)
            except (NotImplementedError, IndexError):
                continue  # each=True needs at least as many edges as atoms
            total, compact = ref.ints.numel(), ref.compact_ints
            assert 0 < compact <= total and compact == int(ref.header[_lib.H_IN_SRC4])
            assert total - compact == 4 * ((4 * ref.N + 3) // 4 * 4)
            dst = torch.full((total + 8,), -7, dtype=torch.int32, device="cuda")
            ref.upload_into(dst)
            torch.cuda.synchronize()
            assert torch.equal(dst[:total].cpu(), ref.ints), f"{name}: first difference at int {int((dst[:total].cpu() != ref.ints).nonzero()[0])}"
            assert bool((dst[total:] == -7).all()), "wrote past the packed batch"
            assert torch.equal(ref.to("cuda").ints.cpu(), ref.ints)


@pytest.mark.gpu
def test_device_pack_bench_shape_and_train_step_on_top():
    """The bench shape (8192 fixed-shape molecules) packed on the device: identical buffer, identical step."""
    from graphchem_b200.nn.gcn import MoleculeGCN
    store = make_molecules(8192, seed=11)
    idx = np.arange(8192)
    ref = store.collate_pack(idx, 4, 64, 32)
    pb = store.collate_raw(idx, 4, 64, 32).to("cuda").pack()
    assert torch.equal(pb.ints.cpu(), ref.ints)
    torch.manual_seed(0)
    model = MoleculeGCN(64, 32, 1, embedding_dim=128, n_messages=4, compute_dtype=torch.bfloat16).cuda()
    outs = []
    for batch in (ref.to("cuda"), pb):
        model.zero_grad()
        pred = model(batch)[0]
        torch.nn.functional.mse_loss(pred, batch.y).backward()
        outs.append((pred.detach().clone(), model.flat_parameters().grad if model.flat_parameters().grad is not None
                     else torch.cat([p.grad.reshape(-1) for p in model.parameters()])))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.gpu
@pytest.mark.parametrize("slots", [0, 3])
def test_loader_with_device_pack_delivers_the_host_packed_batches(slots):
    """PackedLoader(device_pack=True): raw collate on host threads, pack kernels on the copy stream; the batches
    equal a direct host collate_pack, across epochs (recycled pinned buffers) and through the device slot ring."""
    from graphchem_b200.data import PackedLoader
    store = make_molecules(1000, seed=5, variable=(slots == 0))
    loader = PackedLoader(store, 128, 4, 64, 32, device="cuda", workers=4, prefetch=2, device_slots=slots, device_pack=True,
                          drop_last=bool(slots))
    for epoch in range(3):
        n = 0
        for k, b in enumerate(loader):
            ref = store.collate_pack(np.arange(128 * k, min(1000, 128 * k + 128)), 4, 64, 32, pin=False)
            assert b.ints.is_cuda and torch.equal(b.ints.cpu(), ref.ints) and torch.equal(b.y.cpu(), ref.y)
            assert (b.header == ref.header).all()
            n += 1
        assert n == (7 if slots else 8)


def test_loader_device_pack_argument_checks():
    from graphchem_b200.data import PackedLoader
    with pytest.raises(ValueError):
        PackedLoader(CASES["small"], 8, 2, AV, BV, device="cpu", device_pack=True)
    with pytest.raises(ValueError):
        PackedLoader(CASES["small"], 8, 2, AV, BV, device="cuda", device_pack=True, each=True)


@pytest.mark.gpu
def test_thread_per_molecule_kernels_in_a_subprocess():
    """The default device packer is the warp-per-molecule kernel (every other GPU test here runs it); the
    thread-per-molecule kernels it replaced stay selectable (GCB_PACK_WARP=0, read once per process) and must keep
    producing the same buffers."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, torch\n"
        "from tests.test_pack_device import CASES, _orders, AV, BV\n"
        "for name, store in CASES.items():\n"
        "    for idx in _orders(store):\n"
        "        ref = store.collate_pack(idx, 2, AV, BV)\n"
        "        got = store.collate_raw(idx, 2, AV, BV).to('cuda').pack().ints.cpu()\n"
        "        assert torch.equal(got, ref.ints), name\n"
        "print('thread pack ok')\n")
    env = dict(os.environ, GCB_PACK_WARP="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "thread pack ok" in res.stdout, res.stdout + res.stderr
