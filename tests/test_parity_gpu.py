"""-m gpu: the CUDA path (through the C-ABI) against the CPU oracle on identical inputs and weights.

Tolerances (norm-wise, max|a-b| / max|oracle|), as BASELINE.json's north_star states them:
  fp32 forward 1e-5, fp32 gradients 1e-4; bf16 forward 1e-2 per message step (grows with depth,
  bound stated per test), bf16 gradients 5e-2.  Integer structures are compared bit-exactly in
  tests/test_pack.py.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import make_pair, nerr, oracle_fp64, random_batch, random_graph_batch

pytestmark = pytest.mark.gpu

FWD_F32, GRAD_F32 = 1e-5, 1e-4
FWD_BF16, GRAD_BF16 = 2e-2, 6e-2

ACTS = [F.softplus, F.sigmoid, F.relu, F.tanh, F.leaky_relu]


def _to_dev(data, device="cuda"):
    from graphchem_b200.data import Data
    d = Data(x=data.x.to(device), edge_index=data.edge_index.to(device), edge_attr=data.edge_attr.to(device),
             y=None if data.y is None else data.y.to(device))
    if getattr(data, "batch", None) is not None:
        d.batch = data.batch.to(device)
        d.num_graphs = int(data.batch.max()) + 1 if data.batch.numel() else 0
    return d


def _check_forward(oracle, model, data, tol):
    """CUDA forward against the float64 evaluation of the oracle (the CUDA path's own error, tests/util.py
    oracle_fp64); returns (errors vs fp64, rounding noise of the fp32 oracle vs fp64) per output."""
    oracle.eval(); model.eval()
    o64 = oracle_fp64(oracle)
    with torch.no_grad():
        r64 = o64(data)
        r32 = oracle(data)
        rm = model(_to_dev(data))
    errs = [nerr(a.float(), b) for a, b in zip(rm, r64)]
    noise = [nerr(a, b) for a, b in zip(r32, r64)]
    assert all(e <= tol for e in errs), (f"forward errors vs fp64 oracle (out, out_atom, out_bond) = {errs} > {tol}; "
                                         f"fp32 oracle's own rounding noise = {noise}")
    return errs


def _check_backward(oracle, model, data, tol, seed=0):
    """All parameter gradients of a loss that touches the three outputs, CUDA autograd.Function against torch
    autograd through the float64 oracle."""
    oracle.train(); model.train()  # p_dropout == 0 in these tests: train() only enables grads bookkeeping
    o64 = oracle_fp64(oracle)
    g = torch.Generator().manual_seed(seed)
    ro = o64(data)
    wa = torch.randn(ro[1].shape, generator=g) * 0.1
    wb = torch.randn(ro[2].shape, generator=g) * 0.1
    y = torch.randn(ro[0].shape, generator=g)
    lo = (F.mse_loss(ro[0], y.double()) + (ro[1] * wa.double()).sum() / max(ro[1].numel(), 1)
          + (ro[2] * wb.double()).sum() / max(ro[2].numel(), 1))
    o64.zero_grad(); lo.backward()
    rm = model(_to_dev(data))
    lm = (F.mse_loss(rm[0], y.cuda()) + (rm[1].float() * wa.cuda()).sum() / max(ro[1].numel(), 1)
          + (rm[2].float() * wb.cuda()).sum() / max(ro[2].numel(), 1))
    model.zero_grad(); lm.backward()
    errs = {}
    for (name, po), (_, pm) in zip(o64.named_parameters(), model.named_parameters()):
        assert pm.grad is not None, name
        if po.grad is None or float(po.grad.abs().max()) == 0.0:
            assert float(pm.grad.abs().max()) <= 1e-6, name
            continue
        errs[name] = nerr(pm.grad, po.grad)
    bad = {k: v for k, v in errs.items() if v > tol}
    assert not bad, f"gradient errors vs fp64 oracle above {tol}: {bad} (all: {errs})"
    return errs


@pytest.mark.parametrize("L,n_readout,D,R", [(2, 2, 128, 64), (3, 3, 128, 128), (1, 1, 64, 32), (0, 2, 64, 64),
                                             (6, 0, 64, 64), (4, 2, 256, 64)])
def test_forward_backward_fp32_molecules(L, n_readout, D, R):
    data, _ = random_batch(seed=L * 7 + n_readout, n_graphs=16)
    oracle, model = make_pair(12, 7, 1 if n_readout else None, embedding_dim=D, n_messages=L, n_readout=n_readout,
                              readout_dim=R)
    _check_forward(oracle, model, data, FWD_F32)
    _check_backward(oracle, model, data, GRAD_F32)


def test_per_layer_activations_fp32():
    """Weights are shared across message steps, so a model with n_messages=l returns the layer-l
    activations: compare every layer, not only the last."""
    data, _ = random_batch(seed=5, n_graphs=24)
    for l in range(0, 5):
        oracle, model = make_pair(12, 7, 1, seed=3, embedding_dim=128, n_messages=l)
        _check_forward(oracle, model, data, FWD_F32)


@pytest.mark.parametrize("act", ACTS)
def test_activations_fp32(act):
    data, _ = random_batch(seed=11, n_graphs=8)
    oracle, model = make_pair(12, 7, 2, embedding_dim=64, n_messages=3, n_readout=2, readout_dim=48, act_fn=act)
    _check_forward(oracle, model, data, FWD_F32)
    _check_backward(oracle, model, data, GRAD_F32 * (3 if act is F.relu or act is F.leaky_relu else 1))


def test_aggr_mean_fp32():
    data, _ = random_batch(seed=13, n_graphs=8)
    oracle, model = make_pair(12, 7, 1, embedding_dim=64, n_messages=3, aggr="mean")
    _check_forward(oracle, model, data, FWD_F32)
    _check_backward(oracle, model, data, GRAD_F32)


@pytest.mark.parametrize("dtype,ftol,gtol", [(torch.float32, FWD_F32, GRAD_F32), (torch.bfloat16, FWD_BF16, GRAD_BF16)])
def test_global_mean_pool(dtype, ftol, gtol):
    """pool="mean" (BASELINE north_star item 4; PyG global_mean_pool = segment sum / max(count, 1)), batched and bare"""
    data, graphs = random_batch(seed=23, n_graphs=64 if dtype == torch.bfloat16 else 8, max_atoms=30)
    oracle, model = make_pair(12, 7, 2, dtype=dtype, embedding_dim=128 if dtype == torch.bfloat16 else 64, n_messages=2, pool="mean")
    _check_forward(oracle, model, data, ftol)
    _check_backward(oracle, model, data, gtol)
    if dtype == torch.float32:
        _check_forward(oracle, model, graphs[0], ftol)


@pytest.mark.parametrize("act", [F.softplus, F.relu])
def test_aggr_max_fp32(act):
    """GeneralConv(aggr="max"), gcn.py:84-86: transform-then-aggregate per edge, amax with even tie split in
    backward (relu makes exact ties -- several in-edges at 0 -- common)."""
    data, _ = random_batch(seed=17, n_graphs=8)
    oracle, model = make_pair(12, 7, 1, embedding_dim=64, n_messages=3, aggr="max", act_fn=act)
    _check_forward(oracle, model, data, FWD_F32)
    _check_backward(oracle, model, data, GRAD_F32 * (3 if act is F.relu else 1))


@pytest.mark.parametrize("N,E,G", [(3, 5, 1), (40, 40, 1), (50, 200, 4), (64, 70, 3)])
def test_aggr_max_random_graphs_fp32(N, E, G):
    """the reference's own test inputs (tests/test_gcn.py:51-56): duplicate edges, isolated atoms (empty max = 0), E == N,
    bond rows >= N that stay act(0)"""
    data = random_graph_batch(seed=N + E + 1, N=N, E=E, av=10, bv=5, G=G)
    oracle, model = make_pair(10, 5, 2, embedding_dim=40, n_messages=2, n_readout=1, readout_dim=32, aggr="max")
    _check_forward(oracle, model, data, FWD_F32)
    _check_backward(oracle, model, data, GRAD_F32)


def test_aggr_max_without_edges_and_bare_molecules():
    """aggr="max" on a lone atom (no edges at all: every maximum is the empty one, 0) and on bare molecules (batch=None)"""
    from oracle.pyg_restated import Data as OData
    _, graphs = random_batch(seed=3, n_graphs=3, min_atoms=3)
    oracle, model = make_pair(12, 7, 1, embedding_dim=64, n_messages=2, aggr="max")
    single = OData(x=torch.tensor([3], dtype=torch.int32), edge_index=torch.zeros(2, 0, dtype=torch.long),
                   edge_attr=torch.zeros(0, dtype=torch.int32), y=torch.zeros(1, 1))
    _check_forward(oracle, model, single, FWD_F32)
    _check_backward(oracle, model, single, GRAD_F32)
    for gr in graphs:
        _check_forward(oracle, model, gr, FWD_F32)
        _check_backward(oracle, model, gr, GRAD_F32)


def test_aggr_max_bf16():
    data, _ = random_batch(seed=19, n_graphs=64, max_atoms=30)
    oracle, model = make_pair(12, 7, 1, dtype=torch.bfloat16, embedding_dim=128, n_messages=2, aggr="max")
    _check_forward(oracle, model, data, FWD_BF16)
    _check_backward(oracle, model, data, GRAD_BF16)


@pytest.mark.parametrize("N,E,G", [(3, 5, 1), (40, 40, 1), (50, 200, 4), (64, 70, 3)])
def test_random_graphs_duplicates_isolated_fp32(N, E, G):
    """The reference's own test inputs (tests/test_gcn.py:51-56): random edge_index with duplicate
    edges, isolated nodes and E == N."""
    data = random_graph_batch(seed=N + E, N=N, E=E, av=10, bv=5, G=G)
    oracle, model = make_pair(10, 5, 2, embedding_dim=64, n_messages=3, n_readout=3, readout_dim=128)
    _check_forward(oracle, model, data, FWD_F32)
    _check_backward(oracle, model, data, GRAD_F32)


def test_batch_none_single_molecule_and_single_atom():
    """predict_cn.ipynb:333-334 evaluates bare graphs (batch=None); 'C' has one atom and no bonds."""
    from oracle.pyg_restated import Data as OData
    _, graphs = random_batch(seed=2, n_graphs=3, min_atoms=3)
    oracle, model = make_pair(12, 7, 1, embedding_dim=64, n_messages=2)
    for gr in graphs:
        _check_forward(oracle, model, gr, FWD_F32)
        _check_backward(oracle, model, gr, GRAD_F32)
    single = OData(x=torch.tensor([3], dtype=torch.int32), edge_index=torch.zeros(2, 0, dtype=torch.long),
                   edge_attr=torch.zeros(0, dtype=torch.int32), y=torch.zeros(1, 1))
    errs = _check_forward(oracle, model, single, FWD_F32)
    assert errs[2] == 0.0


def test_index_error_when_edge_index_exceeds_bond_rows():
    """['C','CC'] batched: N=3, E=2 -> the reference raises IndexError at gcn.py:163."""
    from oracle.pyg_restated import Data as OData, collate
    c = OData(x=torch.tensor([2], dtype=torch.int32), edge_index=torch.zeros(2, 0, dtype=torch.long),
              edge_attr=torch.zeros(0, dtype=torch.int32), y=torch.zeros(1, 1))
    cc = OData(x=torch.tensor([2, 2], dtype=torch.int32), edge_index=torch.tensor([[0, 1], [1, 0]]),
               edge_attr=torch.tensor([2, 2], dtype=torch.int32), y=torch.zeros(1, 1))
    data = collate([c, cc])
    oracle, model = make_pair(12, 7, 1, embedding_dim=64, n_messages=2)
    with pytest.raises(IndexError):
        oracle(data)
    with pytest.raises(IndexError):
        model(_to_dev(data))


def test_tie_splitting_in_amax_backward():
    """Two sources with identical bond rows tie in the max; torch splits the gradient evenly."""
    from oracle.pyg_restated import Data as OData
    # star: atoms 1..4 -> 0 and back; all bond tokens equal -> identical rows of the bond matrix
    src = [1, 2, 3, 4, 0, 0, 0, 0]
    dst = [0, 0, 0, 0, 1, 2, 3, 4]
    data = OData(x=torch.tensor([2, 3, 3, 4, 5]), edge_index=torch.tensor([src, dst]),
                 edge_attr=torch.full((8,), 3), y=torch.zeros(1, 1))
    oracle, model = make_pair(12, 7, 1, embedding_dim=64, n_messages=3)
    _check_forward(oracle, model, data, FWD_F32)
    _check_backward(oracle, model, data, GRAD_F32)


@pytest.mark.parametrize("L,D", [(1, 128), (4, 128), (6, 256)])
def test_forward_backward_bf16(L, D):
    data, _ = random_batch(seed=21 + L, n_graphs=64, max_atoms=30)
    oracle, model = make_pair(12, 7, 1, dtype=torch.bfloat16, embedding_dim=D, n_messages=L)
    _check_forward(oracle, model, data, FWD_BF16)
    _check_backward(oracle, model, data, GRAD_BF16)


def test_bf16_matches_fp32_cuda_path_loosely_on_synthetic_batch():
    """Config-3-shaped data (23 atoms / 50 edges per molecule), 256 molecules."""
    from graphchem_b200.synthetic import make_molecules
    from oracle.pyg_restated import Data as OData, collate
    store = make_molecules(256, seed=7)
    graphs = []
    for i in range(256):
        a0, a1 = store.atom_off[i], store.atom_off[i + 1]
        e0, e1 = store.edge_off[i], store.edge_off[i + 1]
        graphs.append(OData(x=store.atoms[a0:a1], edge_attr=store.bonds[e0:e1],
                            edge_index=store.conn[2 * e0:2 * e1].view(2, -1), y=store.targets[i:i + 1]))
    data = collate(graphs)
    oracle, model = make_pair(64, 32, 1, dtype=torch.bfloat16, embedding_dim=128, n_messages=4)
    _check_forward(oracle, model, data, FWD_BF16)
    oracle32, model32 = make_pair(64, 32, 1, embedding_dim=128, n_messages=4)
    _check_forward(oracle32, model32, data, FWD_F32)
    _check_backward(oracle32, model32, data, GRAD_F32)


def test_dropout_semantics():
    """reference tests/test_gcn.py:28-62 runs train mode with p_dropout=0.2 (shapes only)."""
    data = random_graph_batch(seed=1, N=3, E=5, av=10, bv=5)
    oracle, model = make_pair(10, 5, 2, embedding_dim=64, n_messages=3, n_readout=3, readout_dim=128, p_dropout=0.2)
    model.train()
    torch.manual_seed(0)
    out, out_atom, out_bond = model(_to_dev(data))
    assert out.shape == (1, 2) and out_atom.shape == (3, 64) and out_bond.shape == (5, 64)
    torch.manual_seed(0)
    out2, _, _ = model(_to_dev(data))
    assert torch.equal(out, out2), "same torch seed must give the same dropout mask"
    out.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    # statistics on a larger graph: about 20% zeros, survivors scaled by 1/(1-p)
    big, _ = random_batch(seed=3, n_graphs=64)
    oracle, model = make_pair(12, 7, 1, embedding_dim=128, n_messages=1, p_dropout=0.2)
    model.train()
    _, atom_d, bond_d = model(_to_dev(big))
    model.eval()
    with torch.no_grad():
        _, atom_e, bond_e = model(_to_dev(big))
    zero_frac = float((atom_d == 0).float().mean())
    assert 0.17 < zero_frac < 0.23, zero_frac
    keep = bond_d != 0
    assert torch.allclose(bond_d[keep], bond_e[keep] / 0.8, rtol=1e-5, atol=1e-6)
    # eval mode == oracle eval mode
    _check_forward(oracle, model, big, FWD_F32)


@pytest.mark.parametrize("aggr", ["add", "max"])
def test_dropout_gradient_matches_masked_oracle(aggr):
    """Recover the kernel's keep-mask from a train-mode forward (n_messages=1, n_readout=0), replay
    it in the oracle and compare gradients: checks the backward's regenerated masks."""
    data, _ = random_batch(seed=4, n_graphs=8)
    oracle, model = make_pair(12, 7, None, embedding_dim=64, n_messages=1, n_readout=0, p_dropout=0.25, aggr=aggr)
    model.train()
    torch.manual_seed(123)
    out, out_atom, out_bond = model(_to_dev(data))
    mask_a = (out_atom != 0).float().cpu() / 0.75
    mask_b = (out_bond != 0).float().cpu() / 0.75
    out.sum().backward()
    # oracle with explicit masks
    o = oracle
    A0 = o.act_fn(o.emb_atom(data.x)); B0 = o.act_fn(o.emb_bond(data.edge_attr))
    B1 = o.act_fn(o.bond_conv(B0, data.edge_index)) * mask_b
    A1 = o.act_fn(o.atom_conv(A0, data.edge_index, B1)) * mask_a
    o.zero_grad(); A1.sum().backward()
    assert nerr(out_atom, A1) <= FWD_F32
    for (name, po), (_, pm) in zip(o.named_parameters(), model.named_parameters()):
        if po.grad is None:
            continue
        assert nerr(pm.grad, po.grad) <= GRAD_F32, name


def test_mse_and_adam_match_torch():
    import ctypes as C
    from graphchem_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator().manual_seed(0)
    pred, y = torch.randn(4096, 1, generator=g).cuda(), torch.randn(4096, 1, generator=g).cuda()
    loss, dpred = torch.zeros(1, device="cuda"), torch.empty_like(pred)
    s = torch.cuda.current_stream().cuda_stream
    assert lib.gcb_mse_loss(pred.data_ptr(), y.data_ptr(), pred.numel(), loss.data_ptr(), dpred.data_ptr(), 1.0, s) == 0
    pr = pred.clone().requires_grad_(True)
    lr = F.mse_loss(pr, y); lr.backward()
    assert nerr(loss, lr.detach().reshape(1)) <= 1e-6 and nerr(dpred, pr.grad) <= 1e-6
    # Adam: 5 steps against torch.optim.Adam with a changing lr (predict_cn.ipynb:241-242)
    n = 100003
    p0 = torch.randn(n, generator=g).cuda()
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p_ref], lr=1e-3)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    hyper, step = torch.tensor([1e-3, 1.0], device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
    for it in range(5):
        grad = torch.randn(n, generator=g).cuda()
        lr_now = 1e-3 - it * 1e-4
        for grp in opt.param_groups:
            grp["lr"] = lr_now
        p_ref.grad = grad.clone(); opt.step()
        hyper[0] = lr_now
        assert lib.gcb_adam_step(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), n, hyper.data_ptr(),
                                 step.data_ptr(), 0.9, 0.999, 1e-8, s) == 0
        assert nerr(p, p_ref.detach()) <= 1e-6, it
    assert int(step.item()) == 5


def test_cpu_inputs_are_device_transparent_and_notebook_loop_runs():
    """The notebooks never mention a device (examples/predict_cn.ipynb:212-272): CPU batches in,
    CPU tensors out, torch.optim.Adam over model.parameters()."""
    from graphchem_b200.data import Batch, DataLoader, MoleculeDataset, MoleculeGraph
    from graphchem_b200.nn import MoleculeGCN
    _, graphs = random_batch(seed=9, n_graphs=40)
    ds = MoleculeDataset([MoleculeGraph(g.x, g.edge_attr, g.edge_index, g.y) for g in graphs])
    loader = DataLoader(ds, batch_size=16, shuffle=True)
    torch.manual_seed(0)
    model = MoleculeGCN(12, 7, 1, embedding_dim=64, n_messages=2).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    model.train()
    losses = []
    for epoch in range(6):
        tot = 0.0
        for batch in loader:
            opt.zero_grad()
            pred, _, _ = model(batch)
            assert pred.device.type == "cpu"
            loss = F.mse_loss(pred, batch.y)
            loss.backward()
            opt.step()
            tot += loss.detach().item()
        losses.append(tot)
    assert losses[-1] < losses[0]
    model.eval()
    pred, a, b = model(ds[0])
    assert pred.shape == (1, 1) and a.shape[0] == ds[0].x.shape[0]


def test_high_degree_random_graph_bf16_overflows_the_ell4_view():
    """In-/out-degrees well above 4 exercise the CSR tail loop behind the ELL-4 fast path."""
    data = random_graph_batch(seed=77, N=300, E=2400, av=10, bv=5, G=6)
    oracle, model = make_pair(10, 5, 1, dtype=torch.bfloat16, embedding_dim=128, n_messages=2)
    _check_forward(oracle, model, data, FWD_BF16)
    # ~8 candidates per max: bf16 rounding flips near-ties that fp32 resolves, which moves the bond
    # weight gradient by up to ~8% on this adversarial graph (every other parameter stays within 3%)
    _check_backward(oracle, model, data, 0.15)


@pytest.mark.parametrize("n_readout,R,out_dim", [(2, 64, 1), (3, 128, 2), (1, 36, 3)])
def test_large_batch_readout_fp32(n_readout, R, out_dim):
    """More than 1024 graphs: the register-tiled readout kernels (readout.cuh) against the oracle, and
    against the GEMM-per-layer path they replace."""
    import os
    data, _ = random_batch(seed=21, n_graphs=1300, max_atoms=6)
    oracle, model = make_pair(12, 7, out_dim, embedding_dim=64, n_messages=1, n_readout=n_readout, readout_dim=R)
    _check_forward(oracle, model, data, FWD_F32)
    _check_backward(oracle, model, data, GRAD_F32)
    dev = _to_dev(data)
    model.train()

    def grads():
        out = model(dev)[0]
        model.zero_grad()
        (out ** 2).mean().backward()
        return out.detach().clone(), [p.grad.clone() for p in model.parameters()]

    o1, g1 = grads()
    os.environ["GCB_READOUT_GB"] = "32"   # 32 graphs per CTA instead of 64
    try:
        o3, g3 = grads()
    finally:
        os.environ.pop("GCB_READOUT_GB", None)
    assert nerr(o3, o1) <= 1e-6 and all(nerr(a, b) <= 1e-5 for a, b in zip(g3, g1))
    os.environ["GCB_NO_READOUT2"] = "1"
    try:
        o2, g2 = grads()
    finally:
        os.environ.pop("GCB_NO_READOUT2", None)
    assert nerr(o1, o2) <= 1e-6
    for a, b in zip(g1, g2):
        assert nerr(a, b) <= 1e-5


@pytest.mark.parametrize("name", ["gcn_small.pt", "gcn_small_max_meanpool.pt"])
def test_cuda_path_matches_committed_golden_vectors(name):
    """The committed oracle outputs (tests/golden/gcn_small*.pt): forward, gradients and three Adam steps of the
    fused TrainStep, fp32, through the C-ABI; the second file is GeneralConv(aggr="max") with mean pooling."""
    import os
    from graphchem_b200.data import Data
    from graphchem_b200.nn import MoleculeGCN
    from graphchem_b200.train import TrainStep
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name), weights_only=False)
    d = gold["dims"]
    model = MoleculeGCN(d["av"], d["bv"], d["out"], embedding_dim=d["embedding_dim"], n_messages=d["n_messages"],
                        n_readout=d["n_readout"], readout_dim=d["readout_dim"], **gold.get("extra", {}))
    model.load_state_dict(gold["state"])
    model = model.cuda().train()
    dev = Data(x=gold["x"].cuda(), edge_index=gold["edge_index"].cuda(), edge_attr=gold["edge_attr"].cuda(),
               batch=gold["batch"].cuda(), y=gold["y"].cuda(), num_graphs=int(gold["batch"].max()) + 1)
    out, out_atom, out_bond = model(dev)
    assert nerr(out, gold["out"]) <= FWD_F32 and nerr(out_atom, gold["out_atom"]) <= FWD_F32
    assert nerr(out_bond, gold["out_bond"]) <= FWD_F32
    loss = F.mse_loss(out, dev.y)
    assert abs(float(loss) - gold["loss"]) <= 1e-5 * max(1.0, abs(gold["loss"]))
    model.zero_grad(); loss.backward()
    for name, p in model.named_parameters():
        assert nerr(p.grad, gold["grads"][name]) <= GRAD_F32, name
    step = TrainStep(model, lr=1e-3, use_cuda_graph=False)
    pb = model.pack(dev).to("cuda")
    losses = [float(step(pb).item()) for _ in range(3)]
    for a, b in zip(losses, gold["adam_losses"]):
        assert abs(a - b) <= 1e-4 * max(1.0, abs(b))
    from tests.util import assert_adam_states_close
    for name, v in model.state_dict().items():
        assert_adam_states_close(v, gold["state_after_3_adam_steps"][name], 1e-3, 3, name)
