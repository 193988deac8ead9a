"""-m gpu: the tile-staged row kernels (csrc/tile_rows.cuh, the default when the batch has closed row
tiles) and the opt-in fused tcgen05 GEMM + message kernels (csrc/fused_rows.cuh, GCB_FUSED=1) against the
plain row kernels (kernels_v2.cuh, GCB_NO_TILE=1).

The tile kernels, the fused atom backward and the fused bond forward in its GCB_FUSED_EXACT=1 form
round at the same points and sum in the same order as the plain kernels, so they must agree BIT
FOR BIT: predictions, per-atom / per-bond activations and every parameter gradient.  The default
fused path keeps P in fp32 and pre-loads the atom accumulator with residual + bias (different
rounding points, closer to the fp32 reference): it must agree within a few bf16 ulps.  The oracle
parity of the default path is covered by tests/test_parity_gpu.py."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import make_pair, nerr, random_batch

pytestmark = pytest.mark.gpu


def _dev(data):
    from graphchem_b200.data import Data
    return Data(x=data.x.cuda(), edge_index=data.edge_index.cuda(), edge_attr=data.edge_attr.cuda(), y=data.y.cuda(),
                batch=data.batch.cuda(), num_graphs=int(data.batch.max()) + 1)


ENV = {"tile": {"GCB_NO_FUSED": "1", "GCB_NO_TOK1": "1"},
       "tile_split": {"GCB_NO_BOND_MERGE": "1", "GCB_NO_TOK1": "1"}, "tok1": {},
       "fused_exact": {"GCB_FUSED": "1", "GCB_FUSED_EXACT": "1", "GCB_NO_FUSED_ATOM": "1"}, "fused": {"GCB_FUSED": "1"}, "v2": {"GCB_NO_TILE": "1"}}


def _run(model, dev, train, mode):
    os.environ.update(ENV[mode])
    try:
        model.train(train)
        if not train:
            with torch.no_grad():
                return [t.clone() for t in model(dev)], None
        out = model(dev)
        g = torch.Generator(device="cuda").manual_seed(3)
        loss = (F.mse_loss(out[0], dev.y) + (out[1].float() * torch.randn(out[1].shape, device="cuda", generator=g)).mean()
                + (out[2].float() * torch.randn(out[2].shape, device="cuda", generator=g)).mean())
        model.zero_grad()
        loss.backward()
        torch.cuda.synchronize()
        return [t.detach().clone() for t in out], [p.grad.clone() for p in model.parameters()]
    finally:
        for k in ENV[mode]:
            os.environ.pop(k, None)


@pytest.mark.parametrize("n_graphs,max_atoms,L,D", [(1, 12, 1, 128), (7, 12, 2, 128), (64, 30, 3, 128), (700, 40, 4, 128),
                                                      (5, 12, 1, 256), (300, 30, 2, 256)])
@pytest.mark.parametrize("act", [F.softplus, F.relu])
@pytest.mark.parametrize("mode", ["tile", "tile_split", "fused_exact"])
def test_fused_equals_unfused_bitwise(n_graphs, max_atoms, L, D, act, mode):
    if mode == "fused_exact" and D != 128:
        pytest.skip("the fused tcgen05 kernels are D = 128 only")
    data, _ = random_batch(seed=n_graphs, n_graphs=n_graphs, max_atoms=max_atoms)
    _, model = make_pair(12, 7, 1, dtype=torch.bfloat16, embedding_dim=D, n_messages=L, act_fn=act)
    dev = _dev(data)
    from graphchem_b200.data import pack_batch
    assert pack_batch(data, L, 12, 7, pin=False).n_tiles > 0
    for train in (False, True):
        out_f, grad_f = _run(model, dev, train, mode)
        out_u, grad_u = _run(model, dev, train, "v2")
        for k, (a, b) in enumerate(zip(out_f, out_u)):
            assert torch.equal(a, b), f"output {k} differs (train={train}): max diff {float((a.float() - b.float()).abs().max())}"
        if train:
            for (name, _), a, b in zip(model.named_parameters(), grad_f, grad_u):
                assert torch.equal(a, b), f"grad {name} differs: {nerr(a, b)}"


@pytest.mark.parametrize("n_graphs,max_atoms,L,D", [(1, 12, 1, 128), (7, 12, 2, 128), (64, 30, 3, 128), (700, 40, 4, 128),
                                                      (3000, 30, 1, 128), (300, 30, 2, 256)])
@pytest.mark.parametrize("act", [F.softplus, F.relu])
def test_first_step_from_the_token_table(n_graphs, max_atoms, L, D, act):
    """Default path: message step 1 reads [P|Q] of every bond TOKEN (tile::bond_update_tok_kernel; the table comes out
    of the same tcgen05 GEMM, so its rows are bit-identical to the [rows, 2D] matrix they replace) and the backward
    pass sums [dP|dQ] per token before the small table products.  Forward: bit-identical to the tile kernels with the
    table switched off.  Gradients: identical except the three tensors step 1's bond transform feeds, which round
    differently (sum-then-transform) and must agree to a few bf16 ulps of the largest entry."""
    data, _ = random_batch(seed=40 + n_graphs, n_graphs=n_graphs, max_atoms=max_atoms)
    _, model = make_pair(12, 7, 1, dtype=torch.bfloat16, embedding_dim=D, n_messages=L, act_fn=act)
    dev = _dev(data)
    for train in (False, True):
        out_t, grad_t = _run(model, dev, train, "tok1")
        out_u, grad_u = _run(model, dev, train, "tile")
        for k, (a, b) in enumerate(zip(out_t, out_u)):
            assert torch.equal(a, b), f"output {k} differs (train={train}): max diff {float((a.float() - b.float()).abs().max())}"
        if train:
            for (name, _), a, b in zip(model.named_parameters(), grad_t, grad_u):
                if name in ("emb_bond.weight", "bond_conv.nn.0.weight", "bond_conv.nn.0.bias"):
                    assert nerr(a, b) <= 2e-2, (name, nerr(a, b))
                else:
                    assert torch.equal(a, b), f"grad {name} differs: {nerr(a, b)}"


@pytest.mark.parametrize("n_graphs,max_atoms,L", [(1, 12, 1), (7, 12, 2), (64, 30, 3), (700, 40, 4), (3000, 30, 4)])
@pytest.mark.parametrize("act", [F.softplus, F.tanh])  # smooth activations: relu's kink makes two roundings of the same value disagree on 0/1 derivatives
def test_default_fused_path_close_to_unfused(n_graphs, max_atoms, L, act):
    """Default fused kernels (fp32 P, accumulator pre-loaded with residual + deg * bias) against the plain row
    kernels: a different but equally valid set of bf16 rounding points (two bf16 paths that are each within the
    bf16 bound of the fp32 oracle, tests/test_parity_gpu.py, may be twice that apart)."""
    data, _ = random_batch(seed=100 + n_graphs, n_graphs=n_graphs, max_atoms=max_atoms)
    _, model = make_pair(12, 7, 1, dtype=torch.bfloat16, embedding_dim=128, n_messages=L, act_fn=act)
    dev = _dev(data)
    for train in (False, True):
        out_f, grad_f = _run(model, dev, train, "fused")
        out_u, grad_u = _run(model, dev, train, "v2")
        for k, (a, b) in enumerate(zip(out_f, out_u)):
            assert nerr(a.float(), b.float()) <= 4e-2 or float((a.float() - b.float()).abs().max()) <= 2e-2, (k, train, nerr(a.float(), b.float()))
        if train:
            for (name, _), a, b in zip(model.named_parameters(), grad_f, grad_u):
                if float(b.abs().max()) > 0:
                    assert nerr(a, b) <= 1.5e-1, (name, nerr(a, b))


def test_fused_kernels_are_launched_and_relu_ties_match():
    """The fused kernels must actually run (timing facility lists them), including the tie
    bookkeeping of the max aggregation (relu makes exact ties common)."""
    import ctypes as C
    from graphchem_b200 import _lib
    data, _ = random_batch(seed=11, n_graphs=200, max_atoms=25)
    _, model = make_pair(12, 7, 1, dtype=torch.bfloat16, embedding_dim=128, n_messages=2, act_fn=F.relu)
    dev = _dev(data)
    from graphchem_b200.train import TrainStep
    pb = model.pack(data).to("cuda")
    step = TrainStep(model, use_cuda_graph=False)  # forward and backward on this thread (the timing table is thread-local)
    lib = _lib.lib()

    def kernels_of_a_step(env):
        os.environ.update(env)
        lib.gcb_timing_enable(1)
        try:
            step(pb)
            torch.cuda.synchronize()
        finally:
            lib.gcb_timing_enable(0)
            for k in env:
                os.environ.pop(k, None)
        buf = C.create_string_buffer(1 << 16)
        lib.gcb_timing_report(buf, len(buf))
        return buf.value.decode()

    names = kernels_of_a_step({"GCB_FUSED": "1"})
    for k in ("fused_bond_fwd_kernel", "fused_atom_fwd_kernel", "fused_atom_bwd_kernel", "tile_bond_bwd_kernel"):
        assert k in names, (k, names)
    for k in ("tile_bond_update_kernel", "tile_atom_gather_kernel", "tile_atom_bwd_kernel"):
        assert k not in names, (k, names)
    names = kernels_of_a_step({})  # the default: tile kernels + stand-alone tcgen05 GEMMs, step 1 from the token table
    for k in ("tile_bond_update_tok_kernel", "tile_bond_update_kernel", "tile_atom_gather_kernel", "tile_atom_bwd_kernel",
              "tile_bond_bwd_kernel", "bond_table_bwd_kernel"):
        assert k in names, (k, names)
    for k in ('"bond_update_kernel"', '"atom_gather_kernel"', "atom_bwd_gather_kernel", "bond_bwd_scatter_kernel", "bond_bwd_tie_kernel"):
        assert k not in names, (k, names)
    names = kernels_of_a_step({"GCB_NO_BOND_MERGE": "1"})
    assert "tile_bond_scatter_kernel" in names and "bond_bwd_tie_kernel" in names, names


def test_wide_component_falls_back_to_unfused():
    """A connected component wider than a tile has no closed tiling: n_tiles == 0 and the unfused
    kernels run (results still match the oracle within the bf16 bound)."""
    from graphchem_b200.data import pack_batch
    from oracle.pyg_restated import Data as OData
    n = 300  # a 300-atom chain
    src = torch.arange(n - 1)
    ei = torch.stack([torch.cat([src, src + 1]), torch.cat([src + 1, src])])
    g = torch.Generator().manual_seed(0)
    d = OData(x=torch.randint(2, 12, (n,), generator=g).int(), edge_index=ei,
              edge_attr=torch.randint(2, 7, (ei.size(1),), generator=g).int(), y=torch.zeros(1, 1))
    d.batch = torch.zeros(n, dtype=torch.long)
    assert pack_batch(d, 2, 12, 7, pin=False).n_tiles == 0
    oracle, model = make_pair(12, 7, 1, dtype=torch.bfloat16, embedding_dim=128, n_messages=2)
    oracle.eval(); model.eval()
    with torch.no_grad():
        ro = oracle(d)
        rm = model(_dev(d))
    assert nerr(rm[0].float(), ro[0]) <= 2e-2 and nerr(rm[1].float(), ro[1]) <= 2e-2


@pytest.mark.parametrize("chain", [127, 128, 129])
def test_tile_boundary_molecule_sizes(chain):
    """A molecule that exactly fills a 128-row tile (and its neighbours 127 / 129: the latter has no closed
    tiling and runs untiled) next to small molecules: forward and gradients against the oracle."""
    from graphchem_b200.data import pack_batch
    from oracle.pyg_restated import Data as OData, collate as ocollate
    g = torch.Generator().manual_seed(chain)
    src = torch.arange(chain - 1)
    big = OData(x=torch.randint(2, 12, (chain,), generator=g).int(),
                edge_index=torch.stack([torch.cat([src, src + 1]), torch.cat([src + 1, src])]),
                edge_attr=torch.randint(2, 7, (2 * (chain - 1),), generator=g).int(), y=torch.randn(1, 1, generator=g))
    _, small = random_batch(seed=chain, n_graphs=20, max_atoms=9)
    data = ocollate(small[:10] + [big] + small[10:])
    pb = pack_batch(data, 2, 12, 7, pin=False)
    assert (pb.n_tiles > 0) == (chain <= 128)
    oracle, model = make_pair(12, 7, 1, dtype=torch.bfloat16, embedding_dim=128, n_messages=2)
    dev = _dev(data)
    oracle.eval(); model.eval()
    with torch.no_grad():
        ro, rm = oracle(data), model(dev)
    for a, b in zip(rm, ro):
        assert nerr(a.float(), b) <= 2e-2
    oracle.train(); model.train()
    lo = F.mse_loss(oracle(data)[0], data.y)
    oracle.zero_grad(); lo.backward()
    lm = F.mse_loss(model(dev)[0], dev.y)
    model.zero_grad(); lm.backward()
    for (name, po), (_, pm) in zip(oracle.named_parameters(), model.named_parameters()):
        if po.grad is not None and float(po.grad.abs().max()) > 0:
            assert nerr(pm.grad, po.grad) <= 6e-2, name


def test_single_atom_molecules_force_the_untiled_path():
    """More atoms than edges (single-atom molecules): not every atom row is a live bond row, the tile
    kernels decline, results still match the oracle."""
    from oracle.pyg_restated import Data as OData, collate as ocollate
    _, small = random_batch(seed=3, n_graphs=4, max_atoms=4)
    lone = [OData(x=torch.tensor([3], dtype=torch.int32), edge_index=torch.zeros(2, 0, dtype=torch.long),
                  edge_attr=torch.zeros(0, dtype=torch.int32), y=torch.zeros(1, 1)) for _ in range(40)]
    data = ocollate(small + lone)
    assert data.edge_index.size(1) < data.x.numel()
    oracle, model = make_pair(12, 7, 1, dtype=torch.bfloat16, embedding_dim=128, n_messages=2)
    oracle.eval(); model.eval()
    with torch.no_grad():
        ro, rm = oracle(data), model(_dev(data))
    for a, b in zip(rm, ro):
        assert nerr(a.float(), b) <= 2e-2
