"""-m gpu: parity at the BENCHMARKED shapes (BASELINE.json configs[2]/[3]: 4096 and 8192 synthetic molecules of
23 atoms / 50 directed edges, D = 128, 4 message steps, bf16; configs[4]: D = 256, 6 message steps, forward only).

These sizes exercise what the 64-graph parity tests cannot: multi-wave persistent tcgen05 GEMMs (1472 row tiles
over 148 CTAs), ~1700 closed row tiles, the weight-gradient partial reduction over 148 CTAs, the register-tiled
readout.  The checker is the CPU oracle (fp32, all host threads: a few seconds per case); the product path is
the fused `TrainStep` pieces through the C-ABI (`TrainStep.local_gradient` = forward + MSE + backward) and the
module forward.

Tolerances (norm-wise, max|a-b| / max|oracle|): bf16 forward 1e-2 for up to 4 message steps (north star), 1.5e-2
at 6 steps and D = 256 (error grows with depth: every layer rounds its activations to bf16 once); bf16 gradients
5e-2.  Measured values: profiles/r2_parity_errors.jsonl (profiles/micro/parity_errors.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import nerr

pytestmark = pytest.mark.gpu

FWD_BF16_L4, FWD_BF16_L6, GRAD_BF16 = 1e-2, 1.5e-2, 5e-2


def synthetic_case(n_mol, D, L, seed=4321):
    """(oracle batch, PackedBatch on the GPU, oracle model, CUDA model) of the bench workload."""
    from graphchem_b200.nn import MoleculeGCN
    from graphchem_b200.synthetic import make_molecules
    from oracle.gcn_oracle import OracleMoleculeGCN
    from oracle.pyg_restated import Data as OData, collate
    store = make_molecules(n_mol, seed=seed)
    graphs = []
    for i in range(n_mol):
        a0, a1 = store.atom_off[i], store.atom_off[i + 1]
        e0, e1 = store.edge_off[i], store.edge_off[i + 1]
        graphs.append(OData(x=store.atoms[a0:a1], edge_attr=store.bonds[e0:e1],
                            edge_index=store.conn[2 * e0:2 * e1].view(2, -1), y=store.targets[i:i + 1]))
    data = collate(graphs)
    torch.manual_seed(0)
    oracle = OracleMoleculeGCN(64, 32, 1, embedding_dim=D, n_messages=L, n_readout=2, readout_dim=64)
    model = MoleculeGCN(64, 32, 1, embedding_dim=D, n_messages=L, n_readout=2, readout_dim=64, compute_dtype=torch.bfloat16)
    model.load_state_dict(oracle.state_dict())
    model = model.cuda()
    pb = store.collate_pack(np.arange(n_mol), L, 64, 32).to("cuda")
    return data, pb, oracle, model


def forward_errors(data, pb, oracle, model):
    oracle.eval(); model.eval()
    with torch.no_grad():
        ro = oracle(data)
        rm = model._run_forward(pb, 0, 0)
    return [nerr(a.float(), b) for a, b in zip(rm[:3], ro)]


def gradient_errors(data, pb, oracle, model):
    from graphchem_b200.train import TrainStep
    oracle.train(); model.train()
    oracle.zero_grad()
    lo = F.mse_loss(oracle(data)[0], data.y)
    lo.backward()
    step = TrainStep(model, use_cuda_graph=False)
    flat_grad = step.local_gradient(pb).clone()
    torch.cuda.synchronize()
    errs = {}
    off = 0
    for (name, po), pm in zip(oracle.named_parameters(), model._ordered_params()):
        g = flat_grad[off:off + pm.numel()].view(pm.shape)
        off += pm.numel()
        errs[name] = nerr(g, po.grad)
    return errs, abs(float(step.loss.item()) - float(lo)) / abs(float(lo))


@pytest.mark.parametrize("n_mol", [4096, 8192])
def test_bench_shape_bf16_forward_and_all_gradients(n_mol):
    data, pb, oracle, model = synthetic_case(n_mol, 128, 4)
    assert pb.n_tiles > 1400 * (n_mol // 4096) // 2  # the tile kernels and multi-wave GEMMs are the path under test
    errs = forward_errors(data, pb, oracle, model)
    assert max(errs) <= FWD_BF16_L4, errs
    gerrs, loss_err = gradient_errors(data, pb, oracle, model)
    assert loss_err <= 2e-2, loss_err
    bad = {k: v for k, v in gerrs.items() if v > GRAD_BF16}
    assert not bad, (bad, gerrs)


def test_screening_shape_bf16_forward():
    """configs[4]: D = 256, 6 message steps, forward only, 4096 molecules (streamed-weight tcgen05 GEMMs)."""
    data, pb, oracle, model = synthetic_case(4096, 256, 6)
    errs = forward_errors(data, pb, oracle, model)
    assert max(errs) <= FWD_BF16_L6, errs


def test_wide_model_bf16_training_gradients():
    """D = 256 training (2048 molecules, 4 message steps): the [256 x 512] / [512 x 256] weight gradients come from the
    column-blocked tcgen05 TN kernel; all gradients against the CPU oracle."""
    data, pb, oracle, model = synthetic_case(2048, 256, 4)
    errs = forward_errors(data, pb, oracle, model)
    assert max(errs) <= FWD_BF16_L4, errs
    gerrs, loss_err = gradient_errors(data, pb, oracle, model)
    assert loss_err <= 2e-2, loss_err
    bad = {k: v for k, v in gerrs.items() if v > GRAD_BF16}
    assert not bad, (bad, gerrs)
