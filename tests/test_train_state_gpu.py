"""-m gpu: `TrainStep` state handling: CUDA-graph cache across batch sizes,
checkpoint / resume through `state_dict()` (reference: best-epoch `deepcopy(model.state_dict())` and
`load_state_dict`, examples/comparison/train_cn.ipynb:220-226), device-side snapshot / restore."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from graphchem_b200.synthetic import make_molecules

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(dtype=torch.float32, seed=0, D=64, L=2):
    from graphchem_b200.nn import MoleculeGCN
    torch.manual_seed(seed)
    return MoleculeGCN(16, 8, 1, embedding_dim=D, n_messages=L, n_readout=2, readout_dim=32, compute_dtype=dtype).cuda()


def _batches(sizes, L=2, seed=3):
    store = make_molecules(sum(sizes), seed=seed, atom_vocab=16, bond_vocab=8, variable=True)
    out, lo = [], 0
    for n in sizes:
        out.append(store.collate_pack(np.arange(lo, lo + n), L, 16, 8).to("cuda"))
        lo += n
    return out


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_alternating_batch_sizes_graphed_equals_eager(dtype):
    """full / partial / full batches over several epochs with cached graphs (the last batch of an epoch is smaller):
    graphs captured for one batch size must stay valid when another size re-sizes the prediction buffers."""
    from graphchem_b200.train import TrainStep
    D = 128 if dtype == torch.bfloat16 else 64
    batches = _batches([96, 96, 41, 96, 17])
    results = []
    for graphed in (True, False):
        model = _model(dtype, D=D)
        step = TrainStep(model, lr=1e-2, use_cuda_graph=graphed, capture_after=1)
        canary = torch.full((4096,), 7.0, device="cuda")  # an allocation that would land in a freed prediction buffer
        losses = []
        for epoch in range(4):
            for pb in batches:
                losses.append(step(pb).clone())
            junk = [torch.zeros(96, 1, device="cuda") for _ in range(8)]  # recycle small blocks between epochs
            del junk
        torch.cuda.synchronize()
        assert bool((canary == 7.0).all())
        results.append((model.flat_parameters().clone(), torch.stack(losses).cpu()))
        if graphed:
            assert len(step._graphs) == len(batches)
    assert torch.equal(results[0][0], results[1][0])
    assert torch.equal(results[0][1], results[1][1])


def test_graphs_are_captured_on_second_sighting_and_lru_bounded():
    from graphchem_b200.train import TrainStep
    batches = _batches([32] * 6)
    step = TrainStep(_model(), max_graphs=4)
    for pb in batches:
        step(pb)
    assert len(step._graphs) == 0  # a streaming epoch of never-repeating batches pays no capture
    for pb in batches:
        step(pb)
    assert len(step._graphs) == 4  # least recently used graphs were evicted
    eager = TrainStep(_model(), use_cuda_graph=False)
    ref = TrainStep(_model(), max_graphs=4)
    for _ in range(3):
        for pb in batches:
            eager(pb); ref(pb)
    assert torch.equal(eager.model.flat_parameters(), ref.model.flat_parameters())


_RESUME = r"""
import sys, torch, numpy as np
sys.path.insert(0, {root!r})
from tests.test_train_state_gpu import _model, _batches
from graphchem_b200.train import TrainStep
state = torch.load({ckpt!r}, weights_only=False)
model = _model(torch.bfloat16, seed=99, D=128)   # different initial weights: everything must come from the checkpoint
step = TrainStep(model, lr=123.0)
step.load_state_dict(state)
batches = _batches([64, 64, 30])
for i in range(5, 12):
    step.lr = 1e-2 * (1 - 0.03 * i)
    step(batches[i % 3])
torch.save({{"flat": model.flat_parameters().cpu(), "exp_avg": step.exp_avg.cpu(), "steps": int(step.step_count.item())}}, {out!r})
"""


def test_checkpoint_resume_in_a_new_process_is_bit_identical(tmp_path):
    """save -> new process -> load -> continue == an uninterrupted run (parameters, Adam moments, step counter, lr)."""
    from graphchem_b200.train import TrainStep
    batches = _batches([64, 64, 30])
    model = _model(torch.bfloat16, D=128)
    step = TrainStep(model, lr=1e-2)
    ckpt, out = str(tmp_path / "ckpt.pt"), str(tmp_path / "resumed.pt")
    for i in range(12):
        if i == 5:
            torch.save(step.state_dict(), ckpt)
        step.lr = 1e-2 * (1 - 0.03 * i)
        step(batches[i % 3])
    torch.cuda.synchronize()
    res = subprocess.run([sys.executable, "-c", _RESUME.format(root=ROOT, ckpt=ckpt, out=out)], capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    got = torch.load(out, weights_only=False)
    assert got["steps"] == 12 == int(step.step_count.item())
    assert torch.equal(got["flat"], model.flat_parameters().cpu())
    assert torch.equal(got["exp_avg"], step.exp_avg.cpu())
    # the model half of the checkpoint is a plain reference-keyed state_dict
    sd = torch.load(ckpt, weights_only=False)["model"]
    assert set(sd) == set(model.state_dict())


def test_best_epoch_snapshot_restore():
    """train_cn.ipynb:220-226: remember the best epoch's weights, keep training, go back."""
    from graphchem_b200.train import TrainStep
    batches = _batches([48, 48])
    model = _model()
    step = TrainStep(model, lr=1e-2)
    for pb in batches:
        step(pb)
    snap = step.snapshot()
    best = model.flat_parameters().clone()
    sd_best = {k: v.clone() for k, v in model.state_dict().items()}
    for _ in range(3):
        for pb in batches:
            step(pb)
    assert not torch.equal(model.flat_parameters(), best)
    step.restore(snap)
    assert torch.equal(model.flat_parameters(), best)
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd_best[k]), k
    # the kernel-side weight copies follow: a forward after restore equals one on a fresh model with those weights
    from graphchem_b200.nn import MoleculeGCN
    fresh = MoleculeGCN(16, 8, 1, embedding_dim=64, n_messages=2, n_readout=2, readout_dim=32).cuda()
    fresh.load_state_dict(sd_best)
    model.eval(); fresh.eval()
    with torch.no_grad():
        assert torch.equal(model(batches[0])[0], fresh(batches[0])[0])
    # and the optimiser state went back with it: continuing reproduces the first continuation
    a = step(batches[0]).clone()
    step.restore(snap)
    b = step(batches[0]).clone()
    assert torch.equal(a, b)


def test_fused_adam_state_dict_resumes_bit_identically():
    """FusedAdam kept its moments outside Optimizer.state, so state_dict() dropped them.  A run that is
    checkpointed after two steps and resumed in a fresh optimiser must equal the uninterrupted run."""
    import copy
    import torch.nn.functional as F
    from graphchem_b200.optim import FusedAdam
    from tests.util import random_batch
    from graphchem_b200.data import Data
    data, _ = random_batch(seed=5, n_graphs=12, av=16, bv=8)
    dev = Data(x=data.x.cuda(), edge_index=data.edge_index.cuda(), edge_attr=data.edge_attr.cuda(),
               batch=data.batch.cuda(), y=data.y.cuda(), num_graphs=12)

    def steps(model, opt, n):
        for _ in range(n):
            opt.zero_grad()
            pred, _, _ = model(dev)
            F.mse_loss(pred, dev.y).backward()
            opt.step()

    ref = _model()
    opt = FusedAdam(ref, lr=1e-2)
    steps(ref, opt, 2)
    ckpt_model, ckpt_opt = copy.deepcopy(ref.state_dict()), copy.deepcopy(opt.state_dict())
    assert int(ckpt_opt["fused"]["step_count"]) == 2 and float(ckpt_opt["fused"]["exp_avg_sq"].abs().sum()) > 0
    steps(ref, opt, 3)
    resumed = _model(seed=9)
    resumed.load_state_dict(ckpt_model)
    opt2 = FusedAdam(resumed, lr=1e-2)
    opt2.load_state_dict(ckpt_opt)
    steps(resumed, opt2, 3)
    assert torch.equal(ref.flat_parameters(), resumed.flat_parameters())
