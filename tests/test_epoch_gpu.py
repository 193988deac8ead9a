"""-m gpu: the epoch runner (`MoleculeStore.collate_pack_epoch` + `TrainStep.run_epoch` = one `gcb_train_steps` call per
epoch, the notebook loop of reference examples/predict_cn.ipynb:243-252 without per-step host work) against the same steps
taken one by one."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,kw", [(torch.float32, {}), (torch.float32, dict(aggr="max", p_dropout=0.0)), (torch.bfloat16, {})])
def test_run_epoch_is_bit_identical_to_stepping(dtype, kw):
    from graphchem_b200.nn import MoleculeGCN
    from graphchem_b200.synthetic import make_molecules
    from graphchem_b200.train import TrainStep
    store = make_molecules(200, seed=9, variable=True)
    index_sets = [np.arange(i, min(200, i + 16)) for i in range(0, 200, 16)]   # 13 batches of different shapes, the last of 8
    D = 128 if dtype == torch.bfloat16 else 64

    def new():
        torch.manual_seed(3)
        return MoleculeGCN(64, 32, 1, embedding_dim=D, n_messages=2, n_readout=2, readout_dim=32, compute_dtype=dtype, **kw).cuda().train()

    a, b = new(), new()
    sa, sb = TrainStep(a, lr=1e-3, use_cuda_graph=False), TrainStep(b, lr=1e-3, use_cuda_graph=False)
    for epoch in range(3):
        sa.lr = sb.lr = 1e-3 - epoch * 1e-5
        order = index_sets[::-1] if epoch == 1 else index_sets
        la = torch.stack([sa(store.collate_pack(idx, 2, 64, 32).to("cuda")).reshape(()).clone() for idx in order])
        lb = sb.run_epoch(store.collate_pack_epoch(order, 2, 64, 32, "cuda"))
        assert torch.equal(la, lb), (epoch, la, lb)
    torch.cuda.synchronize()
    assert torch.equal(a.flat_parameters(), b.flat_parameters())
    assert torch.equal(sa.exp_avg, sb.exp_avg) and torch.equal(sa.exp_avg_sq, sb.exp_avg_sq)
    assert int(sa.step_count.item()) == int(sb.step_count.item()) == 39
    assert sb.run_epoch([]).numel() == 0
