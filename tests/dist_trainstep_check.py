"""Data-parallel parity of `graphchem_b200.train.TrainStep` under NCCL (the reference has no
distributed code, so the check is against a single-GPU run of the same kernels):

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_trainstep_check.py

Every rank trains K steps on ITS OWN batches (seed 1234 + rank, as bench.py shards them) through the captured
CUDA graph that contains the NCCL all-reduce.  Then

  (1) the flat parameter vector, Adam moments and step counter must be BIT-IDENTICAL on all ranks;
  (2) rank 0 replays the same K steps alone: per step it computes the gradient of every rank's batch with the
      same kernels (`TrainStep.local_gradient`), sums them in rank order, and applies Adam with the 1/world
      factor.  For world = 2 the all-reduce is one fp32 addition per element (a + b == b + a), so the replay must
      equal the data-parallel run BIT FOR BIT; for larger worlds NCCL's reduction order is its own and the
      comparison is to 1e-6 relative.

Prints one JSON line on rank 0 and exits non-zero on any mismatch.  Run by tests/test_distributed_gpu.py when
two GPUs are visible, and by hand through `gpurun --gpus 2` (log under profiles/)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def main() -> int:
    from graphchem_b200.distributed import init_from_env
    from graphchem_b200.nn import MoleculeGCN
    from graphchem_b200.synthetic import make_molecules
    from graphchem_b200.train import TrainStep

    rank, world, local = init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    B = int(os.environ.get("GCB_DP_BATCH", "1024"))
    K = int(os.environ.get("GCB_DP_STEPS", "6"))
    n_res = 3
    D, L = 128, 3

    def new_model():
        torch.manual_seed(0)
        return MoleculeGCN(64, 32, 1, embedding_dim=D, n_messages=L, n_readout=2, readout_dim=64,
                           compute_dtype=torch.bfloat16).to(dev)

    def batches_of(r):
        store = make_molecules(n_res * B, seed=1234 + r)
        return [store.collate_pack(np.arange(i * B, (i + 1) * B), L, 64, 32).to(dev) for i in range(n_res)]

    # ---- the product path: graphed step with the all-reduce inside ----
    model = new_model()
    if world > 1:
        dist.broadcast(model.flat_parameters(), src=0)
    init_flat = model.flat_parameters().clone()
    mine = batches_of(rank)
    step = TrainStep(model, lr=1e-3, capture_after=1)
    losses = []
    for k in range(K):
        step.lr = 1e-3 - k * 1e-5  # lr is rewritten between steps in the notebooks (predict_cn.ipynb:241-242)
        losses.append(step(mine[k % n_res]).clone())
    torch.cuda.synchronize()
    flat = model.flat_parameters().clone()
    state = torch.cat([flat, step.exp_avg, step.exp_avg_sq, step.step_count.float()])

    ok = True
    report = {"world": world, "batch_per_rank": B, "steps": K, "D": D, "L": L}
    # (1) identical replicas
    if world > 1:
        gathered = [torch.empty_like(state) for _ in range(world)]
        dist.all_gather(gathered, state)
        same = all(torch.equal(gathered[0], g) for g in gathered[1:])
        report["replicas_bit_identical"] = bool(same)
        ok &= same
    # (2) single-GPU replay on rank 0
    if rank == 0:
        ref = new_model()
        ref.flat_parameters().copy_(init_flat)
        ref.refresh_weights()
        solo = TrainStep(ref, lr=1e-3, use_cuda_graph=False)
        solo.world = 1                      # the replay runs on this rank alone: no collective (the other ranks wait below)
        solo.hyper[1:2].fill_(1.0 / world)  # ... but the same 1/world factor in Adam
        all_batches = [batches_of(r) for r in range(world)]
        for k in range(K):
            solo.lr = 1e-3 - k * 1e-5
            total = None
            for r in range(world):
                g = solo.local_gradient(all_batches[r][k % n_res]).clone()
                total = g if total is None else total + g
            solo.grad.copy_(total)
            solo._enqueue_update()
        torch.cuda.synchronize()
        ref_flat = ref.flat_parameters()
        diff = float((ref_flat - flat).abs().max())
        scale = float(ref_flat.abs().max())
        report["replay_max_abs_diff"] = diff
        report["replay_bit_identical"] = bool(torch.equal(ref_flat, flat))
        report["moments_bit_identical"] = bool(torch.equal(solo.exp_avg, step.exp_avg) and torch.equal(solo.exp_avg_sq, step.exp_avg_sq))
        report["params_moved"] = float((flat - init_flat).abs().max())
        report["loss_first_last"] = [float(losses[0]), float(losses[-1])]
        if world <= 2:
            ok &= report["replay_bit_identical"] and report["moments_bit_identical"]
        else:
            ok &= diff <= 1e-6 * max(scale, 1.0)
        ok &= report["params_moved"] > 0
    flag = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        report["ok"] = bool(flag.item())
        print(json.dumps(report), flush=True)
    rc = 0 if bool(flag.item()) else 1
    if world > 1:
        step._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
