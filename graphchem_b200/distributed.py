"""Data-parallel plumbing (new work: the reference has no distributed code).

Molecules (whole graphs) are sharded across ranks; every rank collates its own PackedBatch, so the
EdgeConv-on-bond-matrix coupling of reference gcn.py:163 acts within the rank-local batch.  The only
exchange step is one all-reduce (sum) of the flat fp32 gradient per step; with equal per-rank batch
sizes, scaling the sum by 1/world gives the gradient of the global-mean MSE loss.
"""
from __future__ import annotations

import os
from typing import Sequence

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend: str = "nccl") -> tuple:
    """(rank, world, local_rank) from torchrun's environment; initialises the process group."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kwargs)
    return rank, world, local


def shard_indices(n_items: int, rank: int, world: int, per_rank_batch: int, step: int, seed: int = 0,
                  shuffle: bool = True) -> np.ndarray:
    """Molecule indices of `rank`'s batch at `step`: a global permutation per epoch (same on every
    rank, derived from `seed` and the epoch), cut into world*per_rank_batch slices, rank-major."""
    global_batch = per_rank_batch * world
    steps_per_epoch = max(n_items // global_batch, 1)
    epoch, k = divmod(step, steps_per_epoch)
    order = np.random.default_rng(seed + epoch).permutation(n_items) if shuffle else np.arange(n_items)
    lo = k * global_batch + rank * per_rank_batch
    return order[lo:lo + per_rank_batch]


def allreduce_sum_(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """Sum-all-reduce the flat gradient in place (`TrainStep` folds 1/world into its Adam kernel, so the captured
    step has no extra scaling launch).  NCCL on CUDA tensors (captured into the step's CUDA graph), gloo on CPU."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def allreduce_mean_(flat_grad: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """Sum-all-reduce the flat gradient in place and scale by 1/world (mean of per-rank means)."""
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        flat_grad.mul_(1.0 / world)
    return flat_grad
