"""Fused training step: the whole notebook loop body (reference examples/predict_cn.ipynb:247-251:
zero_grad / model(batch) / F.mse_loss / loss.backward() / optimizer.step()) as one sequence of
C-ABI calls on one CUDA stream, captured into a CUDA graph per resident batch buffer, with the
data-parallel gradient all-reduce (new) between backward and Adam.

No autograd, no per-step host synchronisation (the reference's `loss.item()` every step,
predict_cn.ipynb:252, is replaced by a device-side loss scalar the caller reads when it wants).
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .data.packed import PackedBatch
from .distributed import allreduce_sum_


class TrainStep:
    """One fused training step per call (see module docstring).

    Device state: the model's flat fp32 parameter vector, Adam's `exp_avg` / `exp_avg_sq`, the step counter and
    the learning rate all live in device memory at fixed addresses, so captured CUDA graphs stay valid across
    `lr` changes, `load_state_dict` and `restore`.  `state_dict()` / `load_state_dict()` carry everything a resume
    needs (the reference keeps `deepcopy(model.state_dict())` of the best epoch,
    examples/comparison/train_cn.ipynb:220-226; optimiser state is added here so that a resumed run continues
    bit-identically); `snapshot()` / `restore()` are the device-side form of that best-epoch copy.

    CUDA graphs: a batch buffer is run eagerly the first `capture_after - 1` times it is seen and captured on the
    next sighting, so streaming loaders whose batches never repeat pay no capture cost; the cache is LRU with
    `max_graphs` entries (raise it to the number of resident batches)."""

    def __init__(self, model, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 use_cuda_graph: bool = True, process_group=None, capture_after: int = 2, max_graphs: int = 64):
        self.model = model
        self.betas, self.eps = betas, eps
        dev = model._ensure_device()
        self.device = dev
        flat = self._flat_now = model.flat_parameters()
        self._flat_ptr = flat.data_ptr()
        self.grad = torch.zeros_like(flat)
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.world = 1
        self.pg = process_group
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        self.hyper = torch.tensor([lr, 1.0 / self.world], dtype=torch.float32, device=dev)
        self._lr = lr
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.use_cuda_graph = use_cuda_graph
        self.capture_after = max(1, int(capture_after))
        self._ws: Optional[torch.Tensor] = None
        self._pred_buf: Optional[torch.Tensor] = None   # [G_max, out]; `_pred` / `_dpred` are its leading rows
        self._dpred_buf: Optional[torch.Tensor] = None
        self._pred: Optional[torch.Tensor] = None
        self._dpred: Optional[torch.Tensor] = None
        self._graphs: "OrderedDict[tuple, torch.cuda.CUDAGraph]" = OrderedDict()
        self._keepalive: Dict[tuple, PackedBatch] = {}
        self._seen: "OrderedDict[tuple, int]" = OrderedDict()
        self._warm = False
        self.max_graphs = max_graphs
        self.launches_per_step = 0
        model.refresh_weights()

    # lr stays mutable between steps (predict_cn.ipynb:241-242) without re-capturing the graph
    @property
    def lr(self) -> float:
        return self._lr

    @lr.setter
    def lr(self, value: float) -> None:
        self._lr = float(value)
        self.hyper[0:1].fill_(self._lr)

    # ---- checkpoint / resume ----------------------------------------------------
    def state_dict(self) -> dict:
        """Everything a bit-identical resume needs, as host tensors: model parameters (reference keys), Adam
        moments over the flat parameter vector, step counter, learning rate and the Adam constants."""
        return {
            "model": {k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()},
            "exp_avg": self.exp_avg.detach().cpu().clone(),
            "exp_avg_sq": self.exp_avg_sq.detach().cpu().clone(),
            "step_count": int(self.step_count.item()),
            "lr": float(self._lr), "betas": tuple(self.betas), "eps": float(self.eps),
        }

    def load_state_dict(self, state: dict) -> None:
        self.model.load_state_dict(state["model"])
        self._rebind()
        self.exp_avg.copy_(state["exp_avg"])
        self.exp_avg_sq.copy_(state["exp_avg_sq"])
        self.step_count.fill_(int(state["step_count"]))
        self.betas, self.eps = tuple(state["betas"]), float(state["eps"])
        self.lr = float(state["lr"])
        self.model.refresh_weights()

    def snapshot(self) -> dict:
        """Device-side copy of the training state (one D2D copy per tensor, no host sync): the best-epoch
        `deepcopy(model.state_dict())` of examples/comparison/train_cn.ipynb:220."""
        return {"flat": self.model.flat_parameters().clone(), "exp_avg": self.exp_avg.clone(),
                "exp_avg_sq": self.exp_avg_sq.clone(), "step_count": self.step_count.clone(), "lr": self._lr}

    def restore(self, snap: dict, optimizer_state: bool = True) -> None:
        """Back to a `snapshot()` (train_cn.ipynb:226 `model.load_state_dict(best_state)`); addresses are kept,
        captured graphs stay valid."""
        self._rebind()
        self.model.flat_parameters().copy_(snap["flat"])
        if optimizer_state:
            self.exp_avg.copy_(snap["exp_avg"])
            self.exp_avg_sq.copy_(snap["exp_avg_sq"])
            self.step_count.copy_(snap["step_count"])
            self.lr = snap["lr"]
        self.model.refresh_weights()

    def _rebind(self) -> None:
        """The captured graphs hold the address of the flat parameter vector: if the model re-flattened its
        parameters (`.to()`, externally replaced `.data`), drop them."""
        flat = self._flat_now = self.model.flat_parameters()  # verified once per step; the enqueue helpers reuse it
        if flat.data_ptr() != self._flat_ptr:
            self._flat_ptr = flat.data_ptr()
            self._drop_graphs()

    def _drop_graphs(self) -> None:
        # captured graphs hold stale addresses; the sighting counts stay valid (a batch seen before is still a repeat)
        self._graphs.clear()
        self._keepalive.clear()

    def _buffers(self, pb: PackedBatch, flags: int):
        cfg = self.model._cfg()
        need = int(_lib.check(int(_lib.lib().gcb_workspace_bytes(C.byref(cfg), pb.N, pb.E, pb.G, flags)),
                              "gcb_workspace_bytes"))
        if self._ws is None or self._ws.numel() < need:
            self._drop_graphs()
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out_cols = self.model._dims[2] if self.model._dims[4] > 0 else self.model._dims[3]
        if self._pred_buf is None or self._pred_buf.shape[0] < pb.G or self._pred_buf.shape[1] != out_cols:
            # graphs captured earlier hold the old buffers' addresses: they go with them
            self._drop_graphs()
            self._pred_buf = torch.empty(pb.G, out_cols, dtype=torch.float32, device=self.device)
            self._dpred_buf = torch.empty_like(self._pred_buf)
        # leading rows of the capacity buffers: the same base address for every batch size
        self._pred, self._dpred = self._pred_buf[:pb.G], self._dpred_buf[:pb.G]

    def _enqueue_grad(self, pb: PackedBatch) -> None:
        """Forward, MSE and backward of one batch on the current stream: `self.loss`, `self.grad` (this rank's)."""
        lib, model = _lib.lib(), self.model
        cfg = model._cfg()
        flat = self._flat_now
        flags = _lib.TRAINING | _lib.SAVE_FOR_BWD
        s = torch.cuda.current_stream(self.device).cuda_stream
        seed = 0
        if model._p_dropout > 0:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
        hdr = pb.header.ctypes.data
        _lib.check(lib.gcb_forward(C.byref(cfg), pb.ints.data_ptr(), hdr, flat.data_ptr(), model._derived.data_ptr(),
                                   self._ws.data_ptr(), self._ws.numel(), flags, seed, self._pred.data_ptr(), None,
                                   None, s), "gcb_forward")
        _lib.check(lib.gcb_mse_loss(self._pred.data_ptr(), pb.y.data_ptr(), self._pred.numel(), self.loss.data_ptr(),
                                    self._dpred.data_ptr(), 1.0, s), "gcb_mse_loss")
        _lib.check(lib.gcb_backward(C.byref(cfg), pb.ints.data_ptr(), hdr, flat.data_ptr(), model._derived.data_ptr(),
                                    self._ws.data_ptr(), self._ws.numel(), flags, seed, self._dpred.data_ptr(), None,
                                    None, self.grad.data_ptr(), 0, s), "gcb_backward")

    def _enqueue_update(self) -> None:
        """Data-parallel gradient sum, Adam on the flat parameter vector, refresh of the derived weights."""
        lib, model = _lib.lib(), self.model
        cfg = model._cfg()
        flat = self._flat_now
        s = torch.cuda.current_stream(self.device).cuda_stream
        if self.world > 1:
            allreduce_sum_(self.grad, group=self.pg)  # 1/world is folded into the Adam kernel (hyper[1])
        b1, b2 = self.betas
        _lib.check(lib.gcb_adam_step(flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                     self.exp_avg_sq.data_ptr(), flat.numel(), self.hyper.data_ptr(),
                                     self.step_count.data_ptr(), b1, b2, self.eps, s), "gcb_adam_step")
        _lib.check(lib.gcb_prepare_weights(C.byref(cfg), flat.data_ptr(), model._derived.data_ptr(), s),
                   "gcb_prepare_weights")

    def _enqueue(self, pb: PackedBatch) -> None:
        """Enqueue one full training step on the current stream."""
        self._enqueue_grad(pb)
        self._enqueue_update()

    def local_gradient(self, pb: PackedBatch) -> torch.Tensor:
        """This rank's flat gradient of the batch's MSE loss (no collective, no parameter update); the building
        block of the data-parallel parity check (tests/dist_trainstep_check.py)."""
        self._rebind()
        self._buffers(pb, _lib.TRAINING | _lib.SAVE_FOR_BWD)
        self._enqueue_grad(pb)
        self._warm = True
        return self.grad

    def run_epoch(self, batches) -> torch.Tensor:
        """One optimiser step per batch of `batches` (device-resident PackedBatches, in order), enqueued by ONE library
        call (`gcb_train_steps`): the notebook loop `for batch in loader: ...` of examples/predict_cn.ipynb:243-252 without
        the per-step Python / ctypes work, which is what bounds the 16-molecule regime.  Bit-identical to calling the
        step once per batch (with dropout active the masks differ: batch b is seeded with seed0 + b, one draw per epoch).  Returns the per-batch losses [n] as a device tensor (no synchronisation).  Single process
        only: with a process group the batches are stepped one by one (the all-reduce lives in `__call__`)."""
        batches = list(batches)
        if self.world > 1:
            return torch.stack([self(pb).reshape(()).clone() for pb in batches])
        n = len(batches)
        losses = torch.empty(n, dtype=torch.float32, device=self.device)
        if n == 0:
            return losses
        for pb in batches:
            if pb.device != self.device or pb.y is None:
                raise ValueError("run_epoch needs device-resident PackedBatches with targets")
        self._rebind()
        lib, model = _lib.lib(), self.model
        cfg = model._cfg()
        flags = _lib.TRAINING | _lib.SAVE_FOR_BWD
        need = max(int(_lib.check(int(lib.gcb_workspace_bytes(C.byref(cfg), pb.N, pb.E, pb.G, flags)), "gcb_workspace_bytes"))
                   for pb in batches)
        if self._ws is None or self._ws.numel() < need:
            self._drop_graphs()
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        g_max = max(pb.G for pb in batches)
        out_cols = model._dims[2] if model._dims[4] > 0 else model._dims[3]
        if self._pred_buf is None or self._pred_buf.shape[0] < g_max or self._pred_buf.shape[1] != out_cols:
            self._drop_graphs()
            self._pred_buf = torch.empty(g_max, out_cols, dtype=torch.float32, device=self.device)
            self._dpred_buf = torch.empty_like(self._pred_buf)
        ptrs = (C.c_void_p * n)(*[pb.ints.data_ptr() for pb in batches])
        hdrs = (C.c_void_p * n)(*[pb.header.ctypes.data for pb in batches])
        ys = (C.c_void_p * n)(*[pb.y.data_ptr() for pb in batches])
        seed0 = 0
        if model._p_dropout > 0:
            seed0 = int(torch.empty((), dtype=torch.int64).random_().item()) & ((1 << 62) - 1)
        flat = self._flat_now
        b1, b2 = self.betas
        s = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(lib.gcb_train_steps(C.byref(cfg), n, ptrs, hdrs, ys, flat.data_ptr(), model._derived.data_ptr(),
                                       self._ws.data_ptr(), self._ws.numel(), seed0, self._pred_buf.data_ptr(),
                                       self._dpred_buf.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                       self.exp_avg_sq.data_ptr(), self.hyper.data_ptr(), self.step_count.data_ptr(), b1, b2,
                                       self.eps, losses.data_ptr(), s), "gcb_train_steps")
        self._warm = True
        return losses

    def _capture(self, pb: PackedBatch) -> torch.cuda.CUDAGraph:
        if not self._warm:
            # one eager step on a side stream before the first capture (lazy one-time initialisation inside the
            # library must not happen under capture), undone so that it does not change the training trajectory
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            tensors = (self.model.flat_parameters(), self.exp_avg, self.exp_avg_sq, self.step_count)
            state = [t.clone() for t in tensors]
            with torch.cuda.stream(side):
                self._enqueue(pb)
            torch.cuda.current_stream(self.device).wait_stream(side)
            for t, s0 in zip(tensors, state):
                t.copy_(s0)
            self.model.refresh_weights()
            self._warm = True
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._enqueue(pb)
        return g

    def __call__(self, pb: PackedBatch) -> torch.Tensor:
        """Run one step on a device-resident PackedBatch; returns the device loss scalar (no sync)."""
        if pb.device != self.device:
            raise RuntimeError("TrainStep needs a device-resident PackedBatch (use .to(device) / a prefetcher)")
        if pb.y is None:
            raise ValueError("PackedBatch has no targets")
        self._rebind()
        self._buffers(pb, _lib.TRAINING | _lib.SAVE_FOR_BWD)
        graphable = self.use_cuda_graph and self.model._p_dropout == 0
        if not graphable:
            self._enqueue(pb)
            self._warm = True
            return self.loss
        key = (pb.ints.data_ptr(), pb.y.data_ptr(), pb.header.tobytes())
        g = self._graphs.get(key)
        if g is not None:
            self._graphs.move_to_end(key)
            g.replay()
            return self.loss
        seen = self._seen.get(key, 0) + 1
        if seen < self.capture_after:  # not yet worth a capture: run eagerly
            self._seen[key] = seen
            self._seen.move_to_end(key)
            while len(self._seen) > 4 * max(self.max_graphs, 1):
                self._seen.popitem(last=False)
            self._enqueue(pb)
            self._warm = True
            return self.loss
        self._seen.pop(key, None)
        g = self._capture(pb)
        while len(self._graphs) >= max(self.max_graphs, 1):  # least recently used graph goes
            old, _ = self._graphs.popitem(last=False)
            self._keepalive.pop(old, None)
        self._graphs[key] = g
        self._keepalive[key] = pb
        g.replay()
        return self.loss


class InferStep:
    """Forward-only counterpart of `TrainStep` for screening (BASELINE.json configs[4]): `model.eval()`
    semantics (no dropout), no saved activations, workspace and outputs allocated once, one CUDA graph per
    resident batch buffer.  `step(pb)` returns the predictions [G, out] as a device tensor that is OVERWRITTEN
    by the next call with a batch of the same shape (clone it, or pass `out=`)."""

    def __init__(self, model, use_cuda_graph: bool = True, want_atom_bond: bool = False):
        self.model = model
        self.device = model._ensure_device()
        self.use_cuda_graph = use_cuda_graph
        self.want_atom_bond = want_atom_bond
        self._ws: Optional[torch.Tensor] = None
        self._out: Dict[tuple, tuple] = {}
        self._graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        self._keepalive: Dict[tuple, PackedBatch] = {}
        self.max_graphs = 64
        model.refresh_weights()

    def refresh_weights(self) -> None:
        """Call after the parameters changed (the kernel-side weight copies are derived from them)."""
        self.model.refresh_weights()

    def _buffers(self, pb: PackedBatch):
        model = self.model
        cfg = model._cfg()
        need = int(_lib.check(int(_lib.lib().gcb_workspace_bytes(C.byref(cfg), pb.N, pb.E, pb.G, 0)), "gcb_workspace_bytes"))
        if self._ws is None or self._ws.numel() < need:
            self._graphs.clear(); self._keepalive.clear()
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        key = (pb.N, pb.E, pb.G)
        if key not in self._out:
            av, bv, out_dim, D, n_readout, R = model._dims
            cols = int(out_dim) if n_readout > 0 else D
            out = torch.empty(pb.G, cols, dtype=torch.float32, device=self.device)
            oa = torch.empty(pb.N, D, dtype=model._compute_dtype, device=self.device) if self.want_atom_bond else None
            ob = torch.empty(pb.E, D, dtype=model._compute_dtype, device=self.device) if self.want_atom_bond else None
            self._out[key] = (out, oa, ob)
        return self._out[key]

    def _enqueue(self, pb: PackedBatch, bufs) -> None:
        model = self.model
        out, oa, ob = bufs
        flat = model.flat_parameters()
        s = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().gcb_forward(C.byref(model._cfg()), pb.ints.data_ptr(), pb.header.ctypes.data, flat.data_ptr(),
                                          model._derived.data_ptr(), self._ws.data_ptr(), self._ws.numel(), 0, 0,
                                          out.data_ptr(), None if oa is None else oa.data_ptr(),
                                          None if ob is None else ob.data_ptr(), s), "gcb_forward")

    def __call__(self, pb: PackedBatch):
        if pb.device != self.device:
            raise RuntimeError("InferStep needs a device-resident PackedBatch (use .to(device) / PackedLoader)")
        bufs = self._buffers(pb)
        if not self.use_cuda_graph:
            self._enqueue(pb, bufs)
        else:
            key = (pb.ints.data_ptr(), pb.header.tobytes())
            g = self._graphs.get(key)
            if g is None:
                side = torch.cuda.Stream(self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    self._enqueue(pb, bufs)  # warm-up before capture
                torch.cuda.current_stream(self.device).wait_stream(side)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue(pb, bufs)
                if len(self._graphs) >= self.max_graphs:
                    old = next(iter(self._graphs))
                    del self._graphs[old]
                    self._keepalive.pop(old, None)
                self._graphs[key] = g
                self._keepalive[key] = pb
            g.replay()
        return bufs if self.want_atom_bond else bufs[0]


def screening_shard(n_items: int, rank: int, world: int):
    """Contiguous block of molecule indices that `rank` of `world` screens (BASELINE.json configs[4]: independent
    molecules, replicas only, no collective; blocks differ in size by at most one)."""
    import numpy as np
    if not 0 <= rank < world:
        raise ValueError("rank must be in [0, world)")
    lo = (n_items * rank) // world
    hi = (n_items * (rank + 1)) // world
    return np.arange(lo, hi, dtype=np.int64)


class Screener:
    """Inference-only screening (BASELINE.json configs[4]) with its set-up kept between calls: the pinned host
    pool, the ring of device slots, the pack scratch, the forward workspace and one CUDA graph per slot are built
    by the first call and reused by every later call with the same batch geometry (cudaHostAlloc of the pool and
    the graph captures cost more than screening a million molecules).  `screen()` is the one-shot form.

    `screener(store, indices=None, rank=0, world=1)` returns the predictions [n, out] (host tensor) of the molecules
    `indices` of a `MoleculeStore` (all of them by default; with `world > 1` this rank's `screening_shard` of them),
    in order.  Host threads gather raw batches, the GPU packs them (`device_pack`) and runs the forward pass only,
    `model.eval()` semantics (no dropout); no collective is involved, predictions stay with the rank.

    As with any batched forward of the reference, molecules of one batch see each other's bond rows (gcn.py:163,
    SURVEY.md 3.3): results equal `model(batch)` on the same batches, not `predict_each`.  CUDA graphs are used when
    every batch has the same shape (`use_cuda_graph=None`: decided per call from the store)."""

    def __init__(self, model, batch_size: int = 32768, workers: int = 8, prefetch: int = 4, device_pack: bool = True,
                 use_cuda_graph: Optional[bool] = None):
        self.model, self.batch_size = model, int(batch_size)
        self.workers, self.prefetch, self.device_pack = workers, prefetch, device_pack
        self.use_cuda_graph = use_cuda_graph
        self._buffers: Dict[bool, dict] = {}      # graphed? -> loader buffer set
        self._step: Optional[InferStep] = None

    @torch.no_grad()
    def __call__(self, store, indices=None, rank: int = 0, world: int = 1) -> torch.Tensor:
        import numpy as np
        from .data.loader import PackedLoader
        model = self.model
        dev = model._ensure_device()
        idx = np.arange(len(store), dtype=np.int64) if indices is None else np.asarray(indices, dtype=np.int64)
        if world > 1:
            idx = idx[screening_shard(len(idx), rank, world)]
        n = int(idx.shape[0])
        av, bv, out_dim, D, n_readout, R = model._dims
        cols = int(out_dim) if n_readout > 0 else D
        preds = torch.empty(n, cols, dtype=torch.float32, device=dev)
        if n == 0:
            return preds.cpu()
        L = int(model._n_messages)
        graphed = self.use_cuda_graph
        if graphed is None:  # one graph per device slot pays only when all (full) batches share one shape
            na, ne = store.n_atoms[idx], store.n_edges[idx]
            graphed = bool((na == na[0]).all() and (ne == ne[0]).all()) and n >= self.batch_size
        loader = PackedLoader(store, self.batch_size, L, av, bv, device=dev, workers=self.workers, prefetch=self.prefetch,
                              indices=idx, device_slots=3 if graphed else 0, device_pack=self.device_pack,
                              buffers=self._buffers.get(graphed), lookahead=False)  # one epoch per loader
        if graphed and self._step is None:
            self._step = InferStep(model, use_cuda_graph=True)
        step = self._step if graphed else None
        if step is not None:
            step.refresh_weights()
        was_training = model.training
        model.eval()
        try:
            k = 0
            for pb in loader:
                out = step(pb) if step is not None else model._run_forward(pb, 0, 0)[0]
                preds[k:k + pb.G].copy_(out)
                k += pb.G
        finally:
            model.train(was_training)
            self._buffers[graphed] = loader.export_buffers()
        return preds.cpu()


def screen(model, store, batch_size: int = 32768, indices=None, rank: int = 0, world: int = 1, workers: int = 8,
           prefetch: int = 4, device_pack: bool = True, use_cuda_graph: Optional[bool] = None) -> torch.Tensor:
    """One-shot `Screener` (see there): predictions [n, out] (host tensor) of the molecules `indices` of `store`."""
    return Screener(model, batch_size, workers, prefetch, device_pack, use_cuda_graph)(store, indices, rank, world)
