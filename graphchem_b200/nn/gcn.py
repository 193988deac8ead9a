"""Drop-in `MoleculeGCN` whose forward/backward run as hand-written sm_100a CUDA kernels behind
the C-ABI of include/graphchem_b200.h.

Mirrors reference `graphchem/nn/gcn.py:10-202`: same constructor signature (:35-42), same
attributes (`_p_dropout`, `_n_messages`, `act_fn`, `emb_atom`, `emb_bond`, `atom_conv`,
`bond_conv`, `readout`), same `state_dict()` keys/shapes (:78-114) and the same
`forward(data) -> (out, out_atom, out_bond)` contract (:120-202), including the EdgeConv-on-the-
bond-matrix behaviour of :163.  There is no CPU or eager-PyTorch fallback: without a
CUDA device or without the built library the model raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import warnings
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from .. import _lib
from ..data.packed import PackedBatch, pack_batch

_ACT_IDS = {
    F.softplus: _lib.ACT_SOFTPLUS,
    F.sigmoid: _lib.ACT_SIGMOID, torch.sigmoid: _lib.ACT_SIGMOID,
    F.relu: _lib.ACT_RELU, torch.relu: _lib.ACT_RELU,
    F.tanh: _lib.ACT_TANH, torch.tanh: _lib.ACT_TANH,
    F.leaky_relu: _lib.ACT_LEAKY_RELU,
}


def _act_id(act_fn) -> Optional[int]:
    """Kernel id of `act_fn`: one of the five functional activations above, or the `nn.Module` form of
    the same function with its default hyper-parameters (`nn.Softplus()`, `nn.LeakyReLU()` ...).  The
    reference takes any callable (gcn.py:66-68); here the activation is compiled into the kernels."""
    try:
        if act_fn in _ACT_IDS:
            return _ACT_IDS[act_fn]
    except TypeError:  # unhashable callable
        return None
    if isinstance(act_fn, nn.Softplus) and act_fn.beta == 1 and act_fn.threshold == 20:
        return _lib.ACT_SOFTPLUS
    if isinstance(act_fn, nn.LeakyReLU) and act_fn.negative_slope == 0.01:
        return _lib.ACT_LEAKY_RELU
    for cls, aid in ((nn.Sigmoid, _lib.ACT_SIGMOID), (nn.ReLU, _lib.ACT_RELU), (nn.Tanh, _lib.ACT_TANH)):
        if isinstance(act_fn, cls):
            return aid
    return None


_POISON = os.environ.get("GCB_DEBUG_POISON", "0") == "1"
_AGGR_IDS = {"add": _lib.AGGR_ADD, "sum": _lib.AGGR_ADD, "mean": _lib.AGGR_MEAN, "max": _lib.AGGR_MAX}


class _Linear(nn.Linear):
    """nn.Linear with PyG `Linear`'s initialisation bounds (identical to torch's)."""


class GeneralConvParams(nn.Module):
    """Parameter container with PyG `GeneralConv(D, D, D)`'s names: `lin_msg`, `lin_self`
    (Identity because in == out and skip_linear=False), `lin_edge` (reference gcn.py:84-86)."""

    def __init__(self, channels: int, aggr: str):
        super().__init__()
        self.in_channels = self.out_channels = self.in_edge_channels = channels
        self.aggr = aggr
        self.lin_msg = _Linear(channels, channels)
        self.lin_self = nn.Identity()
        self.lin_edge = _Linear(channels, channels)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("atom_conv is evaluated inside MoleculeGCN.forward by fused CUDA kernels")

    def extra_repr(self) -> str:
        return f"{self.in_channels}, {self.out_channels}, aggr={self.aggr}"


class EdgeConvParams(nn.Module):
    """Parameter container with PyG `EdgeConv(nn)`'s names: `nn.0.{weight,bias}` (gcn.py:89-91)."""

    def __init__(self, mlp: nn.Module):
        super().__init__()
        self.nn = mlp
        self.aggr = "max"

    def forward(self, *args, **kwargs):
        raise NotImplementedError("bond_conv is evaluated inside MoleculeGCN.forward by fused CUDA kernels")


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _GCNFunction(torch.autograd.Function):
    """autograd bridge: one C-ABI call forward, one backward."""

    @staticmethod
    def forward(ctx, model: "MoleculeGCN", packed: PackedBatch, flags: int, seed: int, *params: Tensor):
        out, out_atom, out_bond, ws = model._run_forward(packed, flags, seed)
        ctx.model, ctx.packed, ctx.flags, ctx.seed, ctx.ws = model, packed, flags, seed, ws
        ctx.set_materialize_grads(False)
        return out, out_atom, out_bond

    @staticmethod
    def backward(ctx, d_out, d_out_atom, d_out_bond):
        model: MoleculeGCN = ctx.model
        if ctx.ws is None:
            raise RuntimeError("backward through MoleculeGCN called twice (activations already released)")
        grads = model._run_backward(ctx.packed, ctx.flags, ctx.seed, ctx.ws, d_out, d_out_atom, d_out_bond)
        ctx.ws = None
        return (None, None, None, None) + tuple(grads)


class MoleculeGCN(nn.Module):
    """See module docstring.  Extra keywords (not in the reference): `compute_dtype`
    (torch.float32 = parity mode, default; torch.bfloat16 = bf16 storage / fp32 accumulate) and `pool`
    ("add" = the reference's global_add_pool, gcn.py:181; "mean" = PyG's global_mean_pool)."""

    def __init__(self, atom_vocab_size: int, bond_vocab_size: int, output_dim: Optional[int],
                 embedding_dim: Optional[int] = 128, n_messages: Optional[int] = 2,
                 n_readout: Optional[int] = 2, readout_dim: Optional[int] = 64,
                 p_dropout: Optional[float] = 0.0, aggr: Optional[str] = "add",
                 act_fn: Optional[Callable] = F.softplus, *, compute_dtype: torch.dtype = torch.float32,
                 pool: str = "add"):
        super().__init__()
        self._p_dropout = p_dropout
        self._n_messages = n_messages
        self.act_fn = act_fn
        self._act_id = _act_id(act_fn)
        if self._act_id is None:
            raise ValueError("act_fn must be one of torch.nn.functional.{softplus, sigmoid, relu, tanh, leaky_relu} "
                             "(or the nn.Module of the same function with default arguments): "
                             "the activation is fused into the CUDA kernels")
        if aggr not in _AGGR_IDS:
            raise NotImplementedError(f"aggr={aggr!r} is not implemented by the CUDA kernels (use 'add', 'mean' or 'max')")
        if pool not in ("add", "sum", "mean"):
            raise ValueError("pool must be 'add' (global_add_pool, the reference) or 'mean' (global_mean_pool)")
        self._pool = _lib.POOL_MEAN if pool == "mean" else _lib.POOL_ADD
        if compute_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("compute_dtype must be torch.float32 or torch.bfloat16")
        self._aggr = aggr
        self._compute_dtype = compute_dtype
        self._dims = (int(atom_vocab_size), int(bond_vocab_size), output_dim, int(embedding_dim), int(n_readout),
                      int(readout_dim))

        self.emb_atom = nn.Embedding(atom_vocab_size, embedding_dim)
        self.emb_bond = nn.Embedding(bond_vocab_size, embedding_dim)
        self.atom_conv = GeneralConvParams(embedding_dim, aggr)
        self.bond_conv = EdgeConvParams(nn.Sequential(_Linear(2 * embedding_dim, embedding_dim)))
        if n_readout > 0:
            self.readout = nn.ModuleList()
            self.readout.append(nn.Sequential(_Linear(embedding_dim, readout_dim)))
            for _ in range(n_readout - 1):
                self.readout.append(nn.Sequential(_Linear(readout_dim, readout_dim)))
            self.readout.append(nn.Sequential(_Linear(readout_dim, output_dim)))
        else:
            self.readout = None

        self._flat: Optional[Tensor] = None
        self._derived: Optional[Tensor] = None
        self._cfg_cache = None

    # ---- configuration / parameter plumbing ------------------------------------------------------
    def _cfg(self) -> _lib.ModelCfg:
        av, bv, out_dim, D, n_readout, R = self._dims
        key = (self._compute_dtype, float(self._p_dropout), int(self._n_messages))
        if self._cfg_cache is None or self._cfg_cache[0] != key:
            cfg = _lib.ModelCfg(av, bv, int(out_dim) if (n_readout > 0 and out_dim is not None) else 0, D,
                                int(self._n_messages), n_readout, R, self._act_id, _AGGR_IDS[self._aggr],
                                _lib.BF16 if self._compute_dtype == torch.bfloat16 else _lib.F32,
                                float(self._p_dropout), self._pool)
            offs = (C.c_int64 * (_lib.MAX_PARAM_TENSORS + 1))()
            n = _lib.check(_lib.lib().gcb_param_layout(C.byref(cfg), offs), "gcb_param_layout")
            self._cfg_cache = (key, cfg, [int(offs[i]) for i in range(n + 1)])
        return self._cfg_cache[1]

    def _apply(self, fn, *args, **kwargs):
        self._params_cache = None  # .to() / .cuda() / .float() may replace the Parameter objects
        return super()._apply(fn, *args, **kwargs)

    def _ordered_params(self) -> List[nn.Parameter]:
        # cached: walking the module tree costs ~40 us, and the small-batch training loop asks several times per step
        ps = getattr(self, "_params_cache", None)
        if ps is not None:
            return ps
        ps = [self.emb_atom.weight, self.emb_bond.weight,
              self.atom_conv.lin_msg.weight, self.atom_conv.lin_msg.bias,
              self.atom_conv.lin_edge.weight, self.atom_conv.lin_edge.bias,
              self.bond_conv.nn[0].weight, self.bond_conv.nn[0].bias]
        if self.readout is not None:
            for layer in self.readout:
                ps += [layer[0].weight, layer[0].bias]
        self._params_cache = ps
        return ps

    def flat_parameters(self) -> Tensor:
        """The flat fp32 parameter vector (C-ABI `gcb_param_layout` order) that all parameters alias."""
        self._cfg()
        offs = self._cfg_cache[2]
        params = self._ordered_params()
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("MoleculeGCN parameters must live on a CUDA device (no CPU fallback)")
        flat = self._flat
        ok = flat is not None and flat.device == dev
        if ok:
            base = flat.data_ptr()
            for p, off in zip(params, offs):
                if p.dtype != torch.float32 or p.data_ptr() != base + 4 * off or p.device != dev:
                    ok = False
                    break
        if not ok:
            with torch.no_grad():
                flat = torch.empty(offs[-1], dtype=torch.float32, device=dev)
                for p, off in zip(params, offs):
                    if p.dtype != torch.float32:
                        raise RuntimeError("MoleculeGCN keeps fp32 master parameters; use compute_dtype for bf16")
                    view = flat[off:off + p.numel()].view(p.shape)
                    view.copy_(p.data)
                    p.data = view
            self._flat = flat
        return flat

    def _derived_buf(self, device: torch.device) -> Tensor:
        nbytes = int(_lib.check(int(_lib.lib().gcb_derived_bytes(C.byref(self._cfg()))), "gcb_derived_bytes"))
        if self._derived is None or self._derived.device != device or self._derived.numel() != nbytes:
            self._derived = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self._derived

    def refresh_weights(self) -> None:
        """Re-derive the kernel-side weight copies from the master parameters (one tiny kernel)."""
        flat = self.flat_parameters()
        dv = self._derived_buf(flat.device)
        _lib.check(_lib.lib().gcb_prepare_weights(C.byref(self._cfg()), flat.data_ptr(), dv.data_ptr(),
                                                  _stream_ptr(flat.device)), "gcb_prepare_weights")

    def _ensure_device(self) -> torch.device:
        p = self.emb_atom.weight
        if p.device.type != "cuda":
            if not torch.cuda.is_available():
                raise RuntimeError("graphchem_b200.MoleculeGCN needs a CUDA device (B200, sm_100a); "
                                   "there is no CPU fallback")
            warnings.warn("MoleculeGCN parameters were on the CPU; moving the model to cuda:0", stacklevel=3)
            self.to("cuda")
        return self.emb_atom.weight.device

    # ---- packing ---------------------------------------------------------------------------------
    def pack(self, data, pin: bool = True) -> PackedBatch:
        """Host-side prepack of a PyG-style Batch for this model (cached on the object)."""
        av, bv = self._dims[0], self._dims[1]
        tensors = (data.x, data.edge_index, data.edge_attr, getattr(data, "batch", None))
        sig = (int(self._n_messages) >= 1, av, bv) + tuple(
            (t.data_ptr(), t._version, tuple(t.shape)) if isinstance(t, Tensor) else None for t in tensors)
        cached = getattr(data, "__dict__", {}).get("_gcb_packed")
        if cached is not None and cached[0] == sig:
            return cached[1]
        packed = pack_batch(data, int(self._n_messages), av, bv, pin=pin)
        try:
            object.__setattr__(data, "_gcb_packed", (sig, packed))
        except Exception:
            pass
        return packed

    # ---- C-ABI calls -----------------------------------------------------------------------------
    def _run_forward(self, packed: PackedBatch, flags: int, seed: int):
        cfg = self._cfg()
        flat = self.flat_parameters()
        dev = flat.device
        with torch.cuda.device(dev):
            self.refresh_weights()
            av, bv, out_dim, D, n_readout, R = self._dims
            N, E, G = packed.N, packed.E, packed.G
            ws_bytes = int(_lib.check(int(_lib.lib().gcb_workspace_bytes(C.byref(cfg), N, E, G, flags)),
                                      "gcb_workspace_bytes"))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            if _POISON:  # debug: make any read of unwritten workspace show up as NaN
                ws.fill_(0xFF)
            out_cols = int(out_dim) if n_readout > 0 else D
            out = torch.empty(G, out_cols, dtype=torch.float32, device=dev)
            out_atom = torch.empty(N, D, dtype=self._compute_dtype, device=dev)
            out_bond = torch.empty(E, D, dtype=self._compute_dtype, device=dev)
            rc = _lib.lib().gcb_forward(C.byref(cfg), packed.ints.data_ptr(), packed.header.ctypes.data,
                                        flat.data_ptr(), self._derived.data_ptr(), ws.data_ptr(), ws_bytes, flags,
                                        seed, out.data_ptr(), out_atom.data_ptr(), out_bond.data_ptr(),
                                        _stream_ptr(dev))
            _lib.check(rc, "gcb_forward")
        return out, out_atom, out_bond, (ws if flags & _lib.SAVE_FOR_BWD else None)

    def _run_backward(self, packed: PackedBatch, flags: int, seed: int, ws: Tensor, d_out, d_out_atom, d_out_bond):
        cfg = self._cfg()
        flat = self.flat_parameters()
        dev = flat.device
        offs = self._cfg_cache[2]
        params = self._ordered_params()

        def prep(t, dtype):
            return None if t is None else t.to(device=dev, dtype=dtype).contiguous()

        d_out, d_out_atom, d_out_bond = (prep(d_out, torch.float32), prep(d_out_atom, self._compute_dtype),
                                         prep(d_out_bond, self._compute_dtype))
        with torch.cuda.device(dev):
            grad = torch.empty(offs[-1], dtype=torch.float32, device=dev)
            rc = _lib.lib().gcb_backward(
                C.byref(cfg), packed.ints.data_ptr(), packed.header.ctypes.data, flat.data_ptr(),
                self._derived.data_ptr(), ws.data_ptr(), ws.numel(), flags, seed,
                None if d_out is None else d_out.data_ptr(), None if d_out_atom is None else d_out_atom.data_ptr(),
                None if d_out_bond is None else d_out_bond.data_ptr(), grad.data_ptr(), 0, _stream_ptr(dev))
            _lib.check(rc, "gcb_backward")
        return [grad[off:off + p.numel()].view(p.shape) for p, off in zip(params, offs)]

    # ---- batched per-molecule prediction ---------------------------------------------------------
    @torch.no_grad()
    def predict_each(self, molecules, batch_size: int = 4096, return_all: bool = False):
        """What `[model(g) for g in molecules]` returns (the notebooks' evaluation loops, reference
        examples/predict_cn.ipynb:258-263, 333-334), computed in batched launches.

        A plain batched forward is NOT that: the reference indexes the bond matrix by atom id
        (gcn.py:163), so molecules of one Batch see each other's bond rows.  Here the
        batch is packed so that every molecule sees exactly the bond rows it would see alone
        (GCB_PACK_EACH, include/graphchem_b200.h); dropout is off as under `model.eval()`.

        `molecules`: a MoleculeStore, a MoleculeDataset or a sequence of Data/MoleculeGraph.  Returns the
        predictions [n, out] (CPU), or with `return_all` also the per-molecule lists of out_atom / out_bond."""
        from ..data.packed import MoleculeStore
        dev = self._ensure_device()
        store = molecules if isinstance(molecules, MoleculeStore) else MoleculeStore.from_graphs(list(molecules))
        if int(self._n_messages) < 1:
            raise NotImplementedError("predict_each needs n_messages >= 1 (without message passing a plain batch is already per-molecule)")
        av, bv = self._dims[0], self._dims[1]
        n = len(store)
        cols = self._dims[2] if self._dims[4] > 0 else self._dims[3]
        D = self._dims[3]
        out = torch.zeros(n, cols)
        atoms: List[Optional[Tensor]] = [None] * n
        bonds: List[Optional[Tensor]] = [None] * n
        na_all, ne_all = np.asarray(store.n_atoms), np.asarray(store.n_edges)
        # The "each alone" layout gives atom i of a molecule the molecule's bond row i, so it needs at least as many
        # edges as atoms per molecule.  Molecules with fewer (a lone atom, salts / mixtures with unbonded atoms) are
        # rare: each of them runs as its own plain one-molecule batch, which IS the reference's `model(g)`
        # (including its IndexError when an edge addresses a bond row that does not exist, gcn.py:163).
        odd = np.nonzero(ne_all < na_all)[0]
        regular = np.nonzero(ne_all >= na_all)[0]
        for i in range(0, len(regular), batch_size):
            idx = regular[i:i + batch_size]
            pb = store.collate_pack(idx, int(self._n_messages), av, bv, each=True).to(dev, non_blocking=True)
            o, out_atom, out_bond, _ = self._run_forward(pb, 0, 0)
            out[torch.from_numpy(idx)] = o.cpu()
            if return_all:
                for k, a, b in zip(idx.tolist(), torch.split(out_atom.cpu(), na_all[idx].tolist()),
                                   torch.split(out_bond.cpu(), ne_all[idx].tolist())):
                    atoms[k], bonds[k] = a, b
        for k in odd.tolist():
            pb = store.collate_pack([k], int(self._n_messages), av, bv).to(dev, non_blocking=True)
            o, out_atom, out_bond, _ = self._run_forward(pb, 0, 0)
            out[k] = o.cpu()[0]
            if return_all:
                atoms[k], bonds[k] = out_atom.cpu(), out_bond.cpu().reshape(-1, D)
        return (out, atoms, bonds) if return_all else out

    # ---- forward ---------------------------------------------------------------------------------
    def forward(self, data) -> Tuple[Tensor, Tensor, Tensor]:
        """`data`: PyG-style Data/Batch (x, edge_index, edge_attr, batch) or a `PackedBatch`.
        Returns (out, out_atom, out_bond) on the device of the input (reference gcn.py:202)."""
        dev = self._ensure_device()
        if isinstance(data, PackedBatch):
            packed, in_dev = data, data.device
        else:
            x = data.x
            in_dev = x.device if isinstance(x, Tensor) else torch.device("cpu")
            packed = self.pack(data)
        if packed.device != dev:
            packed = packed.to(dev, non_blocking=True)
        params = self._ordered_params()
        self.flat_parameters()
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        flags = (_lib.TRAINING if self.training else 0) | (_lib.SAVE_FOR_BWD if need_grad else 0)
        seed = 0
        if self.training and self._p_dropout > 0:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
        if need_grad:
            out, out_atom, out_bond = _GCNFunction.apply(self, packed, flags, seed, *params)
        else:
            out, out_atom, out_bond, _ = self._run_forward(packed, flags, seed)
        if in_dev != dev and not isinstance(data, PackedBatch):
            out, out_atom, out_bond = out.to(in_dev), out_atom.to(in_dev), out_bond.to(in_dev)
        return out, out_atom, out_bond
