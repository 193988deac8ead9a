"""Fused Adam over the model's flat parameter vector (one kernel for all 14-16 tensors).

Replaces `torch.optim.Adam(model.parameters(), lr=1e-3)` + `optimizer.step()` of the notebook loop
(reference examples/predict_cn.ipynb:230-232, 251).  `param_groups[0]["lr"]` stays mutable, as the
per-epoch rewrite of predict_cn.ipynb:241-242 requires; `torch.optim.Adam` itself also keeps
working on `model.parameters()` (the parameters are ordinary nn.Parameters).

Differences from `torch.optim.Adam`, by design of the flat update: a parameter whose `.grad` is None is treated as
having a zero gradient (its moments still decay and move it), where torch skips it; the moments live in three flat
buffers (`exp_avg`, `exp_avg_sq`, `step_count`) that `state_dict()` / `load_state_dict()` carry under the key
"fused" next to the usual `param_groups`, so a checkpoint / resume keeps them.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, model, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8):
        self.model = model
        params = list(model.parameters())
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._state_ready = False

    def _init_state(self, flat: torch.Tensor) -> None:
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=flat.device)
        self.hyper = torch.zeros(2, dtype=torch.float32, device=flat.device)
        self._hyper_host = None
        self._state_ready = True

    def _flat_grad(self, flat: torch.Tensor) -> torch.Tensor:
        params = self.model._ordered_params()
        grads = [p.grad for p in params]
        if any(g is None for g in grads):
            grads = [torch.zeros_like(p) if g is None else g for p, g in zip(params, grads)]
        base = grads[0].data_ptr()
        off, contiguous = 0, True
        for g in grads:
            if g.dtype != torch.float32 or not g.is_contiguous() or g.data_ptr() != base + 4 * off:
                contiguous = False
                break
            off += g.numel()
        if contiguous and grads[0]._base is not None and grads[0]._base.numel() >= off and grads[0]._base.data_ptr() == base:
            return grads[0]._base
        return torch.cat([g.reshape(-1).float() for g in grads])

    def state_dict(self):
        sd = super().state_dict()
        if self._state_ready:
            sd["fused"] = {"exp_avg": self.exp_avg.detach().clone(), "exp_avg_sq": self.exp_avg_sq.detach().clone(),
                           "step_count": int(self.step_count.item())}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        fused = state_dict.pop("fused", None)
        super().load_state_dict(state_dict)
        if fused is None:
            self._state_ready = False  # a checkpoint taken before the first step: moments start from zero again
            return
        flat = self.model.flat_parameters()
        if fused["exp_avg"].numel() != flat.numel():
            raise ValueError(f"optimizer state holds {fused['exp_avg'].numel()} elements, the model has {flat.numel()}")
        self._init_state(flat)
        self.exp_avg.copy_(fused["exp_avg"].to(flat.device))
        self.exp_avg_sq.copy_(fused["exp_avg_sq"].to(flat.device))
        self.step_count.fill_(int(fused["step_count"]))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        flat = self.model.flat_parameters()
        if not self._state_ready or self.exp_avg.device != flat.device or self.exp_avg.numel() != flat.numel():
            self._init_state(flat)
        group = self.param_groups[0]
        hyper = (float(group["lr"]), 1.0)
        if self._hyper_host != hyper:
            self.hyper.copy_(torch.tensor(hyper, dtype=torch.float32))
            self._hyper_host = hyper
        grad = self._flat_grad(flat)
        b1, b2 = group["betas"]
        rc = _lib.lib().gcb_adam_step(flat.data_ptr(), grad.data_ptr(), self.exp_avg.data_ptr(),
                                      self.exp_avg_sq.data_ptr(), flat.numel(), self.hyper.data_ptr(),
                                      self.step_count.data_ptr(), b1, b2, group["eps"],
                                      torch.cuda.current_stream(flat.device).cuda_stream)
        _lib.check(rc, "gcb_adam_step")
        return loss
