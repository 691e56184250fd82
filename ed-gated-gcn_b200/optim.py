"""Adam for the parameters of the gated block as ONE kernel launch per step (``edg_adam_multi``).

The reference trains with ``torch.optim.Adam`` (train.py:239-243).  Same update rule here (bias-corrected, ``eps``
added to ``sqrt(v / bc2)``, optional L2 weight decay, no amsgrad); the step counter is a device scalar, so a step
captured in a CUDA graph replays correctly.  CUDA fp32 parameters only; no fallback."""
from __future__ import annotations

import ctypes
from typing import Iterable, List

import torch

from . import _lib as L


class FusedAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdam got no trainable parameters")
        dev = self.params[0].device
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                raise L.EdgError("FusedAdam takes contiguous fp32 CUDA parameters on one device (there is no CPU path)")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        sizes = [p.numel() for p in self.params]
        flat = torch.zeros(2 * sum(sizes), dtype=torch.float32, device=dev)          # exp_avg | exp_avg_sq, one allocation
        self.exp_avg = list(flat[:sum(sizes)].split(sizes))
        self.exp_avg_sq = list(flat[sum(sizes):].split(sizes))
        self.step_count = torch.zeros(len(self.params), dtype=torch.float32, device=dev)   # per tensor, as in torch

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0) -> None:
        items = []
        for i, (p, m, v) in enumerate(zip(self.params, self.exp_avg, self.exp_avg_sq)):
            if p.grad is None:
                continue
            g = p.grad
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.float().contiguous()
            items.append((p, g, m, v, self.step_count[i]))
        for i0 in range(0, len(items), 32):
            part = items[i0:i0 + 32]
            n = len(part)
            P, I64 = ctypes.c_void_p * n, ctypes.c_int64 * n
            L.call("edg_adam_multi", n, *[P(*[L.ptr(it[k]) for it in part]) for k in range(5)],
                   I64(*[it[0].numel() for it in part]), self.lr, self.betas[0], self.betas[1], self.eps,
                   self.weight_decay, float(grad_scale), L.stream())
