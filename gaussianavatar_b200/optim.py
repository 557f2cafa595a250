"""Fused Adam with torch.optim.Adam's semantics and state layout (reference: model/avatar_model.py:148-155 builds
`torch.optim.Adam` over net + geo_feature; `step()` at :264-267).  One kernel launch per parameter tensor; with the flat
parameter buffer of `POP_no_unet` that is two launches per step.  `grad_scale` folds the 1/world_size of the
data-parallel all-reduce(sum) into the update."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ptr


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.grad_scale = 1.0
        self.skip_flag = None      # device int32 tensor: a non-zero word makes the update a no-op (rasterizer overflow guard)

    @torch.no_grad()
    def step(self, closure=None):
        L = _lib.lib()
        st = torch.cuda.current_stream().cuda_stream
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or not p.is_contiguous() or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters")
                s = self.state[p]
                if len(s) == 0:
                    s["step"] = 0
                    s["exp_avg"] = torch.zeros_like(p)
                    s["exp_avg_sq"] = torch.zeros_like(p)
                s["step"] = int(s["step"]) + 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                _lib.check(L.ga_adam_step(p.numel(), ptr(p), ptr(g), ptr(s["exp_avg"]), ptr(s["exp_avg_sq"]), float(group["lr"]),
                                          float(b1), float(b2), float(group["eps"]), s["step"], float(self.grad_scale),
                                          ptr(self.skip_flag), st),
                           "ga_adam_step")
        return None
