"""`FusedAdam` — torch.optim.Optimizer-compatible front end of `vl3d_adam_step`.

Same maths and defaults as the optimiser the reference builds in `MPMeshVid.get_optimizer`
(MPV.py:200-218: Adam(betas=(0.9,0.999), eps=6e-8), no weight decay / amsgrad): parameters whose
`.grad` is None are skipped (uvs, uvs_dyn and _verts never receive one), parameters
with an all-zero gradient are still updated by their momentum, `param_group['lr']` may be changed
between steps (train_3dvid.py:281-287).
"""
from __future__ import annotations

import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)        # preserve_format keeps the texel layout
                    st["exp_avg_sq"] = torch.zeros_like(p)
                if tuple(g.stride()) != tuple(p.stride()):
                    g2 = torch.empty_like(p)
                    g2.copy_(g)
                    g = g2
                st["step"] += 1
                ops.adam_step(p.data, g, st["exp_avg"], st["exp_avg_sq"], st["step"], group["lr"], b1, b2,
                              group["eps"])
        return loss
