"""AdamW over the model's flat parameter / gradient buffers in one HBM-bound kernel per group,
with optional global-norm clipping — what HF Trainer does per step with torch.optim.AdamW
(trainer.py:1143-1217,1755-1760; defaults training_args.py:778-862: betas (0.9,0.999), eps 1e-8,
weight_decay 0.0, max_grad_norm 1.0; biases and LayerNorm weights excluded from decay,
trainer.py:1280-1290).  The same kernel rewrites the bf16 weight shadow, so no separate cast
pass runs after the step.  No host synchronisation: the clip coefficient stays on the device.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class VitkAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr: float = 2e-5, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 max_grad_norm: Optional[float] = None):
        lay = model.layout
        decay = [p for n, p in model.named_parameters() if lay.kinds[n] != "nodecay"]
        nodecay = [p for n, p in model.named_parameters() if lay.kinds[n] == "nodecay"]
        super().__init__([{"params": decay, "weight_decay": weight_decay}, {"params": nodecay, "weight_decay": 0.0}],
                         dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.model = model
        self.max_grad_norm = max_grad_norm
        # step() also clears the flat gradient buffer (same pass over it), so zero_grad() after step() is free:
        # param.grad stay bound to the (now zero) views and the next backward needs no memset.
        self.fused_zero_grad = True
        self._grads_clean = False
        self._step = 0
        self._m = self._v = self._ss = self._scale = None

    def _init_state(self):
        flat = self.model.flat_parameters()
        self._m, self._v = torch.zeros_like(flat), torch.zeros_like(flat)
        self._ss = torch.zeros(1, dtype=torch.float32, device=flat.device)
        self._scale = torch.ones(1, dtype=torch.float32, device=flat.device)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        model, lay = self.model, self.model.layout
        p, g = model.flat_parameters(), model.flat_grads()
        if self._m is None or self._m.device != p.device or self._m.numel() != p.numel():
            self._init_state()
        if any(q.grad is not None and q.grad.data_ptr() != model.engine()._grad_views[n].data_ptr()
               for n, q in model.named_parameters()):
            model.engine()._bind_grads()          # foreign .grad tensors: fold them into the flat buffer
        self._step += 1
        t = self._step
        scale = None
        if self.max_grad_norm is not None:
            self._ss.zero_()
            ops.sumsq(g, self._ss)
            ops.clip_scale(self._ss, float(self.max_grad_norm), self._scale)
            scale = self._scale
        shadow = model.shadow()
        g0, g1 = self.param_groups
        b1, b2 = g0["betas"]
        bc1, bc2 = 1.0 - b1 ** t, 1.0 - b2 ** t
        ge, de, n = lay.gemm_end, lay.decay_end, lay.total
        # GEMM weights: decayed, bf16 shadow rewritten in the same pass
        z = self.fused_zero_grad
        ops.adamw(p[:ge], g[:ge], self._m[:ge], self._v[:ge], shadow, ge, g0["lr"], b1, b2, g0["eps"], g0["weight_decay"], bc1, bc2,
                  scale, z)
        if de > ge:
            ops.adamw(p[ge:de], g[ge:de], self._m[ge:de], self._v[ge:de], None, de - ge, g0["lr"], b1, b2, g0["eps"],
                      g0["weight_decay"], bc1, bc2, scale, z)
        b1n, b2n = g1["betas"]
        ops.adamw(p[de:], g[de:], self._m[de:], self._v[de:], None, n - de, g1["lr"], b1n, b2n, g1["eps"], 0.0,
                  1.0 - b1n ** t, 1.0 - b2n ** t, scale, z)
        self._grads_clean = z
        model.mark_shadow_fresh()
        return loss

    def zero_grad(self, set_to_none: bool = True):
        if self._grads_clean:
            self._grads_clean = False      # already zeroed by step(); keep .grad bound to the flat views
            return
        super().zero_grad(set_to_none=set_to_none)

    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm of the last ``step`` (device scalar; only with max_grad_norm)."""
        return self._ss.sqrt()
