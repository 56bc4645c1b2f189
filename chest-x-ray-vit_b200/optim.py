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
        # step() also clears the flat gradient buffer (same pass over it), so the next backward needs no memset.
        self.fused_zero_grad = True
        self._step = 0
        self._m = self._v = self._ss = self._scale = None
        self._seg_key = None
        self._segs = None

    def _init_state(self):
        flat = self.model.flat_parameters()
        self._m, self._v = torch.zeros_like(flat), torch.zeros_like(flat)
        self._ss = torch.zeros(1, dtype=torch.float32, device=flat.device)
        self._scale = torch.ones(1, dtype=torch.float32, device=flat.device)
        self._ss_scratch = torch.zeros(int(ops._lib.lib().vitk_sumsq_scratch_floats()), dtype=torch.float32, device=flat.device)

    def _segments(self, active):
        """Contiguous [start, end) ranges of the flat buffer to update, split at the gemm / decay / no-decay
        boundaries: (start, end, group index, has bf16 shadow).  ``active[i]`` says whether parameter i (HF order)
        takes part in this step (torch.optim skips parameters whose grad is None, e.g. frozen ones)."""
        key = tuple(active)
        if key == self._seg_key:
            return self._segs
        lay = self.model.layout
        ge, de = lay.gemm_end, lay.decay_end
        if all(active):
            spans = [(0, lay.total)]
        else:
            names = sorted((n for n, a in zip(lay.names, active) if a), key=lambda n: lay.offset[n])
            ends = {n: e for n, e in zip(sorted(lay.names, key=lambda n: lay.offset[n]),
                                         [lay.offset[n] for n in sorted(lay.names, key=lambda n: lay.offset[n])[1:]] + [lay.total])}
            spans = []
            for n in names:                       # padded extent of each active parameter, merged when adjacent
                s, e = lay.offset[n], ends[n]
                if spans and spans[-1][1] == s:
                    spans[-1] = (spans[-1][0], e)
                else:
                    spans.append((s, e))
        segs = []
        for s, e in spans:
            for lo, hi, grp, sh in ((0, ge, 0, True), (ge, de, 0, False), (de, lay.total, 1, False)):
                a, b = max(s, lo), min(e, hi)
                if b > a:
                    segs.append((a, b, grp, sh))
        self._seg_key, self._segs = key, segs
        return segs

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        model, lay = self.model, self.model.layout
        p, g = model.flat_parameters(), model.flat_grads()
        if self._m is None or self._m.device != p.device or self._m.numel() != p.numel():
            self._init_state()
        model.engine().bind_grads()               # gradients living outside the flat buffer are folded into it
        plist = model.param_list()
        active = [q.grad is not None for q in plist]
        if not any(active):
            return loss
        segs = self._segments(active)
        whole = len(segs) == 3 and segs[0][0] == 0 and segs[-1][1] == lay.total
        self._step += 1
        self._enqueue(g, p, segs, whole, self._step, None)
        model._grads_clean = self.fused_zero_grad and whole
        model.mark_shadow_fresh()
        return loss

    def _enqueue(self, g, p, segs, whole, t, dev_state):
        """Launches of one step: (‖g‖² → clip coefficient) → AdamW per segment.  ``dev_state`` = (step counter int64 [1],
        bias corrections fp32 [2]) makes the step count device-resident (vitk_adamw_tick) so that the launches can be
        captured once in a CUDA graph and replayed; otherwise ``t`` is the host-side step count."""
        model = self.model
        scale = None
        if self.max_grad_norm is not None:
            ops.fill_zero(self._ss)
            if whole:
                ops.sumsq(g, self._ss, self._ss_scratch)
            else:
                for a, b, _, _ in segs:
                    ops.sumsq(g[a:b], self._ss, self._ss_scratch)
            ops.clip_scale(self._ss, float(self.max_grad_norm), self._scale)
            scale = self._scale
        shadow = model._flat_shadow if dev_state is not None else model.shadow()
        z = self.fused_zero_grad
        bc_dev = None
        if dev_state is not None:
            if any(gr["betas"] != self.param_groups[0]["betas"] for gr in self.param_groups):
                raise NotImplementedError("VitkAdamW graph replay: all parameter groups must share betas")
            b1, b2 = self.param_groups[0]["betas"]
            ops.adamw_tick(dev_state[0], True, b1, b2, dev_state[1])
            bc_dev = dev_state[1]
        for a, b, grp, has_shadow in segs:
            gr = self.param_groups[grp]
            b1, b2 = gr["betas"]
            ops.adamw(p[a:b], g[a:b], self._m[a:b], self._v[a:b], shadow[a:b] if has_shadow else None, b - a, gr["lr"], b1, b2,
                      gr["eps"], gr["weight_decay"], 1.0 - b1 ** t, 1.0 - b2 ** t, scale, z, bc_dev)

    def zero_grad(self, set_to_none: bool = True):
        """``set_to_none=True`` (the default, and what HF Trainer uses): drop the ``param.grad`` references; the flat
        buffer behind them was already cleared by ``step()``, so the next backward starts without a memset."""
        if set_to_none:
            for q in self.model.param_list():
                q.grad = None
        else:
            super().zero_grad(set_to_none=False)

    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm of the last ``step`` (device scalar; only with max_grad_norm)."""
        return self._ss.sqrt()

    # ------------------------------------------------------------------ checkpointing (Trainer writes optimizer.pt)
    def state_dict(self):
        """torch.optim layout (``state`` + ``param_groups``) with the flat moments stored once under ``vitk``:
        ``exp_avg`` / ``exp_avg_sq`` are the flat fp32 buffers (FlatLayout order), ``step`` the shared counter."""
        sd = super().state_dict()
        sd["vitk"] = {"step": self._step, "layout_total": self.model.layout.total,
                      "exp_avg": None if self._m is None else self._m.detach().clone(),
                      "exp_avg_sq": None if self._v is None else self._v.detach().clone(),
                      "max_grad_norm": self.max_grad_norm}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        vk = state_dict.pop("vitk", None)
        super().load_state_dict(state_dict)
        if vk is None:
            raise ValueError("VitkAdamW.load_state_dict: not a VitkAdamW state dict (no 'vitk' entry)")
        if vk["layout_total"] != self.model.layout.total:
            raise ValueError("VitkAdamW.load_state_dict: optimizer state belongs to a different model layout")
        self._step = int(vk["step"])
        if vk["exp_avg"] is not None:
            if self._m is None:
                self._init_state()
            self._m.copy_(vk["exp_avg"])
            self._v.copy_(vk["exp_avg_sq"])
