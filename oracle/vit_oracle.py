"""CPU oracle for the ViT-B/16 fine-tuning hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain fp32 PyTorch on the CPU, the arithmetic the reference
script executes when it calls ``ViTForImageClassification(pixel_values, labels)``,
``loss.backward()`` and ``AdamW.step()``:

  reference call sites .... /root/reference/ViT-Training.py:83-90 (model ctor),
                            :120-132 (Trainer → model(**inputs), backward, step)
  arithmetic lives in ..... HuggingFace ``transformers`` (third party, NOT vendored
                            by the reference, unpinned in its requirements.txt:3;
                            this image has 5.5.0) on torch 2.11.0 ATen CPU kernels.

Every function cites the ``transformers/models/vit/modeling_vit.py`` (``HF:``) lines
it follows.  The oracle is pinned in two ways (tests/test_oracle.py):
  * against golden vectors frozen from HF 5.5.0 + torch 2.11.0 in this image
    (tests/golden/*.pt, produced by oracle/make_golden.py), and
  * live against ``transformers.ViTForImageClassification`` when importable.
The reference itself ships no tests or golden vectors, so upstream parity is
"unpinned"; the frozen vectors above are the pin this repo provides.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (chest-x-ray-vit_b200/) never does.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F


@dataclass
class OracleConfig:
    """Subset of HF ViTConfig (HF configuration_vit.py:50-65) the path depends on."""
    image_size: int = 384
    patch_size: int = 16
    num_channels: int = 3
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    layer_norm_eps: float = 1e-12
    num_labels: int = 14
    initializer_range: float = 0.02

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def seq_len(self) -> int:
        return self.num_patches + 1


def param_shapes(cfg: OracleConfig) -> Dict[str, Tuple[int, ...]]:
    """HF state-dict names and shapes, in ``named_parameters()`` order."""
    D, Fi, P, C = cfg.hidden_size, cfg.intermediate_size, cfg.patch_size, cfg.num_channels
    s: Dict[str, Tuple[int, ...]] = {}
    s["vit.embeddings.cls_token"] = (1, 1, D)
    s["vit.embeddings.position_embeddings"] = (1, cfg.seq_len, D)
    s["vit.embeddings.patch_embeddings.projection.weight"] = (D, C, P, P)
    s["vit.embeddings.patch_embeddings.projection.bias"] = (D,)
    for i in range(cfg.num_hidden_layers):
        p = f"vit.encoder.layer.{i}."
        for n in ("query", "key", "value"):
            s[p + f"attention.attention.{n}.weight"] = (D, D)
            s[p + f"attention.attention.{n}.bias"] = (D,)
        s[p + "attention.output.dense.weight"] = (D, D)
        s[p + "attention.output.dense.bias"] = (D,)
        s[p + "intermediate.dense.weight"] = (Fi, D)
        s[p + "intermediate.dense.bias"] = (Fi,)
        s[p + "output.dense.weight"] = (D, Fi)
        s[p + "output.dense.bias"] = (D,)
        s[p + "layernorm_before.weight"] = (D,)
        s[p + "layernorm_before.bias"] = (D,)
        s[p + "layernorm_after.weight"] = (D,)
        s[p + "layernorm_after.bias"] = (D,)
    s["vit.layernorm.weight"] = (D,)
    s["vit.layernorm.bias"] = (D,)
    s["classifier.weight"] = (cfg.num_labels, D)
    s["classifier.bias"] = (cfg.num_labels,)
    return s


def init_params(cfg: OracleConfig, seed: int = 0, perturb_seed: int = 123) -> Dict[str, torch.Tensor]:
    """Random init following HF modeling_vit.py:385-398 (trunc_normal std 0.02, zero
    bias, unit LN), then the SURVEY Appendix-C perturbation so every bias / LN affine
    parameter matters.  Not bit-identical to HF's RNG consumption order; tests that
    compare against HF copy HF's state dict instead."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        if "layernorm" in name and name.endswith("weight"):
            t = torch.ones(shape)
        elif name.endswith("bias"):
            t = torch.zeros(shape)
        else:
            t = torch.empty(shape)
            torch.nn.init.trunc_normal_(t, mean=0.0, std=cfg.initializer_range, generator=g)
        out[name] = t
    perturb_params_(out, perturb_seed)
    return out


def perturb_params_(params: Dict[str, torch.Tensor], seed: int = 123) -> torch.Generator:
    """bias = 0.02·randn, LN weight = 1 + 0.05·randn, in
    named_parameters() order.  Returns the generator so inputs continue the stream."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, t in params.items():
            if name.endswith("bias"):
                t.copy_(0.02 * torch.randn(t.shape, generator=g))
            elif "layernorm" in name and name.endswith("weight"):
                t.copy_(1.0 + 0.05 * torch.randn(t.shape, generator=g))
    return g


def synth_inputs(cfg: OracleConfig, batch: int, g: torch.Generator):
    """uint8 grayscale images and Bernoulli(0.1) labels."""
    x8 = torch.randint(0, 256, (batch, 1, cfg.image_size, cfg.image_size), dtype=torch.uint8, generator=g)
    y = (torch.rand(batch, cfg.num_labels, generator=g) < 0.1).float()
    return x8, y


def normalize_gray(x8: torch.Tensor, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)) -> torch.Tensor:
    """ToTensor + Normalize on ``img.convert("RGB")`` of a grayscale image
    (/root/reference/ViT-Training.py:60-66): x_c = (g/255 − mean_c)/std_c, c=0..2."""
    g = x8.to(torch.float32) / 255.0
    if g.dim() == 3:
        g = g.unsqueeze(1)
    m = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    return ((g.expand(-1, 3, -1, -1) - m) / s).contiguous()


# --------------------------------------------------------------------------- ops

def im2col(pixel_values: torch.Tensor, patch: int) -> torch.Tensor:
    """[B,C,H,W] → [B, P, C·p·p] with column k = c·p² + ky·p + kx and patch index
    py·(W/p)+px — the order of ``projection.weight.view(D, -1)`` and of
    ``.flatten(2).transpose(1,2)`` (HF:166)."""
    B, C, H, W = pixel_values.shape
    x = pixel_values.view(B, C, H // patch, patch, W // patch, patch)
    x = x.permute(0, 2, 4, 1, 3, 5).reshape(B, (H // patch) * (W // patch), C * patch * patch)
    return x


def embeddings(p, cfg: OracleConfig, pixel_values: torch.Tensor) -> torch.Tensor:
    """HF:100-128 and :153-167 — Conv2d(k=s=16) as a GEMM, CLS concat, +pos."""
    w = p["vit.embeddings.patch_embeddings.projection.weight"].reshape(cfg.hidden_size, -1)
    b = p["vit.embeddings.patch_embeddings.projection.bias"]
    a = im2col(pixel_values, cfg.patch_size)
    e = a @ w.t() + b
    cls = p["vit.embeddings.cls_token"].expand(pixel_values.shape[0], -1, -1)
    return torch.cat((cls, e), dim=1) + p["vit.embeddings.position_embeddings"]


def layer_norm(x, w, b, eps):
    """nn.LayerNorm (HF:325-326): biased variance over the last dim."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w + b


def gelu_erf(x):
    """HF activations.py:85-86 ("gelu" = exact erf form)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def attention(q, k, v, scale):
    """eager_attention_forward (HF:171-196), mathematically what SDPA computes
    (HF integrations/sdpa_attention.py:92-102): softmax(QKᵀ·scale)·V, no mask."""
    s = (q @ k.transpose(-1, -2)) * scale
    pr = torch.softmax(s, dim=-1)
    return pr @ v


def encoder_layer(p, cfg: OracleConfig, i: int, h: torch.Tensor) -> torch.Tensor:
    """ViTLayer.forward (HF:328-346)."""
    pre = f"vit.encoder.layer.{i}."
    B, T, D = h.shape
    H = cfg.num_attention_heads
    dh = D // H
    n1 = layer_norm(h, p[pre + "layernorm_before.weight"], p[pre + "layernorm_before.bias"], cfg.layer_norm_eps)
    def lin(x, name):
        return x @ p[pre + name + ".weight"].t() + p[pre + name + ".bias"]
    q = lin(n1, "attention.attention.query").view(B, T, H, dh).transpose(1, 2)
    k = lin(n1, "attention.attention.key").view(B, T, H, dh).transpose(1, 2)
    v = lin(n1, "attention.attention.value").view(B, T, H, dh).transpose(1, 2)
    o = attention(q, k, v, dh ** -0.5).transpose(1, 2).reshape(B, T, D)
    h1 = lin(o, "attention.output.dense") + h
    n2 = layer_norm(h1, p[pre + "layernorm_after.weight"], p[pre + "layernorm_after.bias"], cfg.layer_norm_eps)
    a = gelu_erf(lin(n2, "intermediate.dense"))
    return lin(a, "output.dense") + h1


def bce_with_logits_mean(logits, labels):
    """BCEWithLogitsLoss() mean over B·C (HF loss_utils.py:110-112)."""
    return (logits.clamp_min(0) - logits * labels + torch.log1p(torch.exp(-logits.abs()))).mean()


def forward(p: Dict[str, torch.Tensor], cfg: OracleConfig, pixel_values: torch.Tensor, labels=None):
    """ViTForImageClassification.forward (HF:620-653) → (loss | None, logits)."""
    h = embeddings(p, cfg, pixel_values)
    for i in range(cfg.num_hidden_layers):
        h = encoder_layer(p, cfg, i, h)
    z = layer_norm(h, p["vit.layernorm.weight"], p["vit.layernorm.bias"], cfg.layer_norm_eps)[:, 0, :]
    logits = z @ p["classifier.weight"].t() + p["classifier.bias"]
    loss = bce_with_logits_mean(logits, labels) if labels is not None else None
    return loss, logits


def forward_backward(p: Dict[str, torch.Tensor], cfg: OracleConfig, pixel_values, labels):
    """One fwd+bwd in fp32; returns (loss, logits, {name: grad})."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    loss, logits = forward(leaves, cfg, pixel_values, labels)
    loss.backward()
    grads = {k: v.grad for k, v in leaves.items()}
    return loss.detach(), logits.detach(), grads


def adamw_step(p, grads, state, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """torch.optim.AdamW semantics; ``state`` = {name: (step, m, v)}."""
    b1, b2 = betas
    for k in p:
        g = grads[k]
        t, m, v = state.get(k, (0, torch.zeros_like(g), torch.zeros_like(g)))
        t += 1
        p[k].mul_(1 - lr * weight_decay)
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        denom = v.sqrt() / math.sqrt(1 - b2 ** t) + eps
        p[k].addcdiv_(m, denom, value=-lr / (1 - b1 ** t))
        state[k] = (t, m, v)
    return p


TINY = OracleConfig(image_size=64, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                    intermediate_size=256, num_labels=14)
VIT_B16_384 = OracleConfig()
VIT_B16_224 = OracleConfig(image_size=224)
VIT_L16_384 = OracleConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)
