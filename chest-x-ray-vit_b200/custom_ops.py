"""torch custom ops (``torch.library.custom_op``) over the C ABI, with fake (meta) kernels and
autograd formulas, for composing the kernels outside the whole-model engine:

    vitk::layer_norm      aten::native_layer_norm      (HF modeling_vit.py:333,340,455)
    vitk::linear          aten::addmm (+ gelu / residual epilogues)  (HF :228-230,266,297-298,309-311)
    vitk::attention       aten::scaled_dot_product_attention          (HF :232-246)

``functional`` wrappers below give them PyTorch-style signatures; ``register_hf_attention()`` plugs the
attention kernel into HuggingFace's ``AttentionInterface`` (modeling_utils.py:4832-4870) so an unmodified
HF ViT can run with it (``config._attn_implementation = "vitk_b200"``).  All ops need CUDA tensors on an
sm_100 device; there is no CPU implementation (the fake kernels only propagate shapes).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import ops

bf16, f32 = torch.bfloat16, torch.float32


# ----------------------------------------------------------------------------- layer_norm
@torch.library.custom_op("vitk::layer_norm", mutates_args=())
def layer_norm_op(x: Tensor, gamma: Tensor, beta: Tensor, eps: float) -> Tuple[Tensor, Tensor, Tensor]:
    return ops.layernorm_fwd(x.contiguous(), gamma, beta, eps)


@layer_norm_op.register_fake
def _(x, gamma, beta, eps):
    M, D = x.shape
    return x.new_empty((M, D), dtype=bf16), x.new_empty((M,), dtype=f32), x.new_empty((M,), dtype=f32)


@torch.library.custom_op("vitk::layer_norm_bwd", mutates_args=())
def layer_norm_bwd_op(dy: Tensor, x: Tensor, mean: Tensor, rstd: Tensor, gamma: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    dg, db = torch.zeros_like(gamma), torch.zeros_like(gamma)
    dx = ops.layernorm_bwd(dy.contiguous(), x, mean, rstd, gamma, None, dg, db)
    return dx, dg, db


@layer_norm_bwd_op.register_fake
def _(dy, x, mean, rstd, gamma):
    return dy.new_empty(dy.shape, dtype=bf16), torch.empty_like(gamma), torch.empty_like(gamma)


def _ln_setup(ctx, inputs, output):
    x, gamma, _, _ = inputs
    _, mean, rstd = output
    ctx.save_for_backward(x, gamma, mean, rstd)


def _ln_backward(ctx, dy, _dmean, _drstd):
    x, gamma, mean, rstd = ctx.saved_tensors
    dx, dg, db = layer_norm_bwd_op(dy.to(bf16), x, mean, rstd, gamma)
    return dx.to(x.dtype), dg, db, None


layer_norm_op.register_autograd(_ln_backward, setup_context=_ln_setup)


# ----------------------------------------------------------------------------- linear
@torch.library.custom_op("vitk::linear", mutates_args=())
def linear_op(x: Tensor, weight: Tensor, bias: Optional[Tensor], gelu: bool) -> Tensor:
    """y = x·Wᵀ (+ bias) (→ exact-erf GELU).  x bf16 [M,K], weight bf16 [N,K], bias fp32 [N]; y bf16 [M,N]."""
    M, K = x.shape
    N = weight.shape[0]
    y = torch.empty((M, N), dtype=bf16, device=x.device)
    if gelu:
        ops.gemm(x, weight, M, N, K, y, epilogue=ops.EPI_BIAS_GELUG_BF16, bias=bias)
    else:
        ops.gemm(x, weight, M, N, K, y, epilogue=ops.EPI_BIAS_BF16 if bias is not None else ops.EPI_STORE_BF16, bias=bias)
    return y


@linear_op.register_fake
def _(x, weight, bias, gelu):
    return x.new_empty((x.shape[0], weight.shape[0]), dtype=bf16)


@torch.library.custom_op("vitk::linear_bwd", mutates_args=())
def linear_bwd_op(dy: Tensor, x: Tensor, weight: Tensor, need_bias: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """dx = dy·W (bf16), dW = dyᵀ·x (fp32, tcgen05 wgrad with both operands read MN-major), db = Σ_rows dy."""
    M, N = dy.shape
    K = x.shape[1]
    dx = torch.empty((M, K), dtype=bf16, device=dy.device)
    ops.gemm(dy, weight, M, K, N, dx, epilogue=ops.EPI_STORE_BF16, b_mn_major=True)
    dw = torch.zeros((N, K), dtype=f32, device=dy.device)
    ops.gemm(dy, x, N, K, M, dw, epilogue=ops.EPI_ACCUM_F32, a_mn_major=True, b_mn_major=True)
    db = torch.zeros((N,), dtype=f32, device=dy.device)
    if need_bias:
        ops.colsum(dy, db)
    return dx, dw, db


@linear_bwd_op.register_fake
def _(dy, x, weight, need_bias):
    return (dy.new_empty((dy.shape[0], x.shape[1]), dtype=bf16), dy.new_empty(weight.shape, dtype=f32),
            dy.new_empty((weight.shape[0],), dtype=f32))


def _lin_setup(ctx, inputs, output):
    x, weight, bias, gelu = inputs
    if gelu:
        raise RuntimeError("vitk::linear(gelu=True) is the inference form; use vitk::linear_gelu, which also returns gelu'(u) "
                           "for the backward pass, when gradients are needed")
    ctx.save_for_backward(x, weight)
    ctx.has_bias = bias is not None


def _lin_backward(ctx, dy):
    x, weight = ctx.saved_tensors
    dx, dw, db = linear_bwd_op(dy.to(bf16).contiguous(), x, weight, ctx.has_bias)
    return dx, dw.to(weight.dtype), (db if ctx.has_bias else None), None


linear_op.register_autograd(_lin_backward, setup_context=_lin_setup)


# fc1 of the MLP as one trainable op: a = gelu(x·Wᵀ + b) and g' = gelu'(x·Wᵀ + b) from ONE GEMM epilogue
# (VITK_EPI_BIAS_GELUG_BF16, HF modeling_vit.py:296-298 + activations.py:85-86); backward is the multiplier GEMM the engine
# uses for fc2's data gradient turned around: du = dy ∘ g' feeds vitk::linear_bwd.
@torch.library.custom_op("vitk::linear_gelu", mutates_args=())
def linear_gelu_op(x: Tensor, weight: Tensor, bias: Tensor) -> Tuple[Tensor, Tensor]:
    M, K = x.shape
    N = weight.shape[0]
    a = torch.empty((M, N), dtype=bf16, device=x.device)
    gp = torch.empty((M, N), dtype=bf16, device=x.device)
    ops.gemm(x, weight, M, N, K, a, epilogue=ops.EPI_BIAS_GELUG_BF16, bias=bias, d2=gp)
    return a, gp


@linear_gelu_op.register_fake
def _(x, weight, bias):
    y = x.new_empty((x.shape[0], weight.shape[0]), dtype=bf16)
    return y, torch.empty_like(y)


def _lg_setup(ctx, inputs, output):
    x, weight, bias = inputs
    ctx.save_for_backward(x, weight, output[1])


def _lg_backward(ctx, da, dgp):
    x, weight, gp = ctx.saved_tensors
    du = (da.to(bf16) * gp).contiguous()           # elementwise chain rule through the saved derivative
    dx, dw, db = linear_bwd_op(du, x, weight, True)
    return dx, dw.to(weight.dtype), db


linear_gelu_op.register_autograd(_lg_backward, setup_context=_lg_setup)


# ----------------------------------------------------------------------------- attention
@torch.library.custom_op("vitk::attention", mutates_args=())
def attention_op(qkv: Tensor, scale: float) -> Tuple[Tensor, Tensor]:
    """qkv bf16 [B,T,3,H,64] → (o bf16 [B,T,H·64], lse fp32 [B,H,T]); softmax(QKᵀ·scale)·V, no mask/dropout."""
    B, T, _, H, _ = qkv.shape
    o, lse = ops.attn_fwd(qkv.contiguous(), B, T, H, scale)
    return o.view(B, T, H * 64), lse


@attention_op.register_fake
def _(qkv, scale):
    B, T, _, H, dh = qkv.shape
    return qkv.new_empty((B, T, H * dh), dtype=bf16), qkv.new_empty((B, H, T), dtype=f32)


@torch.library.custom_op("vitk::attention_bwd", mutates_args=())
def attention_bwd_op(qkv: Tensor, o: Tensor, do: Tensor, lse: Tensor, scale: float) -> Tensor:
    B, T, _, H, _ = qkv.shape
    return ops.attn_bwd(qkv, o.reshape(B * T, H * 64), do.reshape(B * T, H * 64).contiguous(), lse, B, T, H, scale).view(qkv.shape)


@attention_bwd_op.register_fake
def _(qkv, o, do, lse, scale):
    return torch.empty_like(qkv)


def _attn_setup(ctx, inputs, output):
    qkv, scale = inputs
    o, lse = output
    ctx.save_for_backward(qkv, o, lse)
    ctx.scale = scale


def _attn_backward(ctx, do, _dlse):
    qkv, o, lse = ctx.saved_tensors
    return attention_bwd_op(qkv, o, do.to(bf16), lse, ctx.scale), None


attention_op.register_autograd(_attn_backward, setup_context=_attn_setup)


# ----------------------------------------------------------------------------- functional wrappers
class functional:
    @staticmethod
    def layer_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-12) -> Tensor:
        """fp32 [.., D] → bf16 [.., D] (fp32 statistics)."""
        shp = x.shape
        y, _, _ = layer_norm_op(x.reshape(-1, shp[-1]), weight, bias, eps)
        return y.view(shp)

    @staticmethod
    def linear(x: Tensor, weight_bf16: Tensor, bias: Optional[Tensor] = None, gelu: bool = False) -> Tensor:
        """gelu=True fuses the exact-erf GELU into the GEMM epilogue; with gradients enabled it goes through
        vitk::linear_gelu (same epilogue, which also emits gelu'(u) for the backward pass)."""
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).to(bf16).contiguous()
        if gelu and bias is not None and torch.is_grad_enabled() and (x.requires_grad or weight_bf16.requires_grad or bias.requires_grad):
            y = linear_gelu_op(x2, weight_bf16, bias)[0]
        else:
            y = linear_op(x2, weight_bf16, bias, gelu)
        return y.view(*shp[:-1], weight_bf16.shape[0])

    @staticmethod
    def attention(qkv: Tensor, scale: Optional[float] = None) -> Tensor:
        return attention_op(qkv, float(scale if scale is not None else qkv.shape[-1] ** -0.5))[0]


# ----------------------------------------------------------------------------- HF plug-in point
def hf_attention_forward(module, query: Tensor, key: Tensor, value: Tensor, attention_mask=None, scaling: Optional[float] = None,
                         dropout: float = 0.0, **kwargs):
    """Signature of HF's attention functions (sdpa_attention.py:40-104): q,k,v [B,H,T,dh] → ([B,T,H,dh], None)."""
    if attention_mask is not None or dropout != 0.0 or query.shape[-1] != 64:
        raise ValueError("vitk_b200 attention: no mask, no dropout, head_dim 64 only")
    B, H, T, dh = query.shape
    qkv = torch.stack((query, key, value), dim=2).permute(0, 3, 2, 1, 4).to(bf16).contiguous()    # [B,T,3,H,dh]
    o = functional.attention(qkv, scaling)
    return o.view(B, T, H, dh).to(query.dtype), None


def register_hf_attention(name: str = "vitk_b200") -> str:
    from transformers.modeling_utils import AttentionInterface
    AttentionInterface.register(name, hf_attention_forward)
    return name
