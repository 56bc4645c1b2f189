"""torch custom ops over the C ABI: opcheck-style parity with autograd, and the HF AttentionInterface plug-in."""
import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"
bf16 = torch.bfloat16


@pytest.fixture(scope="module")
def co():
    import chest_x_ray_vit_b200 as pkg
    pkg.ops.check_device(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    return pkg.custom_ops


def test_layer_norm_op_autograd(co):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 37, 768, generator=g).to(dev).requires_grad_(True)
    w = (1 + 0.05 * torch.randn(768, generator=g)).to(dev).requires_grad_(True)
    b = (0.02 * torch.randn(768, generator=g)).to(dev).requires_grad_(True)
    dy = torch.randn(4, 37, 768, generator=g).to(dev)
    y = co.functional.layer_norm(x, w, b, 1e-12)
    y.backward(dy.to(bf16))
    xr, wr, br = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    yr = torch.nn.functional.layer_norm(xr, (768,), wr, br, 1e-12)
    yr.backward(dy.to(bf16).float())
    assert (y.float() - yr).abs().max() < 0.03
    assert (x.grad - xr.grad).abs().max() <= 2 ** -7 * xr.grad.abs().max() + 1e-5
    assert torch.allclose(w.grad, wr.grad, rtol=1e-3, atol=1e-3 * wr.grad.abs().max().item())
    assert torch.allclose(b.grad, br.grad, rtol=1e-3, atol=1e-3 * br.grad.abs().max().item())


def test_linear_op_autograd(co):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 197, 768, generator=g).to(dev).to(bf16).requires_grad_(True)
    w = (0.03 * torch.randn(2304, 768, generator=g)).to(dev).to(bf16).requires_grad_(True)
    b = (0.1 * torch.randn(2304, generator=g)).to(dev).requires_grad_(True)
    dy = (0.1 * torch.randn(3, 197, 2304, generator=g)).to(dev).to(bf16)
    y = co.functional.linear(x, w, b)
    y.backward(dy)
    xr, wr, br = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = torch.nn.functional.linear(xr, wr, br)
    yr.backward(dy.float())
    assert (y.float() - yr).abs().max() <= 0.02 * yr.abs().max()
    for got, ref in ((x.grad, xr.grad), (w.grad, wr.grad), (b.grad, br.grad)):
        cos = torch.nn.functional.cosine_similarity(got.float().flatten(), ref.flatten(), dim=0).item()
        assert cos > 0.9999, cos
    # GELU epilogue: inference form ...
    yg = co.functional.linear(x.detach(), w.detach(), b.detach(), gelu=True)
    assert (yg.float() - torch.nn.functional.gelu(yr.detach())).abs().max() <= 0.03 * yr.abs().max()
    # ... and trainable (vitk::linear_gelu: gelu and gelu' out of one epilogue, chain rule in backward)
    x2, w2, b2 = x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    ya = co.functional.linear(x2, w2, b2, gelu=True)
    ya.backward(dy)
    xr2, wr2, br2 = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr2 = torch.nn.functional.gelu(torch.nn.functional.linear(xr2, wr2, br2))
    yr2.backward(dy.float())
    assert (ya.float() - yr2).abs().max() <= 0.03 * yr2.abs().max()
    for got, ref in ((x2.grad, xr2.grad), (w2.grad, wr2.grad), (b2.grad, br2.grad)):
        cos = torch.nn.functional.cosine_similarity(got.float().flatten(), ref.flatten(), dim=0).item()
        assert cos > 0.999, cos


def test_fake_kernels_propagate_shapes(co):
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        qkv = torch.empty(2, 577, 3, 12, 64, dtype=bf16, device=dev)
        o, lse = co.attention_op(qkv, 0.125)
        assert o.shape == (2, 577, 768) and lse.shape == (2, 12, 577) and o.dtype == bf16
        y = co.linear_op(torch.empty(10, 768, dtype=bf16, device=dev), torch.empty(3072, 768, dtype=bf16, device=dev), None, False)
        assert y.shape == (10, 3072)


def test_hf_vit_with_vitk_attention(co):
    tr = pytest.importorskip("transformers")
    name = co.register_hf_attention()
    torch.manual_seed(0)
    cfg = tr.ViTConfig(image_size=224, num_labels=14, problem_type="multi_label_classification")
    m = tr.ViTForImageClassification(cfg).to(dev).train()
    x = torch.randn(2, 3, 224, 224, device=dev)
    y = (torch.rand(2, 14, device=dev) < 0.2).float()
    m.config._attn_implementation = "sdpa"
    ref = m(pixel_values=x, labels=y)
    ref.loss.backward()
    gref = m.vit.encoder.layer[0].attention.attention.query.weight.grad.clone()
    m.zero_grad()
    m.config._attn_implementation = name
    out = m(pixel_values=x, labels=y)
    out.loss.backward()
    g = m.vit.encoder.layer[0].attention.attention.query.weight.grad
    assert (out.logits - ref.logits).abs().max() < 2e-2
    assert abs(out.loss.item() - ref.loss.item()) < 1e-3 * abs(ref.loss.item())
    assert torch.nn.functional.cosine_similarity(g.flatten(), gref.flatten(), dim=0).item() > 0.999
