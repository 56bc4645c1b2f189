"""tcgen05 GEMM through the C ABI vs an fp32 torch.matmul on the same bf16-rounded operands."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"
bf16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    import chest_x_ray_vit_b200 as pkg
    pkg.ops.check_device(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    return pkg.ops


def _operands(M, N, K, a_mn, b_mn, seed):
    g = torch.Generator().manual_seed(seed)
    a = (torch.randn(M, K, generator=g)).to(dev).to(bf16)
    b = (torch.randn(N, K, generator=g) * 0.05).to(dev).to(bf16)
    a_store = a.t().contiguous() if a_mn else a
    b_store = b.t().contiguous() if b_mn else b
    ref = a.float() @ b.float().t()
    return a_store, b_store, ref


def _tol(K, ref):
    return 3e-3 * ref.abs().max().item() + 1e-3 * math.sqrt(K) * 0.05


SHAPES = [
    (128, 128, 64), (128, 256, 128), (256, 768, 768), (1154, 768, 768), (1154, 2304, 768),
    (1154, 3072, 768), (1154, 768, 3072), (100, 384, 72), (577, 768, 200), (9232, 768, 768),
]


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
def test_store_f32_all_layouts(ops, M, N, K, a_mn, b_mn, variant):
    if a_mn and M % 8:
        pytest.skip("MN-major A needs lda multiple of 8")
    if (not a_mn or not b_mn) and K % 8:
        pytest.skip("K-major operand needs K multiple of 8")
    a, b, ref = _operands(M, N, K, a_mn, b_mn, M + N + K)
    d = torch.full((M, N), float("nan"), device=dev)
    ops.gemm(a, b, M, N, K, d, epilogue=ops.EPI_STORE_F32, a_mn_major=a_mn, b_mn_major=b_mn, variant=variant)
    torch.cuda.synchronize()
    err = (d - ref).abs().max().item()
    assert err <= _tol(K, ref), f"max err {err}"


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("tile_n", [128, 192, 256])
@pytest.mark.parametrize("M,N,K", [(1154, 768, 768), (300, 2304, 192), (9232, 3072, 768), (16, 768, 64)])
def test_tile_shapes(ops, M, N, K, tile_n, b_mn, variant):
    if N % tile_n:
        pytest.skip("N not a multiple of the tile")
    if variant == 2 and tile_n == 192 and b_mn:
        pytest.skip("the CTA-pair kernel needs whole 64-column boxes for an MN-major B half")
    a, b, ref = _operands(M, N, K, False, b_mn, 5)
    d = torch.full((M, N), float("nan"), device=dev, dtype=bf16)
    ops.gemm(a, b, M, N, K, d, epilogue=ops.EPI_STORE_BF16, tile_n=tile_n, b_mn_major=b_mn, variant=variant)
    err = (d.float() - ref).abs().max().item()
    assert err <= _tol(K, ref) + 2 ** -8 * ref.abs().max().item(), f"max err {err}"


@pytest.mark.parametrize("variant", [1, 2])
def test_bias_and_gelu_epilogues(ops, variant):
    M, N, K = 1154, 3072, 768
    a, b, ref = _operands(M, N, K, False, False, 11)
    bias = torch.randn(N, device=dev) * 0.5
    u = torch.empty((M, N), device=dev, dtype=bf16)
    act = torch.empty((M, N), device=dev, dtype=bf16)
    ops.gemm(a, b, M, N, K, u, epilogue=ops.EPI_BIAS_GELU_BF16, d2=act, bias=bias, variant=variant)
    ur = ref + bias
    assert (u.float() - ur).abs().max() <= _tol(K, ur) + 2 ** -8 * ur.abs().max()
    gr = torch.nn.functional.gelu(ur)
    assert (act.float() - gr).abs().max() <= _tol(K, ur) + 2 ** -8 * gr.abs().max()
    d = torch.empty((M, N), device=dev, dtype=bf16)
    ops.gemm(a, b, M, N, K, d, epilogue=ops.EPI_BIAS_BF16, bias=bias, variant=variant)
    assert torch.equal(d, u)
    # training form: gelu(u) and gelu'(u), then the backward multiplier epilogue
    act2 = torch.empty((M, N), device=dev, dtype=bf16)
    gp = torch.empty((M, N), device=dev, dtype=bf16)
    ops.gemm(a, b, M, N, K, act2, epilogue=ops.EPI_BIAS_GELUG_BF16, d2=gp, bias=bias, variant=variant)
    assert (act2.float() - gr).abs().max() <= _tol(K, ur) + 2 ** -8 * gr.abs().max()
    uf = ur.clone().requires_grad_(True)
    torch.nn.functional.gelu(uf).sum().backward()
    assert (gp.float() - uf.grad).abs().max() <= _tol(K, ur) + 2 ** -8 * 1.2
    act3 = torch.full((M, N), float("nan"), device=dev, dtype=bf16)
    ops.gemm(a, b, M, N, K, act3, epilogue=ops.EPI_BIAS_GELUG_BF16, bias=bias, variant=variant)     # inference: no d2
    assert torch.equal(act3, act2)
    mul = torch.empty((M, N), device=dev, dtype=bf16)
    ops.gemm(a, b, M, N, K, mul, epilogue=ops.EPI_MUL_BF16, aux=gp, variant=variant)
    rm = ref * gp.float()
    assert (mul.float() - rm).abs().max() <= _tol(K, rm) + 2 ** -8 * rm.abs().max()


@pytest.mark.parametrize("variant", [1, 2])
def test_gelu_matches_erf_form_pointwise(ops, variant):
    """K=8 identity-like GEMM isolates the epilogue: gelu in the epilogue vs torch erf GELU."""
    M, N, K = 256, 128, 64
    x = torch.linspace(-8, 8, M * N, device=dev).view(M, N)
    a = torch.zeros(M, K, device=dev, dtype=bf16)
    b = torch.zeros(N, K, device=dev, dtype=bf16)
    u = torch.empty((M, N), device=dev, dtype=bf16)
    act = torch.empty((M, N), device=dev, dtype=bf16)
    # acc = 0, so u = bias-free aux path is not available; use the residual epilogue to inject x
    pre = x.to(bf16)
    dgrad = torch.empty((M, N), device=dev, dtype=bf16)
    a[:, 0] = 1.0
    b[:, 0] = 1.0   # acc = 1 everywhere
    ops.gemm(a, b, M, N, K, dgrad, epilogue=ops.EPI_DGELU_BF16, aux=pre, variant=variant)
    xf = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(xf).sum().backward()
    assert (dgrad.float() - xf.grad).abs().max() <= 2 ** -8 * 1.2 + 1e-6


@pytest.mark.parametrize("variant", [1, 2])
def test_bias_resid_and_patch_epilogues(ops, variant):
    M, N, K = 1154, 768, 768
    a, b, ref = _operands(M, N, K, False, False, 13)
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev)
    d = torch.empty((M, N), device=dev)
    ops.gemm(a, b, M, N, K, d, epilogue=ops.EPI_BIAS_RESID_F32, bias=bias, aux=res, variant=variant)
    rr = ref + bias + res
    assert (d - rr).abs().max() <= _tol(K, rr)
    # in-place residual (d == aux) is how the engine updates the residual stream
    d2 = res.clone()
    ops.gemm(a, b, M, N, K, d2, epilogue=ops.EPI_BIAS_RESID_F32, bias=bias, aux=d2, variant=variant)
    assert (d2 - rr).abs().max() <= _tol(K, rr)
    # patch embedding row remap: B images × P patches → rows 1+p of a [B, P+1, N] buffer
    Bn, P = 2, 576
    M = Bn * P
    a, b, ref = _operands(M, N, K, False, False, 14)
    pos = torch.randn(P + 1, N, device=dev)
    out = torch.full((Bn, P + 1, N), 7.0, device=dev)
    ops.gemm(a, b, M, N, K, out, epilogue=ops.EPI_PATCH_F32, bias=bias, aux=pos, ldd=N, ld_aux=N, rows_in=P,
             rows_out=P + 1, row_off=1)
    rr = (ref + bias).view(Bn, P, N) + pos[1:]
    assert (out[:, 1:] - rr).abs().max() <= _tol(K, rr)
    assert (out[:, 0] == 7.0).all()


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("split_k", [0, 1, 3, 8])
def test_wgrad_accumulate_split_k(ops, split_k, variant):
    """wgrad shape: small M×N, long ragged K, both operands MN-major, accumulate into fp32."""
    M, N, K = 768, 768, 9232
    g = torch.Generator().manual_seed(21)
    dy = (torch.randn(K, M, generator=g) * 0.1).to(dev).to(bf16)    # [tokens, out]  → A stored [K, M]
    x = torch.randn(K, N, generator=g).to(dev).to(bf16)             # [tokens, in]   → B stored [K, N]
    ref = dy.float().t() @ x.float()
    d = torch.ones((M, N), device=dev)
    ops.gemm(dy, x, M, N, K, d, epilogue=ops.EPI_ACCUM_F32, a_mn_major=True, b_mn_major=True, split_k=split_k, variant=variant)
    err = (d - 1 - ref).abs().max().item()
    assert err <= _tol(K, ref) * 4, f"max err {err}"


def test_argument_errors(ops):
    a = torch.zeros(128, 64, device=dev, dtype=bf16)
    d = torch.zeros(128, 100, device=dev)
    with pytest.raises(RuntimeError, match="multiple of 128"):
        ops.gemm(a, a, 128, 100, 64, d, epilogue=ops.EPI_STORE_F32)
    with pytest.raises(RuntimeError, match="split_k"):
        ops.gemm(a, a, 128, 128, 64, torch.zeros(128, 128, device=dev), epilogue=ops.EPI_STORE_F32, split_k=2)


@pytest.mark.parametrize("b_mn", [False, True])
@pytest.mark.parametrize("N,K", [(768, 192), (768, 768), (1280, 128)])
def test_mixed_width_tiles_full_batch(ops, N, K, b_mn):
    """M = 9232 (batch 16): 37 bands of 256 rows on 74 CTA pairs.  With N = 768 the pair kernel cuts every band into
    2 tiles of 256 columns + 2 of 128 (one of each per pair) instead of 1.5 waves of 256-wide tiles; every epilogue
    family has to cope with tiles of both widths in one launch (bias slices, aux look-ahead across a width change)."""
    M = 9232
    a, b, ref = _operands(M, N, K, False, b_mn, 21)
    d = torch.full((M, N), float("nan"), device=dev, dtype=bf16)
    ops.gemm(a, b, M, N, K, d, epilogue=ops.EPI_STORE_BF16, b_mn_major=b_mn)
    assert (d.float() - ref).abs().max() <= _tol(K, ref) + 2 ** -8 * ref.abs().max()
    pure = torch.empty_like(d)
    ops.gemm(a, b, M, N, K, pure, epilogue=ops.EPI_STORE_BF16, b_mn_major=b_mn, tile_n=256 if N % 256 == 0 else 128)
    assert torch.equal(d, pure)                      # same K order per element: the tiling must not change a bit
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev)
    out = res.clone()
    ops.gemm(a, b, M, N, K, out, epilogue=ops.EPI_BIAS_RESID_F32, bias=bias, aux=out, b_mn_major=b_mn)
    rr = ref + bias + res
    assert (out - rr).abs().max() <= _tol(K, rr)
    mult = (torch.randn(M, N, device=dev) * 0.5).to(bf16)
    mul = torch.empty((M, N), device=dev, dtype=bf16)
    ops.gemm(a, b, M, N, K, mul, epilogue=ops.EPI_MUL_BF16, aux=mult, b_mn_major=b_mn)
    rm = ref * mult.float()
    assert (mul.float() - rm).abs().max() <= _tol(K, rm) + 2 ** -8 * rm.abs().max()
    act = torch.empty((M, N), device=dev, dtype=bf16)
    gp = torch.empty((M, N), device=dev, dtype=bf16)
    ops.gemm(a, b, M, N, K, act, epilogue=ops.EPI_BIAS_GELUG_BF16, d2=gp, bias=bias, b_mn_major=b_mn)
    ur = ref + bias
    gr = torch.nn.functional.gelu(ur)
    assert (act.float() - gr).abs().max() <= _tol(K, ur) + 2 ** -8 * gr.abs().max()


@pytest.mark.parametrize("M", [9232, 4616, 1154, 300])
def test_sixteen_warp_epilogues_equal_eight_warp(ops, M):
    """fc1 + GELU (+ GELU') runs as a 16-epilogue-warp instantiation of the pair kernel; variant 3 forces the 8-warp
    kernel on the same tiling.  Same K order, same arithmetic per element: the outputs must agree bit for bit, with and
    without the second output, ragged last band included.  The multiplier / fp32-residual epilogues (8 warps either way)
    ride along as a check that variant 3 changes nothing else."""
    g = torch.Generator().manual_seed(M)
    D, F = 768, 3072
    x = torch.randn(M, D, generator=g).to(dev).to(bf16)
    w1 = (torch.randn(F, D, generator=g) * 0.05).to(dev).to(bf16)
    b1 = torch.randn(F, generator=g).to(dev) * 0.5
    outs = {}
    for variant in (2, 3):
        act = torch.full((M, F), float("nan"), device=dev, dtype=bf16)
        gp = torch.full((M, F), float("nan"), device=dev, dtype=bf16)
        ops.gemm(x, w1, M, F, D, act, epilogue=ops.EPI_BIAS_GELUG_BF16, d2=gp, bias=b1, variant=variant, tile_n=256)
        act1 = torch.full((M, F), float("nan"), device=dev, dtype=bf16)
        ops.gemm(x, w1, M, F, D, act1, epilogue=ops.EPI_BIAS_GELUG_BF16, bias=b1, variant=variant, tile_n=256)   # no d2
        # fc2 data gradient: dh [M, D] · W2 [D, F] (MN-major B) × gelu'
        du = torch.full((M, F), float("nan"), device=dev, dtype=bf16)
        w2 = w1.t().contiguous()                                     # stored [D, F]: element (n, k) at k·F + n
        ops.gemm(x, w2, M, F, D, du, epilogue=ops.EPI_MUL_BF16, aux=gp, b_mn_major=True, variant=variant, tile_n=256)
        # out-proj / fc2 forward with the fp32 residual, in place, default tiling (192-wide or mixed) and forced 256
        wo = w1[:D].contiguous()
        bo = b1[:D].contiguous()
        res = torch.randn(M, D, generator=torch.Generator().manual_seed(1)).to(dev)
        h = {}
        for tn in (0, 192, 256):
            h[tn] = res.clone()
            ops.gemm(x, wo, M, D, D, h[tn], epilogue=ops.EPI_BIAS_RESID_F32, bias=bo, aux=h[tn], variant=variant, tile_n=tn)
        h2 = res.clone()
        ops.gemm(act, w2, M, D, F, h2, epilogue=ops.EPI_BIAS_RESID_F32, bias=bo, aux=h2, variant=variant)      # K = 3072
        outs[variant] = (act, gp, act1, du, h[0], h[192], h[256], h2)
    torch.cuda.synchronize()
    names = ("gelu", "gelu'", "gelu (no d2)", "mul", "resid auto", "resid 192", "resid 256", "resid K=3072")
    for n, a16, a8 in zip(names, outs[2], outs[3]):
        assert not torch.isnan(a16.float()).any(), n
        assert torch.equal(a16, a8), n
    ref = torch.nn.functional.gelu(x.float() @ w1.float().t() + b1)
    assert (outs[2][0].float() - ref).abs().max() <= _tol(D, ref) + 2 ** -8 * ref.abs().max()
