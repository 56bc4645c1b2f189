"""Timing of the attention kernels at the model's shape (B=16, T=577, H=12) through the C ABI."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chest_x_ray_vit_b200 as pkg  # noqa: E402

ops = pkg.ops
_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
B, T, H = (int(x) for x in (_pos[:3] if len(_pos) >= 3 else (16, 577, 12)))
dev = "cuda"
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, T, 3, H, 64, generator=g).to(dev).to(torch.bfloat16)
do = (torch.randn(B * T, H * 64, generator=g) * 0.1).to(dev).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
o, lse = ops.attn_fwd(qkv, B, T, H, 0.125)
ws = torch.empty(ops.attn_bwd_workspace_bytes(B, T, H), dtype=torch.uint8, device=dev)
dqkv = torch.empty(B * T, 3 * H * 64, dtype=torch.bfloat16, device=dev)


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(2_000_000)      # host runs ahead of the GPU: the event pairs bracket device time only
    evs = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    ts.sort()
    return ts[len(ts) // 2]


f_fwd = 4.0 * B * H * T * T * 64
tf = timeit(lambda: ops.attn_fwd(qkv, B, T, H, 0.125, o=o, lse=lse))
tb = timeit(lambda: ops.attn_bwd(qkv, o, do, lse, B, T, H, 0.125, dqkv=dqkv, workspace=ws))
print(f"attn fwd {tf * 1e3:7.1f} us  {f_fwd / tf / 1e9:6.0f} TF/s   bwd (delta+main+dq) {tb * 1e3:7.1f} us  {2 * f_fwd / tb / 1e9:6.0f} TF/s")

if "--lib" in sys.argv:
    # Library baselines on the same shape (L2 flushed, same timer): torch SDPA's flash / cuDNN / efficient backends
    # (mma.sync kernels recompiled for sm_100, cuDNN's own Blackwell kernels) and the flash_attn package.
    import torch.nn.functional as Fn
    from torch.nn.attention import SDPBackend, sdpa_kernel
    q, k, v = (qkv[:, :, i].transpose(1, 2).contiguous().requires_grad_(True) for i in range(3))     # [B, H, T, 64]
    dob = do.view(B, T, H, 64).transpose(1, 2).contiguous()
    for name, be in (("flash", SDPBackend.FLASH_ATTENTION), ("cudnn", SDPBackend.CUDNN_ATTENTION), ("efficient", SDPBackend.EFFICIENT_ATTENTION)):
        try:
            with sdpa_kernel([be]):
                out = Fn.scaled_dot_product_attention(q, k, v, scale=0.125)
                tfl = timeit(lambda: Fn.scaled_dot_product_attention(q, k, v, scale=0.125))
                tbl = timeit(lambda: torch.autograd.grad(out, (q, k, v), dob, retain_graph=True))
            print(f"torch SDPA {name:9s} fwd {tfl * 1e3:7.1f} us  {f_fwd / tfl / 1e9:6.0f} TF/s   bwd {tbl * 1e3:7.1f} us  {2 * f_fwd / tbl / 1e9:6.0f} TF/s")
        except Exception as e:       # noqa: BLE001 — a backend that is not built for this device is a data point too
            print(f"torch SDPA {name:9s} unavailable: {str(e).splitlines()[0][:100]}")
    try:
        from flash_attn import flash_attn_func
        qf, kf, vf = (qkv[:, :, i].contiguous().requires_grad_(True) for i in range(3))               # [B, T, H, 64]
        dof = do.view(B, T, H, 64)
        out = flash_attn_func(qf, kf, vf, softmax_scale=0.125)
        tfl = timeit(lambda: flash_attn_func(qf, kf, vf, softmax_scale=0.125))
        tbl = timeit(lambda: torch.autograd.grad(out, (qf, kf, vf), dof, retain_graph=True))
        print(f"flash_attn package   fwd {tfl * 1e3:7.1f} us  {f_fwd / tfl / 1e9:6.0f} TF/s   bwd {tbl * 1e3:7.1f} us  {2 * f_fwd / tbl / 1e9:6.0f} TF/s")
    except Exception as e:           # noqa: BLE001
        print(f"flash_attn package unavailable: {str(e).splitlines()[0][:100]}")
