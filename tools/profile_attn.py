"""Attention forward + backward at the model's shape (B=16, T=577, H=12), twice (first pass warms up): for
ncu -k regex:attn_ -s <launches of pass 1> captures."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chest_x_ray_vit_b200 as pkg
ops = pkg.ops
B, T, H = 16, 577, 12
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, T, 3, H, 64, generator=g).cuda().to(torch.bfloat16)
do = (torch.randn(B * T, H * 64, generator=g) * 0.1).cuda().to(torch.bfloat16)
ws = torch.empty(ops.attn_bwd_workspace_bytes(B, T, H), dtype=torch.uint8, device="cuda")
dqkv = torch.empty(B * T, 3 * H * 64, dtype=torch.bfloat16, device="cuda")
for _ in range(2):
    o, lse = ops.attn_fwd(qkv, B, T, H, 0.125)
    ops.attn_bwd(qkv, o, do, lse, B, T, H, 0.125, dqkv=dqkv, workspace=ws)
    torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
