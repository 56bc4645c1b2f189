"""Dump the per-phase clock64 timeline of CTA (0,0,0) of the attention backward kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chest_x_ray_vit_b200 as pkg  # noqa: E402

ops = pkg.ops
B, T, H = 16, 577, 12
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, T, 3, H, 64, generator=g).cuda().to(torch.bfloat16)
do = (torch.randn(B * T, H * 64, generator=g) * 0.1).cuda().to(torch.bfloat16)
o, lse = ops.attn_fwd(qkv, B, T, H, 0.125)
ops.attn_bwd(qkv, o, do, lse, B, T, H, 0.125)
tl = torch.zeros(1024, dtype=torch.int64, device="cuda")
pkg._lib.lib().vitk_debug_timeline(tl.data_ptr())
ops.attn_bwd(qkv, o, do, lse, B, T, H, 0.125)
torch.cuda.synchronize()
pkg._lib.lib().vitk_debug_timeline(None)
t = tl.cpu().tolist()
t0 = min(x for x in t if x > 0)
names = {0: "ctrl: bar_pd done", 1: "ctrl: dV/dK/dQ issued", 2: "ctrl: scores(i+1) issued", 3: "ctrl: bar_g done (reload)",
         8: "w0: start iter", 9: "w0: bar_s done", 10: "w0: tmem loaded", 11: "w0: math done", 12: "w0: stores+fences done",
         13: "w0: reduce_dq(i-1) done"}
for i in range(10):
    print(f"--- query sub-block {i}")
    for k in sorted(names):
        v = t[16 * i + k]
        if v:
            print(f"   {names[k]:32s} {v - t0:8d}")
