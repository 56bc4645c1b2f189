"""clock64 timeline of CTA (0,0,0) of the attention forward kernel (vitk_debug_timeline stamps)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chest_x_ray_vit_b200 as pkg  # noqa: E402

ops = pkg.ops
B, T, H = 16, 577, 12
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, T, 3, H, 64, generator=g).cuda().to(torch.bfloat16)
o, lse = ops.attn_fwd(qkv, B, T, H, 0.125)
tl = torch.zeros(8192, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
pkg._lib.lib().vitk_debug_timeline(tl.data_ptr())
ops.attn_fwd(qkv, B, T, H, 0.125, o=o, lse=lse)
torch.cuda.synchronize()
pkg._lib.lib().vitk_debug_timeline(None)
t = tl.cpu().tolist()
t0 = t[2000]
print(f"entry 0   prologue done {t[2001] - t0}   O ready {t[2002] - t0}   done {t[2003] - t0}")
names = ["ctrl: P(u) ready", "ctrl: PV(u)+S(u+1) issued", "ctrl: K/V buffer free", "ctrl: reload issued", "w0: S(u) ready", "w0: first half done",
         "w0: P(u) published"]
for u in range(10):
    print(f"--- key sub-block {u}")
    for k, n in enumerate(names):
        v = t[2016 + 8 * u + k]
        if v:
            print(f"   {n:28s} {v - t0:8d}")
