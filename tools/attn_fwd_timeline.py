"""Per-CTA residency of the attention forward kernel: globaltimer at entry / exit and %smid of every CTA."""
import os
import sys
from collections import defaultdict

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
n = 5 * H * B
rec = [(t[3 * i], t[3 * i + 1], t[3 * i + 2]) for i in range(n)]
t0 = min(r[0] for r in rec)
end = max(r[2] for r in rec)
print(f"CTAs {n}; kernel span (first entry → last exit) {(end - t0) / 1e3:.1f} us")
dur = sorted((r[2] - r[0]) / 1e3 for r in rec)
print(f"CTA lifetime us: min {dur[0]:.1f}  median {dur[n // 2]:.1f}  p90 {dur[int(n * .9)]:.1f}  max {dur[-1]:.1f}")
per_sm = defaultdict(list)
for r in rec:
    per_sm[r[1]].append(((r[0] - t0) / 1e3, (r[2] - t0) / 1e3))
cnt = sorted(len(v) for v in per_sm.values())
print(f"SMs {len(per_sm)}; CTAs per SM min {cnt[0]} median {cnt[len(cnt) // 2]} max {cnt[-1]}")
last = sorted(max(e for _, e in v) for v in per_sm.values())
print(f"per-SM finish time us: min {last[0]:.1f} median {last[len(last) // 2]:.1f} max {last[-1]:.1f}")
starts = sorted((r[0] - t0) / 1e3 for r in rec)
print("entry times us (every 60th CTA):", [round(x, 1) for x in starts[::60]])
sm0 = sorted(per_sm[rec[0][1]])
print("CTAs on the SM of CTA 0 (entry, exit us):", [(round(a, 1), round(b, 1)) for a, b in sm0])
