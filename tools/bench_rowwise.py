"""Timing of the HBM-bound kernels at the model's shapes (M = 9232, D = 768) with achieved GB/s against the
algorithmic bytes of DESIGN.md §4.  L2 flushed between launches."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chest_x_ray_vit_b200 as pkg  # noqa: E402

ops = pkg.ops
dev = "cuda"
M, D, F = 9232, 768, 3072
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
HBM = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


x = torch.randn(M, D, device=dev)
gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
y, mean, rstd = ops.layernorm_fwd(x, gamma, beta, 1e-12)
dy = torch.randn(M, D, device=dev).to(torch.bfloat16)
dres = torch.randn(M, D, device=dev).to(torch.bfloat16)
dx = torch.empty_like(dy)
dg, db, ds = torch.zeros(D, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev)
du = torch.randn(M, F, device=dev).to(torch.bfloat16)
dbf = torch.zeros(F, device=dev)
n = 85_000_000 // 8 * 8
p, g, m, v = (torch.zeros(n, device=dev) for _ in range(4))
p16 = torch.empty(n, dtype=torch.bfloat16, device=dev)
rows = [
    ("layernorm_fwd", lambda: ops.layernorm_fwd(x, gamma, beta, 1e-12, y=y, mean=mean, rstd=rstd), M * D * 6),
    ("layernorm_bwd (+dres, +dxsum)", lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dres, dg, db, dx=dx, dxsum=ds), M * D * 10),
    ("colsum [M,3072]", lambda: ops.colsum(du, dbf), M * F * 2),
    ("adamw (85M params, +bf16 shadow, +zero grad)", lambda: ops.adamw(p, g, m, v, p16, n, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0.1, 0.001, None, True), n * 34),
]
for name, fn, nbytes in rows:
    t = timeit(fn)
    print(f"{name:46s} {t * 1e3:8.1f} us  {nbytes / t / 1e6:7.0f} GB/s  {100 * nbytes / t / 1e6 / HBM:5.1f}% of measured HBM peak")
