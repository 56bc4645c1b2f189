import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chest_x_ray_vit_b200 as pkg
ops = pkg.ops
B, T, D = 16, 577, 768
dh = torch.randn(B, T, D, device="cuda").to(torch.bfloat16)
dpos = torch.zeros(T, D, device="cuda"); dcls = torch.zeros(D, device="cuda"); dbias = torch.zeros(D, device="cuda")
dpatch = torch.empty(B * (T - 1), D, device="cuda", dtype=torch.bfloat16)
for _ in range(3): ops.embed_bwd(dh, B, T, D, dpos, dcls, dbias, dpatch)
torch.cuda.synchronize(); torch.cuda._sleep(2_000_000)
evs = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.embed_bwd(dh, B, T, D, dpos, dcls, dbias, dpatch); e1.record(); evs.append((e0, e1))
torch.cuda.synchronize()
print("embed_bwd us:", sorted(a.elapsed_time(b) * 1e3 for a, b in evs)[5])
