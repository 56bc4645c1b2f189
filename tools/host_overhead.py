"""Host (enqueue) time per training step vs device time: is the step CPU-bound?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chest_x_ray_vit_b200 as pkg
B = 16
torch.manual_seed(0)
m = pkg.ViTForImageClassification(pkg.ViTConfig()).cuda().train()
opt = pkg.VitkAdamW(m, lr=2e-5, max_grad_norm=1.0)
x = torch.randint(0, 256, (B, 384, 384), dtype=torch.uint8).cuda()
y = (torch.rand(B, 14) < 0.1).float().cuda()
def step():
    out = m(pixel_values=x, labels=y); out.loss.backward(); opt.step(); opt.zero_grad(set_to_none=True); return out.loss
for _ in range(5): step()
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
host = []
for _ in range(N):
    a = time.perf_counter(); step(); host.append(time.perf_counter() - a)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue {1e3*sum(host)/N:.2f} ms/step (min {1e3*min(host):.2f}), wall incl. drain {(t2-t0)/N*1e3:.2f} ms/step, queue drain after loop {1e3*(t2-t1):.1f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
