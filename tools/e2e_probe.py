"""Where does the end-to-end (host buffers → DeviceFeeder → step → loss read-back) loop spend host time?"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chest_x_ray_vit_b200 as pkg
from chest_x_ray_vit_b200.data import DeviceFeeder
B = 16
torch.manual_seed(0)
m = pkg.ViTForImageClassification(pkg.ViTConfig()).cuda().train()
opt = pkg.VitkAdamW(m, lr=2e-5, max_grad_norm=1.0)
xh = [torch.randn(B, 3, 384, 384).pin_memory() for _ in range(4)]
yh = [(torch.rand(B, 14) < 0.1).float().pin_memory() for _ in range(4)]
xd, yd = [t.cuda() for t in xh], [t.cuda() for t in yh]
print("pinned?", xh[0].is_pinned(), yh[0].is_pinned())
def step(x, y):
    out = m(pixel_values=x, labels=y); out.loss.backward(); opt.step(); opt.zero_grad(set_to_none=True); return out.loss
for i in range(5): step(xd[i % 4], yd[i % 4])
torch.cuda.synchronize()
def timed(name, fn, n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(n); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name:34s} host {1e3*(t1-t0)/n:6.2f} ms/step   wall {1e3*(t2-t0)/n:6.2f} ms/step")
def resident(n):
    for i in range(n): step(xd[i % 4], yd[i % 4])
def copies_only(n):
    s = torch.cuda.Stream(); d = torch.empty_like(xd[0])
    with torch.cuda.stream(s):
        for i in range(n): d.copy_(xh[i % 4], non_blocking=True)
def feeder(n):
    f = DeviceFeeder(({"pixel_values": xh[i % 4], "labels": yh[i % 4]} for i in range(n)))
    for b in f: step(b["pixel_values"], b["labels"])
def feeder_noop(n):
    f = DeviceFeeder(({"pixel_values": xh[i % 4], "labels": yh[i % 4]} for i in range(n)))
    for b in f: pass
gs = pkg.graph.GraphedTrainStep(m, opt)
def graphed(n):
    for i in range(n): gs(xd[i % 4], yd[i % 4])
def feeder_graphed(n):
    f = DeviceFeeder(({"pixel_values": xh[i % 4], "labels": yh[i % 4]} for i in range(n)))
    for b in f: gs(b["pixel_values"], b["labels"])
graphed(3)
for name, fn in (("device-resident eager", resident), ("H2D copies only (28 MB each)", copies_only), ("feeder only", feeder_noop),
                 ("feeder + eager step", feeder), ("device-resident graphed", graphed), ("feeder + graphed step", feeder_graphed)):
    timed(name, fn)
    timed(name, fn)
