"""Instrumented DeviceFeeder: per-statement host time inside _stage / __iter__ when nothing else loads the GPU."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chest_x_ray_vit_b200 as pkg
from chest_x_ray_vit_b200 import data as D
B = 16
xh = [torch.randn(B, 3, 384, 384).pin_memory() for _ in range(4)]
yh = [(torch.rand(B, 14) < 0.1).float().pin_memory() for _ in range(4)]
acc = {}
def T(): return time.perf_counter()
def add(k, t0): acc[k] = acc.get(k, 0.0) + (T() - t0)
orig_like = D.DeviceFeeder._like
def like(t, ref, **kw):
    t0 = T(); r = orig_like(t, ref, **kw); add("_like(" + ("pin" if "pin_memory" in kw else "dev") + (",reuse" if r is ref else ",alloc") + ")", t0); return r
D.DeviceFeeder._like = staticmethod(like)
orig_stage = D.DeviceFeeder._stage
def stage(self, slot, batch):
    t0 = T(); orig_stage(self, slot, batch); add("_stage total", t0)
D.DeviceFeeder._stage = stage
orig_copy = torch.Tensor.copy_
def copy_(self, src, non_blocking=False):
    t0 = T(); r = orig_copy(self, src, non_blocking=non_blocking); add(f"copy_ {tuple(self.shape)} pinned_src={src.is_pinned() if not src.is_cuda else 'cuda'}", t0); return r
torch.Tensor.copy_ = copy_
for rep in range(3):
    acc.clear()
    torch.cuda.synchronize()
    n = 20
    t0 = T()
    f = D.DeviceFeeder(({"pixel_values": xh[i % 4], "labels": yh[i % 4]} for i in range(n)))
    add("ctor", t0)
    t0 = T()
    for b in f:
        pass
    add("loop total", t0)
    t0 = T(); torch.cuda.synchronize(); add("final sync", t0)
    print(f"rep {rep}: " + " | ".join(f"{k} {1e3 * v:.2f} ms" for k, v in acc.items()))
