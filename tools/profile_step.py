"""One profiled training step (ViT-B/16@384, batch 16) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off`.  Usage: python tools/profile_step.py [batch]"""
import os
import sys

import torch

os.environ.setdefault("VITK_PLAN_GRAPHS", "0")   # kernel-by-kernel launches: ncu sees the step in launch order

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chest_x_ray_vit_b200 as pkg  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
m = pkg.ViTForImageClassification(pkg.ViTConfig()).cuda().train()
opt = pkg.VitkAdamW(m, lr=2e-5, max_grad_norm=1.0)
g = torch.Generator().manual_seed(1)
x = torch.randint(0, 256, (B, 384, 384), dtype=torch.uint8, generator=g).cuda()
y = (torch.rand(B, 14, generator=g) < 0.1).float().cuda()


def step():
    out = m(pixel_values=x, labels=y)
    out.loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    return out.loss


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss.detach()))
