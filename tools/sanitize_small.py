"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck): two training steps of the tiny
config (eager plan, then the CUDA-graph replay of the same plan), a no-grad forward, the evaluation counters."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chest_x_ray_vit_b200 as pkg
from oracle import vit_oracle as O
cfg = O.TINY
m = pkg.ViTForImageClassification(pkg.ViTConfig(image_size=64, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                                intermediate_size=256, num_labels=14))
m.load_state_dict(O.init_params(cfg, 0, 123))
m = m.cuda().train()
opt = pkg.VitkAdamW(m, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
g = torch.Generator().manual_seed(1)
x8, y = O.synth_inputs(cfg, 3, g)
x, yd = x8[:, 0].cuda(), y.cuda()
for _ in range(3):
    out = m(pixel_values=x, labels=yd)
    out.loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
m.eval()
with torch.no_grad():
    logits = m(pixel_values=O.normalize_gray(x8).cuda()).logits
ctr = pkg.metrics.MultilabelCounter(14)
ctr.update(logits, yd)
torch.cuda.synchronize()
print("ok", float(out.loss), ctr.compute()["f1_micro"])
