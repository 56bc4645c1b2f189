"""Per-bucket phase timing of PeerGradSync inside a real training step (run under torchrun, N >= 2)."""
import os, sys, torch, torch.distributed as dist
os.environ["VITK_SYNC_TIMING"] = "1"      # the plan's bucket events carry timestamps
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chest_x_ray_vit_b200 as pkg
from chest_x_ray_vit_b200.parallel import PeerGradSync, broadcast_parameters
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
torch.manual_seed(0)
m = pkg.ViTForImageClassification(pkg.ViTConfig()).cuda().train()
broadcast_parameters(m)
gs = PeerGradSync.attach(m, layers_per_bucket=(3, 3, 3, 2, 1))
opt = pkg.VitkAdamW(m, lr=2e-5, max_grad_norm=1.0)
x = torch.randint(0, 256, (16, 384, 384), dtype=torch.uint8).cuda()
y = (torch.rand(16, 14) < 0.1).float().cuda()
def step():
    out = m(pixel_values=x, labels=y); out.loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(6): step()
torch.cuda.synchronize(); dist.barrier()
gs.timing = []
t0 = torch.cuda.Event(enable_timing=True); t0.record()
step()
t1 = torch.cuda.Event(enable_timing=True); t1.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"step {t0.elapsed_time(t1):.3f} ms; per bucket: ready@ (ms since step start) | wait-for-ready→start, barrier1, pull, mean, barrier2, pull2 (us)")
    for s, e, ready, mk in gs.timing:
        d = [1e3 * mk[i].elapsed_time(mk[i + 1]) for i in range(len(mk) - 1)]
        print(f"  [{(e - s) * 4 / 1e6:6.1f} MB] ready@{t0.elapsed_time(ready):6.3f} start@{t0.elapsed_time(mk[0]):6.3f} end@{t0.elapsed_time(mk[-1]):6.3f} | " + " ".join(f"{v:7.1f}" for v in d))
gs.timing = None
for mode in ("peer sync", "no sync (each rank alone, same launch plan)"):
    if mode.startswith("no"):
        gs._reduce = lambda s, e: None
        gs._barrier = lambda: None
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): step()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"20 steps, {mode}: {t.item():.3f} ms/step (max over ranks)")
dist.barrier(); dist.destroy_process_group()
