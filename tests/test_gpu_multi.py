"""Real NCCL data-parallel step on 2 GPUs (skipped on a single-GPU box): bucketed gradient
all-reduce overlapped with backward must give (a) identical parameters on all ranks after a
step and (b) the same gradients as one process running the concatenated batch."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["VITK_ROOT"])
import chest_x_ray_vit_b200 as pkg
from chest_x_ray_vit_b200.parallel import GradSync, broadcast_parameters
from oracle import vit_oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
cfg = O.TINY
m = pkg.ViTForImageClassification(pkg.ViTConfig(image_size=64, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                                intermediate_size=256, num_labels=14))
if rank == 0:
    m.load_state_dict(O.init_params(cfg, 0, 123))
m = m.cuda().train()
broadcast_parameters(m)
gs = GradSync.attach(m)
g = torch.Generator().manual_seed(7)
x8, y = O.synth_inputs(cfg, 4 * world, g)
lo, hi = pkg.parallel.shard_batch(4 * world, rank, world)
opt = pkg.VitkAdamW(m, lr=1e-3)
out = m(pixel_values=x8[lo:hi, 0].cuda(), labels=y[lo:hi].cuda())
out.loss.backward()
torch.cuda.synchronize()
grads = m.flat_grads().clone()
opt.step()
torch.cuda.synchronize()
flat = m.flat_parameters()
ref = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(ref, flat)
same = all(torch.equal(ref[0], r) for r in ref)
ok_grad = True
if rank == 0:
    m2 = pkg.ViTForImageClassification(m.config)
    m2.load_state_dict(O.init_params(cfg, 0, 123))
    m2 = m2.cuda().train()
    m2(pixel_values=x8[:, 0].cuda(), labels=y.cuda()).loss.backward()
    cos = torch.nn.functional.cosine_similarity(m2.flat_grads().double(), grads.double(), dim=0).item()
    ok_grad = cos > 0.9999
    print("RESULT", same, cos, gs.collectives, gs.bytes_reduced)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (same and ok_grad) else 1)
'''


def test_two_gpu_data_parallel_step(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, VITK_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "RESULT True" in r.stdout
