"""World-size-2 gloo tests (CPU) of the data-parallel host logic: bucket schedule of GradSync,
parameter broadcast, batch sharding, and the identity the design relies on (mean of equal-shard
gradients == full-batch gradient of the mean loss), checked with the fp32 oracle."""
import os
import socket

import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import chest_x_ray_vit_b200 as pkg
from chest_x_ray_vit_b200.parallel import GradSync, shard_batch
from oracle import vit_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["PYTHONPATH"] = ROOT + os.pathsep + os.environ.get("PYTHONPATH", "")      # spawned ranks re-import this module
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TINY = dict(image_size=64, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, num_labels=14)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lay = pkg.modeling.FlatLayout(pkg.ViTConfig(**TINY))
        flat = torch.arange(lay.total, dtype=torch.float32) * (rank + 1)        # rank-dependent "gradients"
        gs = GradSync(lay.layer_range, lay.rest_ranges, layers_per_bucket=1)
        gs.begin(flat)
        for l in reversed(range(len(lay.layer_range))):                          # backward finishes layers top-down
            gs.layer_ready(l)
        gs.rest_ready()
        expect = torch.arange(lay.total, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
        ok_mean = torch.allclose(flat, expect)
        ok_bytes = gs.bytes_reduced == lay.total * 4                             # every element reduced exactly once
        # coalesced buckets give the same result with fewer collectives
        flat2 = torch.arange(lay.total, dtype=torch.float32) * (rank + 1)
        gs2 = GradSync(lay.layer_range, lay.rest_ranges, layers_per_bucket=2)
        gs2.begin(flat2)
        for l in reversed(range(len(lay.layer_range))):
            gs2.layer_ready(l)
        gs2.rest_ready()
        ok_coalesced = torch.allclose(flat2, expect) and gs2.collectives < gs.collectives
        # tapered schedule (sizes in readiness order, last repeats) on a 5-layer layout: buckets {4,3}, {2}, {1}, then the final
        # operation: {0} coalesced with the patch weights in front of it, and the non-GEMM tail
        lay5 = pkg.modeling.FlatLayout(pkg.ViTConfig(**dict(TINY, num_hidden_layers=5)))
        flat3 = torch.arange(lay5.total, dtype=torch.float32) * (rank + 1)
        gs3 = GradSync(lay5.layer_range, lay5.rest_ranges, layers_per_bucket=(2, 1))
        gs3.begin(flat3)
        for l in reversed(range(5)):
            gs3.layer_ready(l)
        gs3.rest_ready()
        expect5 = torch.arange(lay5.total, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
        ok_coalesced = ok_coalesced and torch.allclose(flat3, expect5) and gs3.collectives == 3 + len(lay5.rest_ranges) \
            and gs3.bytes_reduced == lay5.total * 4
        # broadcast of flat parameters
        m = pkg.ViTForImageClassification(pkg.ViTConfig(**TINY))
        if rank != 0:
            with torch.no_grad():
                m.flat_parameters().add_(1.0)
        m._shadow_version = m._param_version()        # pretend the bf16 shadow of the pre-broadcast weights is current
        pkg.parallel.broadcast_parameters(m)
        ref = [torch.zeros_like(m.flat_parameters()) for _ in range(world)]
        dist.all_gather(ref, m.flat_parameters())
        ok_bcast = all(torch.equal(ref[0], r) for r in ref)
        # ... and shadow() must re-derive the bf16 weights afterwards: c10d collectives do not bump tensor version
        # counters, so the broadcast invalidates the shadow explicitly (it used to leave non-source ranks computing with
        # their old bf16 weights)
        ok_bcast = ok_bcast and m._shadow_version == -1
        q.put((rank, ok_mean, ok_bytes, ok_coalesced, ok_bcast))
    except Exception as e:          # surface the failure instead of a queue timeout
        q.put((rank, False, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_gradsync_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, *oks in res:
        assert all(oks), (rank, oks)


def test_bucket_end_layers_follow_the_taper():
    lay = pkg.modeling.FlatLayout(pkg.ViTConfig(**dict(TINY, num_hidden_layers=12)))
    gs = GradSync(lay.layer_range, lay.rest_ranges, layers_per_bucket=(3, 3, 3, 2, 1))
    assert gs.bucket_end_layers() == {9, 6, 3, 1, 0}            # buckets {11,10,9} {8,7,6} {5,4,3} {2,1} {0}
    assert GradSync(lay.layer_range, lay.rest_ranges, layers_per_bucket=5).bucket_end_layers() == {7, 2, 0}
    # layers_ready(lo, hi) reduces exactly the bucket's contiguous range (world size 1: counted, not communicated)
    seen = []
    gs._reduce = lambda s, e: seen.append((s, e))
    gs.begin(torch.zeros(lay.total))
    gs.layers_ready(9, 11)
    assert seen == [(lay.layer_range[9][0], lay.layer_range[11][1])]


def test_shard_batch():
    assert shard_batch(128, 3, 8) == (48, 64)
    with pytest.raises(ValueError):
        shard_batch(30, 0, 4)


def test_mean_of_shard_gradients_is_global_gradient():
    cfg = O.TINY
    p = O.init_params(cfg, 0, 123)
    g = torch.Generator().manual_seed(3)
    x8, y = O.synth_inputs(cfg, 4, g)
    x = O.normalize_gray(x8)
    _, _, full = O.forward_backward(p, cfg, x, y)
    parts = [O.forward_backward(p, cfg, x[i:i + 2], y[i:i + 2])[2] for i in (0, 2)]
    for k in full:
        avg = (parts[0][k] + parts[1][k]) / 2
        assert torch.allclose(avg, full[k], atol=1e-6, rtol=1e-4), k
