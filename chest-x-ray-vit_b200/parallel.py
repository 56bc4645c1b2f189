"""Data-parallel gradient synchronisation for the batch-sharded training path.

The reference trains 8 replicas with gradients averaged implicitly inside the XLA optimizer
step (/root/reference/ViT-Training.py:106,165,170; HF trainer.py:1760,1796).  Here every rank
owns one GPU, the batch is sharded by rank, and the only collective of the path is an
all-reduce (mean) of the flat fp32 gradient buffer, issued bucket by bucket while backward is
still running: the weight gradients of adjacent encoder layers are contiguous in the flat layout
(28.3 MB per ViT-B layer) and are coalesced into buckets of ``layers_per_bucket`` layers (default
3; a sequence such as (3, 3, 3, 2, 1) tapers them) in the order backward finishes them, then one
bucket for everything else.  NCCL runs the buckets on its own stream over NVLink/NVSwitch; the
compute stream only waits after the last bucket, right before the optimizer needs the gradients.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, layer_ranges: Sequence[Tuple[int, int]], rest_ranges: Sequence[Tuple[int, int]],
                 process_group: Optional[dist.ProcessGroup] = None, layers_per_bucket=3):
        self.layer_ranges = list(layer_ranges)
        self.rest_ranges = list(rest_ranges)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        # int: uniform buckets; sequence: bucket sizes in the order the layers finish (top layer first), the last
        # entry repeating — e.g. (3, 3, 3, 2, 1) tapers the buckets so the all-reduce left after backward is short
        sizes = [layers_per_bucket] if isinstance(layers_per_bucket, int) else list(layers_per_bucket)
        self.bucket_sizes = [max(1, int(x)) for x in sizes] or [3]
        self.bucket_index = 0
        self.flat: Optional[torch.Tensor] = None
        self.works: List = []
        self.pending: List[int] = []
        self.bytes_reduced = 0
        self.collectives = 0

    @classmethod
    def attach(cls, model, process_group=None, layers_per_bucket=3) -> "GradSync":
        gs = cls(model.layout.layer_range, model.layout.rest_ranges, process_group, layers_per_bucket)
        eng = model.engine()
        eng.grad_sync = gs
        eng.arenas.clear()          # launch plans bake the SM budget left to the collective
        return gs

    # called by Engine.backward ------------------------------------------------
    def begin(self, model_or_flat) -> None:
        self.flat = model_or_flat if isinstance(model_or_flat, torch.Tensor) else model_or_flat.flat_grads()
        self.works, self.pending = [], []
        self.bucket_index = 0

    def layer_ready(self, l: int) -> None:
        """Layer ``l``'s weight gradients are enqueued on the current stream (layers finish in
        descending order).  Adjacent layers are coalesced into one contiguous bucket."""
        self.pending.append(l)
        size = self.bucket_sizes[min(self.bucket_index, len(self.bucket_sizes) - 1)]
        if len(self.pending) >= size or l == 0:
            lo, hi = min(self.pending), max(self.pending)
            self._reduce(self.layer_ranges[lo][0], self.layer_ranges[hi][1])
            self.pending = []
            self.bucket_index += 1

    def rest_ready(self) -> None:
        for s, e in self.rest_ranges:
            self._reduce(s, e)
        self.wait()

    def wait(self) -> None:
        for w in self.works:
            w.wait()
        self.works = []

    # ---------------------------------------------------------------------------
    def _reduce(self, start: int, end: int) -> None:
        if self.world == 1 or end <= start:
            return
        t = self.flat[start:end]
        self.bytes_reduced += t.numel() * t.element_size()
        self.collectives += 1
        if t.is_cuda:
            self.works.append(dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
        else:   # gloo (CPU tests) has no AVG
            w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            w.wait()
            t.mul_(1.0 / self.world)


def broadcast_parameters(model, src: int = 0, process_group=None) -> None:
    """Identical replicas: rank ``src``'s flat fp32 parameters overwrite everyone else's."""
    if dist.is_initialized() and dist.get_world_size(process_group) > 1:
        dist.broadcast(model.flat_parameters(), src=src, group=process_group)
        model.invalidate_shadow()       # the bf16 GEMM weights must be re-derived from the received masters


def shard_batch(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous equal shards of a global batch (rank r owns images [r·n/world, (r+1)·n/world))."""
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by world size {world}: the mean-loss gradient "
                         "average is only exact for equal shards")
    per = n // world
    return rank * per, (rank + 1) * per
