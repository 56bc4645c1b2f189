"""Input edge of the training path: host batches → device, overlapped with compute.

The reference's DataLoader hands ``collate_fn`` output — ``{"pixel_values": fp32 [B,3,H,W], "labels": fp32 [B,C]}``
(/root/reference/ViT-Training.py:77-80) — to ``model(**batch)`` (HF trainer.py:1978).  ``DeviceFeeder`` wraps any
iterable of such batches (or of uint8 grayscale ``[B,H,W]`` images, which is what ``img.convert("RGB")`` of a chest
X-ray replicates three times, ViT-Training.py:60-66 — 12× fewer bytes over PCIe; the patchify kernel normalises them
on the GPU) and yields batches already resident in HBM:

  * each host batch is staged in pinned memory (a no-op when the loader already pins, ``pin_memory=True``),
  * copied on a dedicated copy stream into one of ``depth`` device slots while the previous batch trains,
  * and handed over with a stream-ordered event (no host synchronisation); a slot is reused only after the compute
    stream has finished with it.

Normalisation / im2col happen in ``vitk_patchify_u8`` / ``vitk_patchify_f32``.  The one pixel operation offered here is
the training transform's ``RandomHorizontalFlip`` (ViT-Training.py:61) for uint8 batches: ``hflip_p > 0`` draws a
per-image mask on the host (``generator`` makes it reproducible), copies it with the batch and mirrors the selected
images in place on the copy stream (``vitk_hflip_u8``) — a byte permutation, identical to flipping before
ToTensor+Normalize.  ``RandomResizedCrop`` (PIL's antialiased resampling) stays with the host loader.
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Optional

import torch


class DeviceFeeder:
    def __init__(self, batches: Iterable[Dict[str, torch.Tensor]], device: Optional[torch.device] = None, depth: int = 2,
                 hflip_p: float = 0.0, generator: Optional[torch.Generator] = None, image_key: str = "pixel_values"):
        if depth < 2:
            raise ValueError("DeviceFeeder: depth must be >= 2 (one slot in use, one in flight)")
        if not 0.0 <= hflip_p <= 1.0:
            raise ValueError("DeviceFeeder: hflip_p must be a probability")
        self.hflip_p, self.generator, self.image_key = float(hflip_p), generator, image_key
        self._mask_host: List[Optional[torch.Tensor]] = [None] * depth
        self._mask_dev: List[Optional[torch.Tensor]] = [None] * depth
        self.last_masks: List[Optional[torch.Tensor]] = [None] * depth      # host copy of each slot's latest flip mask
        self.batches = batches
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots: List[Optional[Dict[str, torch.Tensor]]] = [None] * depth
        self._pinned: List[Optional[Dict[str, torch.Tensor]]] = [None] * depth
        self._ready = [torch.cuda.Event() for _ in range(depth)]
        self._freed = [torch.cuda.Event() for _ in range(depth)]
        self.h2d_bytes = 0              # bytes copied host → device so far
        self.batches_fed = 0

    @staticmethod
    def _like(t: torch.Tensor, ref: Optional[torch.Tensor], **kw) -> torch.Tensor:
        if ref is not None and ref.shape == t.shape and ref.dtype == t.dtype:
            return ref
        return torch.empty(t.shape, dtype=t.dtype, **kw)

    def _stage(self, slot: int, batch: Dict[str, torch.Tensor]) -> None:
        dev = self._slots[slot] or {}
        pin = self._pinned[slot] or {}
        new_dev, new_pin = {}, {}
        # device slots are allocated on the CONSUMER's stream (the caching allocator pools memory per stream: a slot
        # allocated under the copy stream could not be recycled by the next feeder's, and every new feeder would
        # cudaMalloc); the copy stream's writes into them are ordered by the freed/ready events
        for k, t in batch.items():
            if isinstance(t, torch.Tensor) and not t.is_cuda:
                dev[k] = self._like(t, dev.get(k), device=self.device)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._freed[slot])          # compute is done with this slot's previous batch
            for k, t in batch.items():
                if not isinstance(t, torch.Tensor):
                    continue
                if t.is_cuda:
                    new_dev[k] = t
                    continue
                if not t.is_pinned():
                    p = self._like(t, pin.get(k), pin_memory=True)
                    # the slot's pinned buffer may still be the source of its previous copy: that copy was enqueued
                    # before _freed[slot] on this stream only in program order, so wait for it on the host
                    self._ready[slot].synchronize()
                    p.copy_(t)
                    new_pin[k] = p
                    t = p
                d = dev[k]
                d.copy_(t, non_blocking=True)
                self.h2d_bytes += t.numel() * t.element_size()
                new_dev[k] = d
                if self.hflip_p > 0.0 and k == self.image_key:
                    self._flip(slot, d)
            self._ready[slot].record(self.copy_stream)
        self._slots[slot], self._pinned[slot] = new_dev, new_pin

    def _flip(self, slot: int, images: torch.Tensor) -> None:
        """RandomHorizontalFlip of the slot's uint8 batch on the copy stream (called under it, after the batch's copy)."""
        if images.dtype != torch.uint8 or images.dim() != 3:
            raise ValueError("DeviceFeeder: hflip_p needs uint8 grayscale batches [B,H,W] (the fp32 collate output is "
                             "already normalised and replicated; flip it in the loader)")
        from . import ops
        B = images.shape[0]
        if self._mask_host[slot] is None or self._mask_host[slot].numel() != B:
            self._mask_host[slot] = torch.empty(B, dtype=torch.uint8, pin_memory=True)
            self._mask_dev[slot] = torch.empty(B, dtype=torch.uint8, device=self.device)
        self._ready[slot].synchronize()                 # the previous copy out of this pinned mask has completed
        mask = (torch.rand(B, generator=self.generator) < self.hflip_p).to(torch.uint8)
        self._mask_host[slot].copy_(mask)
        self.last_masks[slot] = mask
        self._mask_dev[slot].copy_(self._mask_host[slot], non_blocking=True)
        self.h2d_bytes += B
        ops.hflip_u8(images, self._mask_dev[slot])

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        it = iter(self.batches)
        cur = torch.cuda.current_stream(self.device)
        for s in range(self.depth):
            self._freed[s].record(cur)
        pending: List[int] = []
        nxt = 0
        for _ in range(self.depth - 1):                             # prime the pipeline
            b = next(it, None)
            if b is None:
                break
            self._stage(nxt, b)
            pending.append(nxt)
            nxt = (nxt + 1) % self.depth
        while pending:
            b = next(it, None)
            if b is not None:                                       # keep one copy in flight while this batch trains
                self._stage(nxt, b)
                pending.append(nxt)
                nxt = (nxt + 1) % self.depth
            slot = pending.pop(0)
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._ready[slot])
            self.batches_fed += 1
            yield self._slots[slot]
            self._freed[slot].record(torch.cuda.current_stream(self.device))
