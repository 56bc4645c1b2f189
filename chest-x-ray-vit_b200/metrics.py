"""Evaluation metrics of the reference on the device.

/root/reference/ViT-Training.py:112-118 computes, per evaluation, ``f1_score(labels, sigmoid(logits) >= 0.5,
average="micro", zero_division=0)`` and :139-146 a per-class ``classification_report`` on the host from the
concatenated logits of the whole split.  All of that is a function of four integers per class (TP, FP, FN, TN), which
``vitk_multilabel_counts`` accumulates batch by batch next to the logits in HBM; the host reads 14×4 integers once.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops


class MultilabelCounter:
    def __init__(self, num_labels: int, threshold: float = 0.5, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.threshold = float(threshold)
        self.counts = torch.zeros((num_labels, 4), dtype=torch.int64, device=self.device)

    def reset(self) -> None:
        self.counts.zero_()

    def update(self, logits: torch.Tensor, labels: torch.Tensor) -> None:
        """logits fp32 [B,C] on the device (model output), labels {0,1} [B,C] (any float/int dtype)."""
        labels = labels.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        ops.multilabel_counts(logits.to(torch.float32).contiguous(), labels, self.counts, self.threshold)

    def compute(self) -> Dict[str, object]:
        """One device→host read.  Returns sklearn's numbers with zero_division=0: ``f1_micro`` (the reference's
        metric_for_best_model), micro/macro/weighted precision-recall-F1, and per-class precision/recall/f1/support."""
        return scores_from_counts(self.counts.cpu())


def scores_from_counts(counts: torch.Tensor) -> Dict[str, object]:
    c = counts.to(torch.float64)
    tp, fp, fn = c[:, 0], c[:, 1], c[:, 2]

    def div(a, b):
        return torch.where(b > 0, a / b.clamp_min(1), torch.zeros_like(a))

    prec, rec = div(tp, tp + fp), div(tp, tp + fn)
    f1 = div(2 * tp, 2 * tp + fp + fn)
    support = tp + fn
    TP, FP, FN = tp.sum(), fp.sum(), fn.sum()
    out = {"f1_micro": div(2 * TP, 2 * TP + FP + FN).item(), "precision_micro": div(TP, TP + FP).item(),
           "recall_micro": div(TP, TP + FN).item(), "f1_macro": f1.mean().item(), "precision_macro": prec.mean().item(),
           "recall_macro": rec.mean().item(),
           "f1_weighted": div((f1 * support).sum(), support.sum()).item(),
           "per_class": {"precision": prec.tolist(), "recall": rec.tolist(), "f1": f1.tolist(), "support": support.to(torch.int64).tolist()},
           "counts": counts.tolist()}
    return out
