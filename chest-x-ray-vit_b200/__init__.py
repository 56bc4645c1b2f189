"""B200-native (sm_100a) replacement for the ViT-B/16 fine-tuning hot path of
Sam1rShaban1/Chest-X-Ray-ViT: the HuggingFace ViTForImageClassification forward/backward that
ViT-Training.py:83-132 drives.  Public surface:

  ViTConfig, ViTForImageClassification ... drop-in nn.Module (modeling.py)
  ops ................................... tensor-level wrappers over the C ABI (include/vitk.h)
  VitkAdamW ............................. flat-buffer AdamW + grad clip (optim.py)
  GradSync .............................. bucketed NCCL gradient all-reduce (parallel.py)
  graph.GraphedTrainStep / GraphedForward  the same launch plans replayed from one CUDA graph (graph.py)
  data.DeviceFeeder ..................... pinned-memory / copy-stream input edge (data.py)
  checkpoint ............................ HF safetensors interchange: from_pretrained / save_pretrained (checkpoint.py)
  metrics ............................... on-device sigmoid-threshold counters → micro-F1 (metrics.py)
  custom_ops ............................ torch.library custom ops + HF AttentionInterface plug-in
"""
from . import _lib, ops  # noqa: F401

__all__ = ["ops", "_lib"]


def __getattr__(name):
    # heavier modules are imported lazily so `import chest_x_ray_vit_b200` stays cheap
    if name in ("ViTConfig", "ViTForImageClassification", "ImageClassifierOutput"):
        from . import modeling
        return getattr(modeling, name)
    if name == "VitkAdamW":
        from .optim import VitkAdamW
        return VitkAdamW
    if name == "GradSync":
        from .parallel import GradSync
        return GradSync
    if name in ("modeling", "engine", "optim", "parallel", "custom_ops", "graph", "data", "checkpoint", "metrics"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
