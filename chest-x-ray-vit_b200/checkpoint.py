"""HF checkpoint interchange for the drop-in module.

The reference starts from ``ViTForImageClassification.from_pretrained("google/vit-base-patch16-384", num_labels=14,
id2label=…, label2id=…, ignore_mismatched_sizes=True, problem_type="multi_label_classification")``
(/root/reference/ViT-Training.py:83-90) and Trainer writes ``save_pretrained`` directories every epoch (:106).
The on-disk format is HF's: ``config.json`` + ``model.safetensors`` (or ``pytorch_model.bin``) with the state-dict keys of
SURVEY App. B.3 — the same names this module's parameters carry, so interchange is a key-for-key copy.

``ignore_mismatched_sizes=True`` follows HF modeling_utils' behaviour: a checkpoint tensor whose shape differs from the
model's (the 1000-class ImageNet head vs 14 labels) is dropped and the parameter keeps its fresh initialisation
(``_init_weights``, modeling_vit.py:385-398: trunc_normal(std=initializer_range) weight, zero bias); without the flag a
size mismatch raises.  Checkpoints of the bare ``ViTModel`` (keys without the ``vit.`` prefix, plus a pooler this path
never uses) load through the same base-model-prefix rule HF applies.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Union

import torch

_CONFIG_KEYS = ("hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size", "hidden_act",
                "hidden_dropout_prob", "attention_probs_dropout_prob", "initializer_range", "layer_norm_eps", "image_size",
                "patch_size", "num_channels", "qkv_bias")


def _read_state_dict(path: str) -> Dict[str, torch.Tensor]:
    st = os.path.join(path, "model.safetensors") if os.path.isdir(path) else path
    if os.path.isdir(path) and not os.path.exists(st):
        st = os.path.join(path, "pytorch_model.bin")
    if not os.path.exists(st):
        raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin under {path}")
    if st.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(st)
    return torch.load(st, map_location="cpu", weights_only=True)


def load_pretrained(cls, source: Union[str, Dict[str, torch.Tensor]], *, config=None, num_labels: Optional[int] = None,
                    id2label=None, label2id=None, problem_type: Optional[str] = "multi_label_classification",
                    ignore_mismatched_sizes: bool = False, generator: Optional[torch.Generator] = None, **overrides):
    """``source``: an HF checkpoint directory, a weights file, or a state dict.  Returns ``(model, info)`` where ``info``
    lists ``missing_keys``, ``unexpected_keys`` and ``mismatched_keys`` (name, checkpoint shape, model shape)."""
    from .modeling import ViTConfig
    if isinstance(source, str):
        sd = _read_state_dict(source)
        cfg_path = os.path.join(source, "config.json") if os.path.isdir(source) else None
        if config is None:
            if cfg_path is None or not os.path.exists(cfg_path):
                raise ValueError("from_pretrained: no config.json next to the weights; pass config=")
            raw = json.load(open(cfg_path))
            kw = {k: raw[k] for k in _CONFIG_KEYS if k in raw}
            if "id2label" in raw and num_labels is None and id2label is None:
                kw["num_labels"] = len(raw["id2label"])
                kw["id2label"] = {int(k): v for k, v in raw["id2label"].items()}
                kw["label2id"] = raw.get("label2id")
            if raw.get("problem_type") and problem_type is None:
                kw["problem_type"] = raw["problem_type"]
            config = ViTConfig(**kw)
    else:
        sd = dict(source)
        if config is None:
            raise ValueError("from_pretrained: a state dict needs config=")
    if not isinstance(config, ViTConfig):
        config = ViTConfig.from_hf(config)
    if id2label is not None:
        config.id2label = dict(id2label)
        config.num_labels = len(id2label)
    if label2id is not None:
        config.label2id = dict(label2id)
    if num_labels is not None:
        config.num_labels = int(num_labels)
    if problem_type is not None:
        config.problem_type = problem_type
    for k, v in overrides.items():
        setattr(config, k, v)
    model = cls(config)
    if generator is not None:
        model.reset_parameters(generator)
    # bare ViTModel checkpoints: add the base-model prefix, drop the pooler (HF base_model_prefix = "vit")
    if not any(k.startswith("vit.") for k in sd) and any(k.startswith("embeddings.") for k in sd):
        sd = {("vit." + k): v for k, v in sd.items()}
    own = model.state_dict()
    missing = [k for k in own if k not in sd]
    unexpected = [k for k in sd if k not in own]
    mismatched = [(k, tuple(sd[k].shape), tuple(own[k].shape)) for k in own if k in sd and sd[k].shape != own[k].shape]
    if mismatched and not ignore_mismatched_sizes:
        k, a, b = mismatched[0]
        raise RuntimeError(f"size mismatch for {k}: copying a param with shape {a} from checkpoint, the shape in current model "
                           f"is {b}. Pass ignore_mismatched_sizes=True to keep the freshly initialised parameter instead.")
    skip = {k for k, _, _ in mismatched}
    with torch.no_grad():
        for k, p in own.items():
            if k in sd and k not in skip:
                p.copy_(sd[k].to(torch.float32))
    return model, {"missing_keys": missing, "unexpected_keys": unexpected, "mismatched_keys": mismatched}


def save_pretrained(model, save_directory: str, safe_serialization: bool = True) -> None:
    """Writes ``config.json`` + ``model.safetensors`` that ``transformers.ViTForImageClassification.from_pretrained`` reads."""
    os.makedirs(save_directory, exist_ok=True)
    cfg = model.config
    raw = {k: getattr(cfg, k) for k in _CONFIG_KEYS}
    n = cfg.num_labels
    id2label = cfg.id2label or {i: f"LABEL_{i}" for i in range(n)}
    raw.update({"architectures": ["ViTForImageClassification"], "model_type": "vit", "problem_type": cfg.problem_type,
                "id2label": {str(k): v for k, v in id2label.items()},
                "label2id": cfg.label2id or {v: int(k) for k, v in id2label.items()},
                "encoder_stride": 16, "pooler_act": "tanh", "pooler_output_size": cfg.hidden_size, "dtype": "float32"})
    with open(os.path.join(save_directory, "config.json"), "w") as f:
        json.dump(raw, f, indent=2, sort_keys=True)
    sd = {k: v.detach().to("cpu", torch.float32).clone().contiguous() for k, v in model.state_dict().items()}
    if safe_serialization:
        from safetensors.torch import save_file
        save_file(sd, os.path.join(save_directory, "model.safetensors"), metadata={"format": "pt"})
    else:
        torch.save(sd, os.path.join(save_directory, "pytorch_model.bin"))
