"""CPU-only checks of the drop-in module's host logic: HF-compatible names/shapes, state-dict
round trip with the live HF model, flat layout invariants, input validation."""
import math

import pytest
import torch

import chest_x_ray_vit_b200 as pkg
from oracle import vit_oracle as O

TINY = dict(image_size=64, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, num_labels=14)


def test_names_and_shapes_match_hf_inventory():
    m = pkg.ViTForImageClassification(pkg.ViTConfig())
    sd = m.state_dict()
    ref = O.param_shapes(O.VIT_B16_384)
    assert list(sd.keys()) == list(ref.keys())           # same names, same order as HF named_parameters()
    assert all(tuple(sd[k].shape) == ref[k] for k in ref)
    assert sum(v.numel() for v in sd.values()) == 86_101_262


def test_state_dict_round_trip_with_live_hf():
    tr = pytest.importorskip("transformers")
    hc = tr.ViTConfig(**TINY, problem_type="multi_label_classification")
    hf = tr.ViTForImageClassification(hc)
    m = pkg.ViTForImageClassification(pkg.ViTConfig.from_hf(hc))
    res = m.load_state_dict(hf.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in hf.state_dict().items():
        assert torch.equal(m.state_dict()[k], v)
    # and back
    assert not hf.load_state_dict(m.state_dict(), strict=True).missing_keys
    # parameters stayed views of the flat buffer (load_state_dict copies in place)
    flat = m.flat_parameters()
    q = m.vit.encoder.layer[1].attention.attention.query.weight
    assert q.data_ptr() == flat.data_ptr() + 4 * m.layout.offset["vit.encoder.layer.1.attention.attention.query.weight"]


def test_flat_layout_invariants():
    cfg = pkg.ViTConfig()
    lay = pkg.modeling.FlatLayout(cfg)
    D, F = cfg.hidden_size, cfg.intermediate_size
    spans = sorted((lay.offset[n], lay.offset[n] + math.prod(lay.shapes[n])) for n in lay.names)
    assert all(a % 8 == 0 for a, _ in spans)
    assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))       # no overlap
    for i in range(cfg.num_hidden_layers):
        p = f"vit.encoder.layer.{i}.attention.attention."
        assert lay.offset[p + "key.weight"] == lay.offset[p + "query.weight"] + D * D      # fused [3D, D] view
        assert lay.offset[p + "value.weight"] == lay.offset[p + "query.weight"] + 2 * D * D
        assert lay.offset[p + "key.bias"] == lay.offset[p + "query.bias"] + D
        s, e = lay.layer_range[i]
        assert s == lay.offset[p + "query.weight"] and e - s == 4 * D * D + 2 * D * F
    # buckets tile the whole buffer exactly once
    cover = sorted(list(lay.layer_range) + list(lay.rest_ranges))
    assert cover[0][0] == 0 and cover[-1][1] == lay.total
    assert all(cover[i][1] == cover[i + 1][0] for i in range(len(cover) - 1))
    # decay / no-decay split follows HF (bias and LayerNorm excluded)
    for n in lay.names:
        nodecay = n.endswith("bias") or "layernorm" in n
        assert (lay.offset[n] >= lay.decay_end) == nodecay, n


def test_init_follows_hf_scheme():
    m = pkg.ViTForImageClassification(pkg.ViTConfig(**TINY))
    sd = m.state_dict()
    assert sd["classifier.bias"].abs().max() == 0
    assert torch.equal(sd["vit.layernorm.weight"], torch.ones(128))
    w = sd["vit.encoder.layer.0.intermediate.dense.weight"]
    assert 0.015 < w.std() < 0.025 and w.abs().max() <= 0.04 + 1e-6


def test_input_validation_matches_hf_errors():
    m = pkg.ViTForImageClassification(pkg.ViTConfig(**TINY))
    with pytest.raises(ValueError, match="specify pixel_values"):
        m()
    with pytest.raises(ValueError, match="channel dimension"):
        m(pixel_values=torch.zeros(1, 1, 64, 64))
    with pytest.raises(ValueError, match="doesn't match model"):
        m(pixel_values=torch.zeros(1, 3, 32, 32))
    with pytest.raises(ValueError, match="interpolate_pos_encoding"):
        m(pixel_values=torch.zeros(1, 3, 64, 64), interpolate_pos_encoding=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(pixel_values=torch.zeros(1, 3, 64, 64))
    assert m.main_input_name == "pixel_values"
    import inspect
    sig = inspect.signature(m.forward)
    assert list(sig.parameters)[:3] == ["pixel_values", "labels", "interpolate_pos_encoding"]
    assert any(p.kind == p.VAR_KEYWORD for p in sig.parameters.values())     # Trainer passes num_items_in_batch


@pytest.mark.parametrize("bad", [dict(hidden_dropout_prob=0.1), dict(hidden_act="relu"), dict(patch_size=32),
                                 dict(hidden_size=192, num_attention_heads=3), dict(problem_type="regression")])
def test_unsupported_configs_are_rejected(bad):
    with pytest.raises(ValueError, match="unsupported config"):
        pkg.ViTForImageClassification(pkg.ViTConfig(**{**TINY, **bad}))


def test_output_object_access_patterns():
    o = pkg.ImageClassifierOutput(loss=torch.tensor(1.0), logits=torch.zeros(2, 3))
    assert o["loss"] is o.loss and o[0] is o.loss and o[1] is o.logits and len(o) == 2
    o2 = pkg.ImageClassifierOutput(logits=torch.zeros(2, 3))
    assert o2[0] is o2.logits and "loss" not in o2
    with pytest.raises(KeyError):
        o2["loss"]
