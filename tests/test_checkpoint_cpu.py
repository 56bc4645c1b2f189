"""CPU: HF checkpoint interchange — from_pretrained / save_pretrained against the live HF model and the
semantics of the reference's call (/root/reference/ViT-Training.py:83-90: num_labels=14 on a 1000-class checkpoint with
ignore_mismatched_sizes=True)."""
import os

import pytest
import torch

import chest_x_ray_vit_b200 as pkg

TINY = dict(image_size=64, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256)


def _hf(num_labels, **kw):
    tr = pytest.importorskip("transformers")
    return tr, tr.ViTForImageClassification(tr.ViTConfig(**TINY, num_labels=num_labels, **kw))


def test_from_pretrained_reads_hf_directory_and_reinitialises_mismatched_head(tmp_path):
    tr, hf = _hf(1000)                                  # stands in for google/vit-base-patch16-384 (ImageNet-1k head)
    d = str(tmp_path / "ckpt")
    hf.save_pretrained(d)
    assert os.path.exists(os.path.join(d, "model.safetensors"))
    labels = [f"finding_{i}" for i in range(14)]
    with pytest.raises(RuntimeError, match="size mismatch for classifier"):
        pkg.ViTForImageClassification.from_pretrained(d, num_labels=14)
    m, info = pkg.ViTForImageClassification.from_pretrained(
        d, num_labels=14, id2label=dict(enumerate(labels)), label2id={v: i for i, v in enumerate(labels)},
        ignore_mismatched_sizes=True, problem_type="multi_label_classification", output_loading_info=True,
        generator=torch.Generator().manual_seed(0))
    assert sorted(k for k, _, _ in info["mismatched_keys"]) == ["classifier.bias", "classifier.weight"]
    assert not info["missing_keys"] and not info["unexpected_keys"]
    assert m.config.num_labels == 14 and m.config.problem_type == "multi_label_classification" and m.config.id2label[3] == "finding_3"
    hsd = hf.state_dict()
    for k, v in m.state_dict().items():
        if k.startswith("classifier"):
            continue
        assert torch.equal(v, hsd[k]), k
    # the head is freshly initialised the HF way: trunc_normal(std=initializer_range) weight, zero bias
    w, b = m.classifier.weight.detach(), m.classifier.bias.detach()
    assert tuple(w.shape) == (14, 128) and b.abs().max().item() == 0.0
    assert 0.5 * 0.02 < w.std().item() < 1.5 * 0.02 and w.abs().max().item() <= 2 * 0.02 + 1e-6
    # parameters are still views of the flat buffer
    assert m.classifier.weight.data_ptr() == m.flat_parameters().data_ptr() + 4 * m.layout.offset["classifier.weight"]


def test_save_pretrained_is_readable_by_transformers(tmp_path):
    tr = pytest.importorskip("transformers")
    m = pkg.ViTForImageClassification(pkg.ViTConfig(**TINY, num_labels=14))
    d = str(tmp_path / "out")
    m.save_pretrained(d)
    hf = tr.ViTForImageClassification.from_pretrained(d)
    assert hf.config.num_labels == 14 and hf.config.problem_type == "multi_label_classification"
    hsd = hf.state_dict()
    assert list(hsd.keys()) == list(m.state_dict().keys())
    for k, v in m.state_dict().items():
        assert torch.equal(v, hsd[k]), k
    # and our own reader round-trips it (config comes from config.json)
    m2 = pkg.ViTForImageClassification.from_pretrained(d)
    assert m2.config.num_labels == 14 and m2.config.image_size == 64
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_bare_vit_model_checkpoint_and_state_dict_source(tmp_path):
    tr = pytest.importorskip("transformers")
    base = tr.ViTModel(tr.ViTConfig(**TINY), add_pooling_layer=True)
    cfg = pkg.ViTConfig(**TINY, num_labels=14)
    m, info = pkg.ViTForImageClassification.from_pretrained(base.state_dict(), config=cfg, output_loading_info=True)
    assert sorted(info["missing_keys"]) == ["classifier.bias", "classifier.weight"]
    assert all(k.startswith("vit.pooler") for k in info["unexpected_keys"]) and info["unexpected_keys"]
    bsd = base.state_dict()
    for k, v in m.state_dict().items():
        if k.startswith("vit."):
            assert torch.equal(v, bsd[k[4:]]), k
