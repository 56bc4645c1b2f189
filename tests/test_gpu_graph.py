"""CUDA-graph replay of the launch plans (chest_x_ray_vit_b200.graph) against the eager module path: same kernels in
the same order, so forward results are bit-identical and training trajectories agree to atomics-order rounding."""
import pytest
import torch

from oracle import vit_oracle as O

pytestmark = pytest.mark.gpu


def _model(cfg, params):
    import chest_x_ray_vit_b200 as pkg
    m = pkg.ViTForImageClassification(pkg.ViTConfig(image_size=cfg.image_size, hidden_size=cfg.hidden_size,
                                                    num_hidden_layers=cfg.num_hidden_layers,
                                                    num_attention_heads=cfg.num_attention_heads,
                                                    intermediate_size=cfg.intermediate_size, num_labels=cfg.num_labels))
    m.load_state_dict(params)
    return m.cuda()


@pytest.mark.parametrize("dtype", ["u8", "f32"])
def test_graphed_forward_equals_eager(dtype):
    import chest_x_ray_vit_b200 as pkg
    cfg = O.TINY
    m = _model(cfg, O.init_params(cfg, 0, 123)).eval()
    g = torch.Generator().manual_seed(5)
    xs = []
    for _ in range(3):
        x8, _ = O.synth_inputs(cfg, 4, g)
        xs.append(x8[:, 0].cuda() if dtype == "u8" else O.normalize_gray(x8).cuda())
    with torch.no_grad():
        ref = [m(pixel_values=x).logits.clone() for x in xs]
        gf = pkg.graph.GraphedForward(m, xs[0])
        assert gf.launches_per_replay > 0
        for x, r in zip(xs, ref):
            assert torch.equal(gf(x), r)
        with pytest.raises(ValueError):
            gf(xs[0][:2])
        # weights moved: the replay must see the refreshed bf16 shadow
        m.load_state_dict(O.init_params(cfg, 1, 124))
        assert torch.equal(gf(xs[1]), m(pixel_values=xs[1]).logits)


def test_graphed_train_step_follows_eager_trajectory():
    import chest_x_ray_vit_b200 as pkg
    cfg = O.TINY
    g = torch.Generator().manual_seed(6)
    batches = []
    for _ in range(3):
        x8, y = O.synth_inputs(cfg, 4, g)
        batches.append((x8[:, 0].cuda(), y.cuda()))
    ma, mb = _model(cfg, O.init_params(cfg, 0, 123)).train(), _model(cfg, O.init_params(cfg, 0, 123)).train()
    oa = pkg.VitkAdamW(ma, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
    ob = pkg.VitkAdamW(mb, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
    gstep = pkg.graph.GraphedTrainStep(mb, ob)
    la, lb = [], []
    for it in range(7):                        # graph side: capture on the first call, then replays
        x, y = batches[it % 3]
        out = ma(pixel_values=x, labels=y)
        out.loss.backward()
        oa.step()
        oa.zero_grad(set_to_none=True)
        la.append(out.loss.item())
        lb.append(float(gstep(x, y)))
    assert gstep.replays == 7 and gstep.kernel_launches > 0 and ob._step == oa._step == 7
    assert abs(la[0] - lb[0]) < 1e-6           # identical kernels on identical weights (the loss mean is an fp32 atomic sum over images)
    assert max(abs(a - b) for a, b in zip(la, lb)) < 2e-4, (la, lb)
    assert la[-1] < la[0]

    def close(tag):
        # gradients are reproducible only up to the order of fp32 atomics, and Adam turns a sign-uncertain gradient
        # element into a ±lr difference per step: compare statistically
        d = torch.cat([(pa.detach() - pb.detach()).abs().flatten() for pa, pb in zip(ma.parameters(), mb.parameters())])
        frac = (d > 1e-4).float().mean().item()
        print(f"{tag}: {100 * frac:.3f}% of elements differ by more than lr/10, mean |Δ| {d.mean().item():.2e}")
        assert frac < 0.03 and d.mean().item() < 3e-5, (tag, frac, d.mean().item())
    close("graph vs eager after 7 steps")
    # optimizer state carried by the graph is the optimizer's own: an eager step continues from it
    x, y = batches[0]
    for m, o in ((ma, oa), (mb, ob)):
        m(pixel_values=x, labels=y).loss.backward()
        o.step()
        o.zero_grad(set_to_none=True)
    close("after one more eager step on both")
