"""GPU parity of the stand-alone module mirrors a maintainer imports directly (same class names as the reference):
networks.bg_detector.BgDetector / classification_head (networks/bg_detector.py:5-165) and
training.loss_utils.WeightedCrossEntropy / UnbiasedCrossEntropy (training/loss_utils.py:492-588), forward and
backward through autograd, against the CPU oracle.  fp32: 1e-5 relative (3e-5 for the head gradients)."""
import pytest
import torch

from oracle import bacs_oracle as O

pytestmark = pytest.mark.gpu


def close(got, want, rtol=1e-5, atol=None, what=""):
    got = torch.as_tensor(got).detach().double().cpu()
    want = torch.as_tensor(want).detach().double().cpu()
    if atol is None:
        atol = rtol * max(1e-30, float(want.abs().max()))
    err = float((got - want).abs().max()) if got.numel() else 0.0
    assert torch.allclose(got, want, rtol=rtol, atol=atol), "%s max abs err %.3e (atol %.3e)" % (what, err, atol)


def _detector(synth, cfg, inp):
    from bacs_b200.networks.bg_detector import BgDetector
    bg = BgDetector(cfg.D * 4)
    heads = torch.nn.ModuleList([bg.get_classification_head(1) for _ in range(cfg.T)])
    with torch.no_grad():
        for t, head in enumerate(heads):
            head.conv.weight.copy_(inp.head_w[t].view(1, cfg.D, 1, 1))
            head.conv.bias.copy_(inp.head_b[t].view(1))
    bg.seen_not_seen_clf = heads
    return bg.cuda()


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_bg_detector_module(name):
    from bacs_b200 import synth
    cfg = synth.CONFIGS[name]
    inp = synth.make_step_inputs(cfg, seed=8)
    bg = _detector(synth, cfg, inp)
    assert bg.get_penultimate_layer_dim() == cfg.D
    pen, protos = inp.pen.cuda(), inp.protos.cuda()
    # all heads, probabilities at full resolution (get_seen_probs, bg_detector.py:141-165)
    close(bg.get_seen_probs(pen, protos, bg_detect=True), O.seen_probs(inp.pen, inp.protos, inp.head_w, inp.head_b),
          what="seen probs")
    z_all = O.seen_logits_lowres(inp.pen, inp.protos, inp.head_w, inp.head_b)
    close(bg.forward_seen_before(pen, protos), O.bilinear_upsample(z_all, (cfg.H, cfg.W), True), what="seen logits")
    # one head with autograd (get_seen_map_task, bg_detector.py:100-117): gradients to the head and to the features
    t = cfg.T - 1
    g = torch.Generator().manual_seed(4)
    upstream = torch.randn(cfg.B, 1, cfg.H, cfg.W, generator=g)
    xo, wo, bo = (v.clone().requires_grad_(True) for v in (inp.pen, inp.head_w, inp.head_b))
    want = O.bilinear_upsample(O.seen_logits_lowres(xo, inp.protos, wo, bo)[:, t:t + 1], (cfg.H, cfg.W), True)
    want.backward(upstream)
    for stop in (False, True):
        bg.set_stop_gradients(stop)
        head = bg.seen_not_seen_clf[t]
        head.conv.weight.grad = head.conv.bias.grad = None
        xg = inp.pen.clone().cuda().requires_grad_(True)
        got = bg.get_seen_map_task(xg, protos, t)
        assert tuple(got.shape) == (cfg.B, 1, cfg.H, cfg.W)
        close(got, want, what="seen map")
        got.backward(upstream.cuda())
        close(head.conv.weight.grad.reshape(-1), wo.grad[t], atol=3e-5 * float(wo.grad[t].abs().max()), what="dweight")
        close(head.conv.bias.grad.reshape(()), bo.grad[t], atol=3e-5 * float(bo.grad[t].abs()), what="dbias")
        if stop:    # stop_gradients detaches the features (bg_detector.py:28-30)
            assert xg.grad is None or float(xg.grad.abs().max()) == 0.0
        else:
            close(xg.grad, xo.grad, atol=3e-5 * float(xo.grad.abs().max()), what="dfeatures")
    with torch.no_grad():       # inference path: the up-sample runs in the library too
        close(bg.get_seen_map_task(pen, protos, t), want, what="seen map (no grad)")


@pytest.mark.parametrize("ukd", [True, False])
def test_weighted_ce_module(ukd):
    """WeightedCrossEntropy(...)(inputs, targets, seen_not_seen_probs [B,T,H,W], task_num) as base_loss.py:236 calls it"""
    from bacs_b200 import synth
    from bacs_b200.training.loss_utils import WeightedCrossEntropy
    cfg = synth.CONFIGS["small"]
    inp = synth.make_step_inputs(cfg, seed=9)
    probs = O.seen_probs(inp.pen, inp.protos, inp.head_w, inp.head_b)
    x = inp.logits.clone().requires_grad_(True)
    want = O.weighted_ce(x, inp.mask, probs.max(1)[0], cfg.old_cl, 2.0, 0.5, ukd)
    want.backward()
    mod = WeightedCrossEntropy(gamma=2, old_cl=cfg.old_cl, threshold=0.5, ukd=ukd)
    xg = inp.logits.clone().cuda().requires_grad_(True)
    got = mod(xg, inp.mask.cuda(), probs.cuda(), cfg.T - 1)
    (3.0 * got).backward()
    close(got, want, what="weighted ce")
    close(xg.grad, 3.0 * x.grad, atol=1e-5 * float(3.0 * x.grad.abs().max()), what="weighted ce grad")


def test_unbiased_ce_module():
    from bacs_b200 import synth
    from bacs_b200.training.loss_utils import UnbiasedCrossEntropy
    cfg = synth.CONFIGS["small"]
    inp = synth.make_step_inputs(cfg, seed=10)
    x = inp.logits.clone().requires_grad_(True)
    want = O.unbiased_ce(x, inp.mask, cfg.old_cl)
    want.backward()
    xg = inp.logits.clone().cuda().requires_grad_(True)
    got = UnbiasedCrossEntropy(old_cl=cfg.old_cl)(xg, inp.mask.cuda())
    got.backward()
    close(got, want, what="unbiased ce")
    close(xg.grad, x.grad, atol=1e-5 * float(x.grad.abs().max()), what="unbiased ce grad")
    with pytest.raises(NotImplementedError):
        UnbiasedCrossEntropy(old_cl=3, reduction="none")
