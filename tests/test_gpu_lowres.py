"""GPU parity of the fused low-res path: bacs_pixel_loss_lowres evaluates the network's final
bilinear up-sample (networks/deeplab_v3.py:154-160, align_corners=False), the per-pixel loss and the adjoint of the
up-sample in one kernel.  Oracle: the up-sample in fp32 followed by the oracle's loss, gradients by autograd.
Tolerances: 1e-5 relative on losses, 2e-5 of max |grad| on fp32 gradients (sums of ~256 fp32 atomics per low-res
cell), 1 ulp of the storage type for 16-bit gradients; arg-max and distill mask bit-exact away from fp32 near-ties."""
import pytest
import torch

from oracle import bacs_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def ops():
    from bacs_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def synth():
    from bacs_b200 import synth as _synth
    return _synth


def close(got, want, rtol=RTOL, atol=None, what=""):
    got = torch.as_tensor(got).detach().double().cpu()
    want = torch.as_tensor(want).detach().double().cpu()
    if atol is None:
        atol = rtol * max(1e-30, float(want.abs().max()))
    err = float((got - want).abs().max()) if got.numel() else 0.0
    assert torch.allclose(got, want, rtol=rtol, atol=atol), "%s max abs err %.3e (atol %.3e)" % (what, err, atol)


def _sem(cfg, stride, seed, dtype, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(cfg.B, cfg.K, cfg.H // stride, cfg.W // stride, generator=g) * scale).to(dtype)


def _check_preds(preds, up):
    want = O.argmax_first(up)
    top2 = up.topk(2, dim=1)[0]
    gap = top2[:, 0] - top2[:, 1]
    diff = preds.cpu() != want
    assert int((diff & (gap > 1e-5)).sum()) == 0, "arg-max differs away from near-ties"
    assert int(diff.sum()) <= max(2, preds.numel() // 100000)


def _seen_z(inp):
    return O.seen_logits_lowres(inp.pen.float(), inp.protos, inp.head_w, inp.head_b)


@pytest.mark.parametrize("name,dtype,ukd,stride", [
    ("tiny", torch.float32, True, 16), ("small", torch.float32, True, 16), ("small", torch.bfloat16, False, 16),
    ("row512", torch.float32, True, 16), ("row512", torch.bfloat16, True, 16), ("row512", torch.float16, False, 16),
    ("row1024", torch.float32, True, 16), ("row512_k40", torch.float32, True, 16), ("row512_k7", torch.bfloat16, True, 16),
    ("voc15-1_cpu", torch.float32, True, 16), ("small", torch.float32, True, 8), ("row512_t11", torch.float32, False, 8)])
def test_lowres_training_step(ops, synth, name, dtype, ukd, stride):
    """weighted CE + focal term of one seen head + distill mask + arg-max + d/d sem_logits + d/d z in one launch"""
    from bacs_b200 import _cabi
    cfg = synth.CONFIGS[name]
    if name == "voc15-1_cpu":     # 528 x 528 -> 33 x 33 (not a power of two), two images are enough
        cfg = synth.StepConfig("lr528", B=2, K=21, old_cl=20, T=6, H=528, W=528, D=32, A=16)
    inp = synth.make_step_inputs(cfg, seed=5, dtype=torch.float32)
    g = torch.Generator().manual_seed(3)
    mask = synth.make_labels(cfg, g, classes=list(range(1, cfg.K)))
    mask[0, 9, 40:47] = cfg.K + 3                       # outside [0,K), not ignore: counted, treated as ignore
    clean = torch.where((mask >= cfg.K) & (mask != 255), torch.full_like(mask, 255), mask)
    t = cfg.T - 1
    z = _seen_z(inp).clone().requires_grad_(True)
    upz = O.bilinear_upsample(z, (cfg.H, cfg.W), True)
    smax = torch.sigmoid(upz).max(1)[0].detach()
    sem = _sem(cfg, stride, 11, dtype, 2.0)
    x = sem.float().clone().requires_grad_(True)
    up = O.upsample_sem_logits(x, (cfg.H, cfg.W))
    want = O.weighted_ce(up, clean, smax, cfg.old_cl, 2.0, 0.5, ukd)
    want.backward()
    kept = int((clean != 255).sum())
    want_f = O.focal_seen_loss(upz[:, t:t + 1], clean, 2.0, 0.25)
    want_f.backward()
    scale = 1024.0 if dtype == torch.float16 else 1.0
    out = ops.pixel_loss(sem.cuda(), mask.cuda(), _cabi.PIX_WEIGHTED_CE, want_grad=True, z=z.detach().cuda(),
                         want_distill_mask=True, old_cl=cfg.old_cl, ukd=ukd, grad_scale=scale, focal_head=t,
                         focal_alpha=0.25, lowres=True)
    assert out["variant"] == 3
    assert tuple(out["dlogits"].shape) == tuple(sem.shape) and out["dlogits"].dtype == dtype
    N = cfg.B * cfg.H * cfg.W
    acc = out["acc"].cpu()
    close(acc[_cabi.ACC_LOSS] / N, want, what="loss")
    close(acc[_cabi.ACC_FOCAL] / kept, want_f, what="focal loss")
    close(out["gz"] / kept, z.grad[:, t], atol=2e-5 * float(z.grad.abs().max()), what="gz")
    _check_preds(out["preds"], up.detach())
    want_g = (x.grad * scale).to(dtype).float()
    tol = 2e-5 if dtype == torch.float32 else 2.0 ** (-7 if dtype == torch.bfloat16 else -10)
    close(out["dlogits"].float(), want_g, atol=tol * float(want_g.abs().max()), what="d sem_logits")
    want_m = (clean == 0) & (smax > 0.5)
    diff = out["distill_mask"].cpu().bool() != want_m
    assert int((diff & ((smax - 0.5).abs() > 1e-6)).sum()) == 0
    assert int(diff.sum()) <= 2
    assert int(acc[_cabi.ACC_KEPT]) == kept and int(acc[_cabi.ACC_VALID]) == kept
    assert int(acc[_cabi.ACC_BG]) == int((clean == 0).sum())
    assert int(acc[_cabi.ACC_INVALID]) == 7
    assert int(acc[_cabi.ACC_DISTILL_PIX]) == int(out["distill_mask"].sum())
    # evaluation: same loss, no gradient
    out2 = ops.pixel_loss(sem.cuda(), mask.cuda(), _cabi.PIX_WEIGHTED_CE, want_grad=False, z=z.detach().cuda(),
                          old_cl=cfg.old_cl, ukd=ukd, lowres=True)
    close(out2["acc"][_cabi.ACC_LOSS] / N, want, what="loss (no grad)")
    assert out2["dlogits"] is None and torch.equal(out2["preds"], out["preds"])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_lowres_ce_modes(ops, synth, dtype):
    """plain / class-weighted CE (task 0, evaluation, dark++), MiB unbiased CE and the per-image score"""
    from bacs_b200 import _cabi
    cfg = synth.CONFIGS["small"]
    g = torch.Generator().manual_seed(2)
    mask = synth.make_labels(cfg, g, classes=list(range(1, cfg.K)))
    sem = _sem(cfg, 16, 4, dtype, 3.0)
    w = torch.zeros(cfg.K)
    w[1:cfg.old_cl] = 1
    tol = 2e-5 if dtype == torch.float32 else 2.0 ** -7
    for weight in (None, w):
        x = sem.float().clone().requires_grad_(True)
        up = O.upsample_sem_logits(x, (cfg.H, cfg.W))
        want = O.cross_entropy(up, mask, weight)
        want.backward()
        out = ops.pixel_loss(sem.cuda(), mask.cuda(), _cabi.PIX_CE, want_grad=True,
                             class_w=None if weight is None else weight.cuda(), grad_scale=0.2, lowres=True)
        acc = out["acc"]
        close(acc[_cabi.ACC_LOSS] / acc[_cabi.ACC_WSUM], want, what="ce")
        wg = (0.2 * x.grad).to(dtype).float()
        close(out["dlogits"].float(), wg, atol=tol * float(wg.abs().max()), what="ce grad")
        _check_preds(out["preds"], up.detach())
    up = O.upsample_sem_logits(sem, (cfg.H, cfg.W))
    out = ops.pixel_loss(sem.cuda(), mask.cuda(), _cabi.PIX_CE, want_grad=False, lowres=True)
    close(out["acc"][_cabi.ACC_LOSS] / out["acc"][_cabi.ACC_WSUM], O.cross_entropy(up, mask))
    x = sem.float().clone().requires_grad_(True)
    want = O.unbiased_ce(O.upsample_sem_logits(x, (cfg.H, cfg.W)), mask, cfg.old_cl)
    want.backward()
    out = ops.pixel_loss(sem.cuda(), mask.cuda(), _cabi.PIX_UNBIASED_CE, want_grad=True, old_cl=cfg.old_cl, lowres=True)
    close(out["acc"][_cabi.ACC_LOSS] / out["acc"][_cabi.ACC_WSUM], want, what="uce")
    wg = x.grad.to(dtype).float()
    close(out["dlogits"].float(), wg, atol=tol * float(wg.abs().max()), what="uce grad")
    w2 = torch.ones(cfg.K)
    w2[0] = 0
    out = ops.pixel_loss(sem.cuda(), mask.cuda(), _cabi.PIX_SCORE, want_grad=False, class_w=w2.cuda(),
                         want_score=True, want_preds=False, lowres=True)
    close(out["score"], O.cross_entropy_per_image_score(up, mask, w2), what="score")


def test_lowres_matches_fullres_kernel(ops, synth):
    """the fused path against the full-resolution kernel fed torch's own up-sample (the reference's data flow)"""
    from bacs_b200 import _cabi
    cfg = synth.CONFIGS["row512"]
    inp = synth.make_step_inputs(cfg, seed=7, dtype=torch.float32)
    g = torch.Generator().manual_seed(8)
    mask = synth.make_labels(cfg, g, classes=list(range(1, cfg.K))).cuda()
    z = _seen_z(inp).cuda()
    sem = _sem(cfg, 16, 3, torch.float32, 2.0).cuda().requires_grad_(True)
    full = torch.nn.functional.interpolate(sem, size=(cfg.H, cfg.W), mode="bilinear", align_corners=False)
    ref = ops.pixel_loss(full.detach(), mask, _cabi.PIX_WEIGHTED_CE, want_grad=True, z=z, want_distill_mask=True,
                         old_cl=cfg.old_cl)
    full.backward(ref["dlogits"])
    out = ops.pixel_loss(sem.detach(), mask, _cabi.PIX_WEIGHTED_CE, want_grad=True, z=z, want_distill_mask=True,
                         old_cl=cfg.old_cl, lowres=True)
    close(out["acc"][_cabi.ACC_LOSS], ref["acc"][_cabi.ACC_LOSS], what="loss")
    close(out["dlogits"], sem.grad, atol=2e-5 * float(sem.grad.abs().max()), what="d sem_logits")
    assert torch.equal(out["distill_mask"], ref["distill_mask"])
    assert int((out["preds"] != ref["preds"]).sum()) <= 2


def test_lowres_rejects_unsupported_geometry(ops):
    from bacs_b200 import _cabi
    sem = torch.randn(1, 5, 8, 8).cuda()
    lab = torch.zeros(1, 32, 32, dtype=torch.int64).cuda()      # x4: not a network stride of the reference
    with pytest.raises(_cabi.BacsError):
        ops.pixel_loss(sem, lab, _cabi.PIX_CE, want_grad=False, lowres=True)


@pytest.mark.parametrize("name,dtype", [("tiny", torch.float32), ("small", torch.float32), ("row512", torch.float32),
                                        ("row512", torch.bfloat16)])
def test_full_step_with_fused_logit_upsample(synth, name, dtype):
    """BACSLoss(fused_logit_upsample=True): the network hands over sem_logits (return_sem_logits=True) and the whole
    step -- prototypes, seen heads, weighted CE, focal, distill, DER replay, dark++ -- matches the oracle's step fed
    the up-sampled logits; gradients arrive at the low-res logits."""
    cfg = synth.CONFIGS[name]
    inp = synth.make_step_inputs(cfg, seed=11, dtype=dtype)
    g = torch.Generator().manual_seed(21)
    sem = (torch.randn(cfg.B, cfg.K, cfg.h, cfg.w, generator=g) * 2).to(dtype)
    rsem = (torch.randn(cfg.Br, cfg.K, cfg.h, cfg.w, generator=g) * 2).to(dtype)
    # ---- oracle: the reference's data flow (up-sample, then the step) ----
    leaf = lambda t: t.float().clone().requires_grad_(True)
    sem_o, rsem_o, na, hw, hb = leaf(sem), leaf(rsem), leaf(inp.new_att), leaf(inp.head_w), leaf(inp.head_b)
    rp = inp.replay
    rdl = leaf(rp["sem_logits"])
    rp2 = dict(rp, logits=O.upsample_sem_logits(rsem_o, (cfg.H, cfg.W)), sem_logits=rdl, pen=rp["pen"].float(),
               n_classes=rp["n_classes"].numpy(), memory_logits=rp["memory_logits"].float())
    want = O.bacs_step(O.upsample_sem_logits(sem_o, (cfg.H, cfg.W)), inp.pen.float(), inp.old_att.float(), na, inp.mask,
                       inp.protos, inp.counts, hw, hb, initial_classes=cfg.initial_classes, increment=cfg.increment,
                       old_cl=cfg.old_cl, task_num=cfg.T - 1, first_task=False, epoch=3, max_epochs=30, replay=rp2,
                       nb_current_classes=cfg.K, proto_mode="exact")
    want["loss"].backward()
    # ---- the product path ----
    loss_fn, net, batch, leaves = synth.build_bacs_step(cfg, inp, fused_logit_upsample=True)
    sem_g = sem.clone().cuda().requires_grad_(True)
    rsem_g = rsem.clone().cuda().requires_grad_(True)
    net.register_sem(batch["main"][0], sem_g)
    net.register_sem(batch["buffer"][0], rsem_g)
    loss, preds = loss_fn.compute_loss(batch, net, train=True)
    loss.backward()
    close(loss, want["loss"], what="loss")
    up = O.upsample_sem_logits(sem, (cfg.H, cfg.W))
    _check_preds(preds, up)
    close(loss_fn.prototypes, want["protos"], what="prototypes")
    tol = 2e-5 if dtype == torch.float32 else 2.0 ** -7
    for got, wnt, what in ((sem_g.grad, sem_o.grad, "d sem_logits"), (rsem_g.grad, rsem_o.grad, "d replay sem_logits"),
                           (leaves["replay_sem"].grad, rdl.grad, "d DER logits")):
        w = wnt.to(dtype).float()
        close(got.float(), w, atol=tol * float(w.abs().max()), what=what)
    w = na.grad.to(dtype).float()
    close(leaves["new_att"].grad.float(), w, atol=3 * tol * float(w.abs().max()), what="d new_att")
    assert leaves["logits"].grad is None          # the full-resolution logits were never touched
    t = cfg.T - 1
    close(leaves["head_w"].grad.reshape(-1), hw.grad[t], atol=3e-5 * float(hw.grad[t].abs().max()), what="dhead_w")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_lowres_confident_logits(ops, synth, dtype):
    """|logit| up to ~80 (far beyond random init): finite, and loss / gradient still match the oracle"""
    from bacs_b200 import _cabi
    cfg = synth.CONFIGS["row512"]
    inp = synth.make_step_inputs(cfg, seed=9, dtype=torch.float32)
    g = torch.Generator().manual_seed(5)
    mask = synth.make_labels(cfg, g, classes=list(range(1, cfg.K)))
    z = _seen_z(inp)
    smax = torch.sigmoid(O.bilinear_upsample(z, (cfg.H, cfg.W), True)).max(1)[0]
    sem = _sem(cfg, 16, 13, dtype, 20.0)
    x = sem.float().clone().requires_grad_(True)
    up = O.upsample_sem_logits(x, (cfg.H, cfg.W))
    want = O.weighted_ce(up, mask, smax, cfg.old_cl, 2.0, 0.5, True)
    want.backward()
    out = ops.pixel_loss(sem.cuda(), mask.cuda(), _cabi.PIX_WEIGHTED_CE, want_grad=True, z=z.cuda(),
                         want_distill_mask=True, old_cl=cfg.old_cl, ukd=True, lowres=True)
    assert bool(torch.isfinite(out["acc"]).all()) and bool(torch.isfinite(out["dlogits"].float()).all())
    N = cfg.B * cfg.H * cfg.W
    close(out["acc"][_cabi.ACC_LOSS] / N, want, rtol=2e-5, what="loss")
    want_g = x.grad.to(dtype).float()
    tol = 3e-5 if dtype == torch.float32 else 2.0 ** -7
    close(out["dlogits"].float(), want_g, atol=tol * float(want_g.abs().max()), what="d sem_logits")
    _check_preds(out["preds"], up.detach())
