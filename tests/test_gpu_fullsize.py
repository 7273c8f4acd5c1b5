"""Parity at BASELINE.json's full sizes (VOC 15-1 B=24 512x512 K=21 bf16; Cityscapes 1024x2048 evaluation).

The CPU oracle needs minutes at these sizes, so the checks are (a) the oracle's own torch code evaluated on
device tensors (an independent implementation of the same formulas: different kernels, fp32, no fusion),
(b) per-image parity against the CPU oracle on sampled images, and (c) size-independent properties of the
domain: exact power-of-two homogeneity of the distill norm, additivity over images, checksums of counts."""
import numpy as np
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


@pytest.fixture(scope="module")
def headline():
    from bacs_b200 import synth
    cfg = synth.CONFIGS["voc15-1_b24"]
    inp = synth.make_step_inputs(cfg, seed=3, dtype=torch.bfloat16, device="cuda")
    return cfg, inp


def test_pixel_kernel_headline_size(headline):
    from bacs_b200 import _cabi, ops
    cfg, inp = headline
    z = ops.seen_logits(inp.pen, inp.protos, inp.head_w, inp.head_b)
    t = cfg.T - 1
    out = ops.pixel_loss(inp.logits, inp.mask, _cabi.PIX_WEIGHTED_CE, want_grad=True, z=z, want_distill_mask=True,
                         old_cl=cfg.old_cl, ukd=True, focal_head=t)
    assert out["variant"] == 2
    # (a) the oracle's formulas on device tensors, whole batch
    zr = z.clone().requires_grad_(True)
    up = O.bilinear_upsample(zr, (cfg.H, cfg.W), True)
    smax = torch.sigmoid(up).max(1)[0].detach()
    x = inp.logits.float().requires_grad_(True)
    want = O.weighted_ce(x, inp.mask, smax, cfg.old_cl, 2.0, 0.5, True)
    want.backward()
    N = cfg.pixels
    acc = out["acc"].cpu()
    close(acc[_cabi.ACC_LOSS] / N, want, what="loss")
    want_g = x.grad.to(torch.bfloat16).float()
    close(out["dlogits"].float(), want_g, atol=2.0 ** -7 * float(want_g.abs().max()), what="dlogits")
    del x, want_g
    kept = int((inp.mask != 255).sum())
    want_f = O.focal_seen_loss(up[:, t:t + 1], inp.mask, 2.0, None)
    want_f.backward()
    close(acc[_cabi.ACC_FOCAL] / kept, want_f, what="focal loss")
    close(out["gz"] / kept, zr.grad[:, t], atol=3e-5 * float(zr.grad.abs().max()), what="gz")
    # first-index arg-max, exact
    lf = inp.logits.float()
    first = torch.where(lf == lf.max(1, keepdim=True)[0], torch.arange(cfg.K, device="cuda").view(1, -1, 1, 1),
                        torch.full((), cfg.K, device="cuda")).min(1)[0]
    assert torch.equal(out["preds"], first)
    want_m = (inp.mask == 0) & (smax > 0.5)
    diff = out["distill_mask"].bool() != want_m
    assert int((diff & ((smax - 0.5).abs() > 1e-6)).sum()) == 0
    assert int(diff.sum()) <= 8
    # (c) checksums
    assert int(acc[_cabi.ACC_KEPT]) == kept
    assert int(acc[_cabi.ACC_BG]) == int((inp.mask == 0).sum())
    assert int(acc[_cabi.ACC_INVALID]) == 0
    assert int(acc[_cabi.ACC_DISTILL_PIX]) == int(out["distill_mask"].sum())
    # (b) CPU oracle on two sampled images: the gradient of an image only depends on that image (x 1/B)
    for b in (0, cfg.B - 1):
        xb = inp.logits[b:b + 1].float().cpu().requires_grad_(True)
        wb = O.weighted_ce(xb, inp.mask[b:b + 1].cpu(), smax[b:b + 1].cpu(), cfg.old_cl, 2.0, 0.5, True)
        wb.backward()
        wg = (xb.grad / cfg.B).to(torch.bfloat16).float()
        close(out["dlogits"][b:b + 1].float(), wg, atol=2.0 ** -7 * float(wg.abs().max()), what="dlogits[%d]" % b)


def test_distill_headline_size_properties(headline):
    from bacs_b200 import ops
    cfg, inp = headline
    g = torch.Generator().manual_seed(1)
    m = ((inp.mask.cpu() == 0) & (torch.rand(cfg.B, cfg.H, cfg.W, generator=g) > 0.4)).to(torch.uint8).cuda()
    tot, dnew = ops.teacher_distill(inp.old_att, inp.new_att, m, (cfg.H, cfg.W), 1.0, True)
    # identical maps: exactly zero loss and gradient (start of every task)
    t0, d0 = ops.teacher_distill(inp.old_att, inp.old_att, m, (cfg.H, cfg.W), 1.0, True)
    assert float(t0) == 0.0 and float(d0.float().abs().max()) == 0.0
    # homogeneity of degree 2, exact for a power-of-two factor
    t2, d2 = ops.teacher_distill(inp.old_att * 2, inp.new_att * 2, m, (cfg.H, cfg.W), 1.0, True)
    assert float(t2) == 4.0 * float(tot)
    assert torch.equal(d2.float(), 2.0 * dnew.float())
    # additivity over images + CPU oracle on two of them
    parts = [ops.teacher_distill(inp.old_att[b:b + 1], inp.new_att[b:b + 1], m[b:b + 1], (cfg.H, cfg.W), 1.0, True)
             for b in range(cfg.B)]
    close(sum(float(p[0]) for p in parts), float(tot), rtol=1e-6, what="sum over images")
    assert torch.equal(torch.cat([p[1] for p in parts]), dnew)
    for b in (0, cfg.B - 1):
        new = inp.new_att[b:b + 1].float().cpu().requires_grad_(True)
        lab = torch.where(m[b:b + 1].cpu().bool(), 0, 1)
        want = O.teacher_distill(inp.old_att[b:b + 1].float().cpu(), new, lab, None, lkd=1.0)
        want.backward()
        close(float(parts[b][0]) / (cfg.A * cfg.H), want, what="distill loss image %d" % b)
        wg = (new.grad * (cfg.A * cfg.H)).to(torch.bfloat16).float()
        close(parts[b][1].float().cpu(), wg, atol=3 * 2.0 ** -8 * float(wg.abs().max()), what="distill grad image %d" % b)


def test_prototype_sums_headline_size(headline):
    from bacs_b200 import ops
    cfg, inp = headline
    lut = torch.from_numpy(O.class_task_lut(cfg.initial_classes, cfg.increment).astype(np.int32)).cuda()
    task, rank, n_bt, _ = ops.label_downsample_task(inp.mask, cfg.h, cfg.w, lut, cfg.T)
    feats = inp.pen.double().permute(1, 0, 2, 3).reshape(cfg.D, -1)             # [D, B*h*w]
    tk = task.reshape(-1).long()
    for mode in (0, 1):
        sums, counts = ops.proto_accumulate(inp.pen, task, rank, n_bt, cfg.T, mode=mode)
        assert torch.equal(counts.long(), torch.bincount(tk[tk >= 0], minlength=cfg.T))
        assert torch.equal(n_bt.sum(0).long(), counts.long())
        # the grand total per task is the same whichever way the D x N view cuts the flat sequence (mode 0)
        tot = torch.stack([feats[:, tk == t].sum() for t in range(cfg.T)])
        close(sums.sum(1), tot, rtol=1e-5, atol=1e-5 * float(feats.abs().sum()) / cfg.D, what="task totals")
        if mode == 1:
            per = torch.stack([feats[:, tk == t].sum(1) for t in range(cfg.T)])
            close(sums, per, atol=1e-5 * float(per.abs().max()), what="per-channel sums")


def test_confusion_matrix_cityscapes_eval_size():
    from bacs_b200 import ops
    K, B, H, W = 19, 4, 1024, 2048
    g = torch.Generator(device="cuda").manual_seed(4)
    preds = torch.randint(0, K, (B, H, W), generator=g, device="cuda")
    target = torch.randint(0, K, (B, H, W), generator=g, device="cuda")
    target[torch.rand(B, H, W, generator=g, device="cuda") < 0.1] = 255
    cm = torch.zeros(K, K, dtype=torch.int64, device="cuda")
    for _ in range(2):                                           # accumulates
        ops.confmat_accumulate(preds, target, K, cm)
    valid = target != 255
    want = torch.bincount(target[valid] * K + preds[valid], minlength=K * K).view(K, K) * 2
    assert torch.equal(cm, want)
    assert int(cm.sum()) == 2 * int(valid.sum())


def _device_oracle_pixel_check(cfg, B, dtype, lowres, seed=5):
    """Whole-batch parity of the fused pixel kernel (full-resolution or low-res-logit variant) against the oracle's
    formulas evaluated on device tensors, plus checksums; shapes of BASELINE.json's other configs."""
    from bacs_b200 import _cabi, ops, synth
    g = torch.Generator(device="cuda").manual_seed(seed)
    mask = synth.make_labels(cfg, torch.Generator().manual_seed(seed), classes=list(range(1, cfg.K)), B=B).cuda()
    z = torch.randn(B, cfg.T, cfg.h, cfg.w, device="cuda", generator=g)
    t = cfg.T - 1
    if lowres:
        src = (torch.randn(B, cfg.K, cfg.h, cfg.w, device="cuda", generator=g) * 2).to(dtype)
    else:
        src = torch.randn(B, cfg.K, cfg.H, cfg.W, device="cuda", generator=g).to(dtype)
    out = ops.pixel_loss(src, mask, _cabi.PIX_WEIGHTED_CE, want_grad=True, z=z, want_distill_mask=True,
                         old_cl=cfg.old_cl, ukd=True, focal_head=t, lowres=lowres)
    zr = z.clone().requires_grad_(True)
    up_z = O.bilinear_upsample(zr, (cfg.H, cfg.W), True)
    smax = torch.sigmoid(up_z).max(1)[0].detach()
    x = src.float().requires_grad_(True)
    full = O.upsample_sem_logits(x, (cfg.H, cfg.W)) if lowres else x
    want = O.weighted_ce(full, mask, smax, cfg.old_cl, 2.0, 0.5, True)
    want.backward()
    N = B * cfg.H * cfg.W
    acc = out["acc"].cpu()
    close(acc[_cabi.ACC_LOSS] / N, want, what="loss")
    want_g = x.grad.to(dtype).float()
    tol = 2e-5 if dtype == torch.float32 else 2.0 ** -7
    close(out["dlogits"].float(), want_g, atol=tol * float(want_g.abs().max()), what="gradient")
    kept = int((mask != 255).sum())
    want_f = O.focal_seen_loss(up_z[:, t:t + 1], mask, 2.0, None)
    want_f.backward()
    close(acc[_cabi.ACC_FOCAL] / kept, want_f, what="focal loss")
    close(out["gz"] / kept, zr.grad[:, t], atol=3e-5 * float(zr.grad.abs().max()), what="gz")
    lf = full.detach()
    top2 = lf.topk(2, dim=1)[0]
    first = torch.where(lf == top2[:, :1], torch.arange(cfg.K, device="cuda").view(1, -1, 1, 1),
                        torch.full((), cfg.K, device="cuda")).min(1)[0]
    diff = out["preds"] != first
    if lowres:      # the kernel's fp32 up-sample differs from torch's in the last bit: near-ties may flip
        assert int((diff & ((top2[:, 0] - top2[:, 1]) > 1e-5)).sum()) == 0
        assert int(diff.sum()) <= max(2, N // 100000)
    else:
        assert int(diff.sum()) == 0
    want_m = (mask == 0) & (smax > 0.5)
    dm = out["distill_mask"].bool() != want_m
    assert int((dm & ((smax - 0.5).abs() > 1e-6)).sum()) == 0 and int(dm.sum()) <= 8
    assert int(acc[_cabi.ACC_KEPT]) == kept and int(acc[_cabi.ACC_BG]) == int((mask == 0).sum())
    assert int(acc[_cabi.ACC_DISTILL_PIX]) == int(out["distill_mask"].sum())
    return out["variant"]


@pytest.mark.parametrize("name,B,dtype,variant", [
    ("voc10-1_der", 24, torch.bfloat16, 2),       # T = 11 seen heads: the 3-stage ring of the training-step kernel
    ("cityscapes", 12, torch.bfloat16, 2),        # 512 x 1024 crops, K = 20
    ("ade100-50", 4, torch.bfloat16, 5),          # K = 151, 16-bit: one pass, the channel column of a pixel pair in registers
    ("ade100-50", 24, torch.bfloat16, 5),         # ... at the benchmarked batch
    ("ade100-50", 2, torch.float32, 4)])          # K = 151, fp32: two streaming passes (online soft-max, then gradient)
def test_pixel_kernel_other_baseline_configs(name, B, dtype, variant):
    from bacs_b200 import synth
    assert _device_oracle_pixel_check(synth.CONFIGS[name], B, dtype, lowres=False) == variant


@pytest.mark.parametrize("name,B,dtype", [("voc15-1_b24", 24, torch.bfloat16), ("cityscapes", 12, torch.float32),
                                          ("ade100-50", 4, torch.bfloat16)])
def test_lowres_kernel_baseline_configs(name, B, dtype):
    """the fused up-sample + loss + adjoint kernel at BASELINE.json's shapes (sem_logits at stride 16)"""
    from bacs_b200 import synth
    assert _device_oracle_pixel_check(synth.CONFIGS[name], B, dtype, lowres=True) == 3
