"""GPU parity tests, op by op: every CUDA entry point (through the C ABI) against the CPU
oracle on the same seeded inputs.  Integer work is bit-exact; floating point is within the
1e-5 relative tolerance BASELINE.json's north_star states (gradients of 16-bit tensors are
compared after the oracle's gradient is rounded to the same storage type)."""
import math
import os

import numpy as np
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


def dev(t):
    return t.cuda() if isinstance(t, torch.Tensor) else t


def close(got, want, rtol=RTOL, atol=None, what=""):
    got = torch.as_tensor(got).detach().double().cpu()
    want = torch.as_tensor(want).detach().double().cpu()
    if atol is None:
        atol = rtol * max(1e-30, float(want.abs().max()))
    err = float((got - want).abs().max()) if got.numel() else 0.0
    assert torch.allclose(got, want, rtol=rtol, atol=atol), "%s max abs err %.3e (atol %.3e)" % (what, err, atol)


# --------------------------------------------------------------------------------------
# labels
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 7, 8, 1000, 512 * 512 * 3 + 5])
def test_label_hist(ops, n):
    g = torch.Generator().manual_seed(n)
    lab = torch.randint(0, 256, (n,), generator=g)
    if n > 10:
        lab[3] = -4
        lab[5] = 300
    hist = ops.label_hist(lab.cuda()).cpu()
    inside = lab[(lab >= 0) & (lab < 256)]
    assert torch.equal(hist[:256], torch.bincount(inside, minlength=256))
    assert int(hist[256]) == int(((lab < 0) | (lab >= 256)).sum())


def test_label_remap_sequential_aliasing(ops):
    rng = np.random.RandomState(0)
    for trial in range(12):
        keys = rng.choice(35, size=14, replace=False) - 1          # raw ids incl. -1
        d1 = {int(k): int(rng.randint(0, 20)) for k in keys}
        d2 = {int(k): int(v) for k, v in zip(rng.choice(20, 7, replace=False), rng.randint(0, 9, 7))}
        d2[255] = 255
        n_img, H, W = 3, 37, 53
        lbl = rng.randint(-1, 34, size=(n_img, H, W)).astype(np.int64)
        if trial % 3 == 0:
            lbl[0, :4] = 255
        want = np.stack([O.transform_label(lbl[i], d1, 255, d2, 0) for i in range(n_img)])
        lo, n_dom = -1, 257
        m1 = np.full(n_dom, 255, dtype=np.int32)
        for k, v in d1.items():
            m1[k - lo] = v
        m2 = np.full(n_dom, 0, dtype=np.int32)
        for k, v in d2.items():
            m2[k - lo] = v
        got = ops.label_remap(torch.from_numpy(lbl).cuda(), torch.from_numpy(m1).cuda(), 255,
                              torch.from_numpy(m2).cuda(), 0, lo=lo).cpu().numpy()
        assert np.array_equal(got, want), trial
        # single pass
        want1 = np.stack([O.sequential_remap(lbl[i], d1, 255) for i in range(n_img)])
        got1 = ops.label_remap(torch.from_numpy(lbl).cuda(), torch.from_numpy(m1).cuda(), 255, lo=lo).cpu().numpy()
        assert np.array_equal(got1, want1), trial


@pytest.mark.parametrize("shape", [(2, 512, 512, 32, 32), (3, 528, 528, 33, 33), (2, 513, 513, 33, 33),
                                   (2, 64, 96, 4, 6), (1, 100, 75, 7, 5), (2, 512, 1024, 32, 64)])
def test_label_downsample_task(ops, shape):
    B, H, W, h, w = shape
    g = torch.Generator().manual_seed(H + W)
    lab = torch.randint(0, 22, (B, H, W), generator=g)
    lab[torch.rand(B, H, W, generator=g) < 0.1] = 255
    init, inc, T = 16, 1, 6
    lut = torch.from_numpy(O.class_task_lut(init, inc)).int()
    lut[lut >= T] = -1
    task, rank, n_bt, down = ops.label_downsample_task(lab.cuda(), h, w, lut.cuda(), T, want_labels_down=True)
    want_down = O.downsample_labels(lab, h, w)
    assert torch.equal(down.cpu(), want_down)
    want_task = lut[want_down].to(torch.int8)
    assert torch.equal(task.cpu(), want_task)
    tk, rk = want_task.view(B, -1), rank.cpu().view(B, -1)
    for b in range(B):
        for t in range(T):
            sel = tk[b] == t
            assert int(n_bt[b, t]) == int(sel.sum())
            assert torch.equal(rk[b][sel].long(), torch.arange(int(sel.sum())))


# --------------------------------------------------------------------------------------
# prototypes
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,B,dtype", [("tiny", 1, torch.float32), ("tiny", 2, torch.float32),
                                          ("small", 3, torch.float32), ("small", 3, torch.bfloat16)])
def test_proto_accumulate_update(ops, synth, name, B, dtype):
    cfg = synth.CONFIGS[name]
    inp = synth.make_step_inputs(cfg, seed=3, dtype=dtype)
    g = torch.Generator().manual_seed(5)
    mask = synth.make_labels(cfg, g, classes=list(range(1, cfg.K)), B=B)
    pen = inp.pen[:B]
    lut = torch.from_numpy(O.class_task_lut(cfg.initial_classes, cfg.increment)).int()
    lut[lut >= cfg.T] = -1
    task, rank, n_bt, _ = ops.label_downsample_task(mask.cuda(), cfg.h, cfg.w, lut.cuda(), cfg.T)
    for mode, mname in [(0, "exact"), (1, "channel")]:
        sums, counts = ops.proto_accumulate(pen.cuda(), task, rank, n_bt, cfg.T, mode)
        want_s, want_n = O.proto_accumulate(pen.float(), mask, cfg.initial_classes, cfg.increment, cfg.T, mode=mname)
        assert torch.equal(counts.cpu().long(), want_n)
        close(sums, want_s, atol=2e-5 * float(want_s.abs().max()), what="sums " + mname)
    # running-mean update with float32 counts (Q3) and with int64 counts (task 0)
    sums, counts = ops.proto_accumulate(pen.cuda(), task, rank, n_bt, cfg.T, 0)
    want_s, want_n = O.proto_accumulate(pen.float(), mask, cfg.initial_classes, cfg.increment, cfg.T, mode="exact")
    for cdtype in (torch.float32, torch.int64):
        proto = inp.protos.clone()
        cnt = torch.tensor([0, 1000, 3, 0, 7, 1][:cfg.T] + [5] * max(0, cfg.T - 6)).to(cdtype)
        p_gpu, c_gpu = proto.cuda(), cnt.cuda()
        ready = ops.proto_update(p_gpu, c_gpu, sums, counts)
        wp, wc = O.proto_update(proto, cnt, sums.cpu().float(), want_n)
        close(p_gpu, wp, what="proto")
        assert torch.equal(c_gpu.cpu(), wc)
        assert bool(ready.item()) == O.prototypes_ready(wc)
        # the single-process step fuses the update into the finalize launch of the sums: identical bits
        for mode in (0, 1):
            s2, n2 = ops.proto_accumulate(pen.cuda(), task, rank, n_bt, cfg.T, mode)
            p_sep, c_sep = proto.cuda(), cnt.cuda()
            r_sep = ops.proto_update(p_sep, c_sep, s2, n2)
            p_fus, c_fus = proto.cuda(), cnt.cuda()
            s3, n3, r_fus = ops.proto_accumulate_update(pen.cuda(), task, rank, n_bt, mode, p_fus, c_fus)
            assert torch.equal(s3, s2) and torch.equal(n3, n2)
            assert torch.equal(p_fus, p_sep) and torch.equal(c_fus, c_sep) and int(r_fus.item()) == int(r_sep.item())


def test_proto_no_foreground_is_noop(ops, synth):
    cfg = synth.CONFIGS["tiny"]
    inp = synth.make_step_inputs(cfg, seed=1)
    mask = torch.zeros(cfg.B, cfg.H, cfg.W, dtype=torch.int64)
    mask[:, :5] = 255
    lut = torch.from_numpy(O.class_task_lut(cfg.initial_classes, cfg.increment)).int().cuda()
    task, rank, n_bt, _ = ops.label_downsample_task(mask.cuda(), cfg.h, cfg.w, lut, cfg.T)
    sums, counts = ops.proto_accumulate(inp.pen.cuda(), task, rank, n_bt, cfg.T, 0)
    assert float(counts.abs().sum()) == 0 and float(sums.abs().sum()) == 0
    proto, cnt = inp.protos.clone().cuda(), torch.zeros(cfg.T).cuda()
    ready = ops.proto_update(proto, cnt, sums, counts)
    assert torch.equal(proto.cpu(), inp.protos) and int(ready.item()) == 0
    _, _, ready2 = ops.proto_accumulate_update(inp.pen.cuda(), task, rank, n_bt, 0, proto, cnt)
    assert torch.equal(proto.cpu(), inp.protos) and int(ready2.item()) == 0 and float(cnt.abs().sum()) == 0


# --------------------------------------------------------------------------------------
# seen heads
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,dtype", [("tiny", torch.float32), ("small", torch.float32), ("small", torch.bfloat16)])
def test_seen_logits_and_upsample(ops, synth, name, dtype):
    cfg = synth.CONFIGS[name]
    inp = synth.make_step_inputs(cfg, seed=1, dtype=dtype)
    z = ops.seen_logits(inp.pen.cuda(), inp.protos.cuda(), inp.head_w.cuda(), inp.head_b.cuda())
    want = O.seen_logits_lowres(inp.pen.float(), inp.protos, inp.head_w, inp.head_b)
    close(z, want, what="z")
    up = ops.seen_upsample(z, 16, apply_sigmoid=True)
    close(up, O.seen_probs(inp.pen.float(), inp.protos, inp.head_w, inp.head_b), what="seen probs")


def test_seen_logits_from_head_parameters(ops, synth):
    """bacs_seen_logits_heads: the heads as T separate parameter tensors, and the cleared accumulator"""
    cfg = synth.CONFIGS["small"]
    inp = synth.make_step_inputs(cfg, seed=2, dtype=torch.bfloat16)
    pen, protos = inp.pen.cuda(), inp.protos.cuda()
    ws = [inp.head_w[t].clone().reshape(1, -1, 1, 1).cuda() for t in range(cfg.T)]
    bs = [inp.head_b[t].clone().reshape(1).cuda() for t in range(cfg.T)]
    want = ops.seen_logits(pen, protos, inp.head_w.cuda(), inp.head_b.cuda())
    gz = torch.full((cfg.B, cfg.h, cfg.w), 7.0, device="cuda")
    got = ops.seen_logits_heads(pen, protos, ws, bs, zero_out=gz)
    assert torch.equal(got, want) and float(gz.abs().sum()) == 0
    got2 = ops.seen_logits_heads(pen, protos, ws[:2], bs[:2])               # fewer heads than prototypes
    assert torch.equal(got2, want[:, :2])
    with pytest.raises(TypeError):
        ops.seen_logits_heads(pen, protos, [w.half() for w in ws], bs)


def test_teacher_distill_adds_to_a_loss_scalar(ops, synth):
    cfg = synth.CONFIGS["small"]
    inp = synth.make_step_inputs(cfg, seed=2, dtype=torch.bfloat16)
    old, new = inp.old_att.cuda(), inp.new_att.cuda()
    m = (torch.rand(cfg.B, cfg.H, cfg.W, generator=torch.Generator().manual_seed(1)) > 0.4).to(torch.uint8).cuda()
    base = torch.tensor(3.25, device="cuda")
    for mode in (1, 0):                                                      # FMA kernel, then whatever the shape gets
        ops.distill_set_mode(mode)
        try:
            _, d0, l0 = ops.teacher_distill(old, new, m, (cfg.H, cfg.W), 1e-3, True, want_scaled=True)
            _, d1, l1 = ops.teacher_distill(old, new, m, (cfg.H, cfg.W), 1e-3, True, want_scaled=True, addend=base)
        finally:
            ops.distill_set_mode(0)
        assert torch.equal(d0, d1)
        assert float(l1) == float(torch.tensor(float(l0), dtype=torch.float32) + torch.tensor(3.25))


def test_seen_head_backward(ops, synth):
    cfg = synth.CONFIGS["tiny"]
    inp = synth.make_step_inputs(cfg, seed=4)
    t = 1
    pen = inp.pen.clone().requires_grad_(True)
    w = inp.head_w[t:t + 1].clone().requires_grad_(True)
    b = inp.head_b[t:t + 1].clone().requires_grad_(True)
    z = O.seen_logits_lowres(pen, inp.protos[t:t + 1], w, b)
    gz = torch.randn(cfg.B, cfg.h, cfg.w, generator=torch.Generator().manual_seed(0))
    (z[:, 0] * gz).sum().backward()
    scale = torch.tensor([0.37])
    dw, db, dpen = ops.seen_head_backward(inp.pen.cuda(), inp.protos[t].cuda(), inp.head_w[t].contiguous().cuda(),
                                          gz.cuda(), scale.cuda(), True)
    close(dw, 0.37 * w.grad[0], what="dweight")
    close(db, 0.37 * b.grad, what="dbias")
    close(dpen, 0.37 * pen.grad, what="dpen")


# --------------------------------------------------------------------------------------
# fused per-pixel kernel
# --------------------------------------------------------------------------------------
def _seen_z(inp):
    return O.seen_logits_lowres(inp.pen.float(), inp.protos, inp.head_w, inp.head_b)


@pytest.mark.parametrize("ukd", [True, False])
@pytest.mark.parametrize("name,dtype", [("tiny", torch.float32), ("small", torch.float32), ("small", torch.bfloat16),
                                        ("small", torch.float16)])
def test_pixel_weighted_ce(ops, synth, name, dtype, ukd):
    from bacs_b200 import _cabi
    cfg = synth.CONFIGS[name]
    inp = synth.make_step_inputs(cfg, seed=2, dtype=dtype)
    g = torch.Generator().manual_seed(9)
    mask = synth.make_labels(cfg, g, classes=list(range(1, cfg.K)))
    z = _seen_z(inp)
    smax = torch.sigmoid(O.bilinear_upsample(z, (cfg.H, cfg.W), True)).max(1)[0]
    x = inp.logits.float().clone().requires_grad_(True)
    want = O.weighted_ce(x, mask, smax, cfg.old_cl, 2.0, 0.5, ukd)
    want.backward()
    scale = 1024.0 if dtype == torch.float16 else 1.0      # fp16 gradients of a mean need a loss scale
    out = ops.pixel_loss(inp.logits.cuda(), mask.cuda(), _cabi.PIX_WEIGHTED_CE, want_grad=True, z=z.cuda(),
                         want_distill_mask=True, old_cl=cfg.old_cl, ukd=ukd, grad_scale=scale)
    N = cfg.B * cfg.H * cfg.W
    close(out["acc"][_cabi.ACC_LOSS] / N, want, what="loss")
    assert torch.equal(out["preds"].cpu(), O.argmax_first(inp.logits.float()))
    want_g = (x.grad * scale).to(dtype).float()
    tol = RTOL if dtype == torch.float32 else 2.0 ** (-7 if dtype == torch.bfloat16 else -10)   # 1 ulp of the storage type
    close(out["dlogits"].float(), want_g, atol=tol * float(want_g.abs().max()), what="dlogits")
    # teacher-distill pixel mask: exact away from the fp32 rounding band of the threshold
    want_m = (mask == 0) & (smax > 0.5)
    diff = out["distill_mask"].cpu().bool() != want_m
    assert int((diff & ((smax - 0.5).abs() > 1e-6)).sum()) == 0
    assert int(diff.sum()) <= 2
    acc = out["acc"].cpu()
    assert int(acc[_cabi.ACC_KEPT]) == int((mask != 255).sum())
    assert int(acc[_cabi.ACC_BG]) == int((mask == 0).sum())
    assert int(acc[_cabi.ACC_INVALID]) == 0


@pytest.mark.parametrize("name,dtype,ukd", [
    ("row512", torch.float32, True), ("row512", torch.bfloat16, True), ("row512", torch.float16, False),
    ("row1024", torch.bfloat16, True), ("row1024", torch.float32, False), ("row512_k17", torch.bfloat16, True),
    ("row512_k17", torch.float32, True), ("row512_k11", torch.bfloat16, False), ("row512_k11", torch.float32, True),
    ("row512_k7", torch.bfloat16, True), ("row512_k7", torch.float32, True),
    ("row512_t11", torch.bfloat16, True), ("row512_t11", torch.float32, True), ("row512_k24", torch.bfloat16, True),
    ("row512_k24", torch.float16, False), ("row512_k40", torch.bfloat16, True), ("row512_k40", torch.float32, True),
    ("row512_k40", torch.float16, False), ("row512_k151", torch.bfloat16, True), ("row512_k151", torch.float32, True),
    ("row512_k151", torch.float16, False)])
def test_pixel_training_step_kernel(ops, synth, name, dtype, ukd):
    """512-pixel row tiles take the specialised training-step kernel: weighted CE + focal term of one head +
    distill mask + arg-max + both gradients in one launch, incl. padded class counts and invalid labels."""
    from bacs_b200 import _cabi
    cfg = synth.CONFIGS[name]
    inp = synth.make_step_inputs(cfg, seed=5, dtype=dtype)
    g = torch.Generator().manual_seed(3)
    mask = synth.make_labels(cfg, g, classes=list(range(1, cfg.K)))
    mask[0, 9, 40:47] = cfg.K + 3                       # outside [0,K), not ignore: counted, treated as ignore
    t = cfg.T - 1
    z = _seen_z(inp).clone().requires_grad_(True)
    up = O.bilinear_upsample(z, (cfg.H, cfg.W), True)
    smax = torch.sigmoid(up).max(1)[0].detach()
    x = inp.logits.float().clone().requires_grad_(True)
    clean = torch.where(mask > 255, torch.full_like(mask, 255), mask)
    clean = torch.where((clean >= cfg.K) & (clean != 255), torch.full_like(mask, 255), clean)
    want = O.weighted_ce(x, clean, smax, cfg.old_cl, 2.0, 0.5, ukd)
    want.backward()
    kept = int((clean != 255).sum())
    want_f = O.focal_seen_loss(up[:, t:t + 1], clean, 2.0, 0.25)
    want_f.backward()
    scale = 1024.0 if dtype == torch.float16 else 1.0
    out = ops.pixel_loss(inp.logits.cuda(), mask.cuda(), _cabi.PIX_WEIGHTED_CE, want_grad=True, z=z.detach().cuda(),
                         want_distill_mask=True, old_cl=cfg.old_cl, ukd=ukd, grad_scale=scale, focal_head=t,
                         focal_alpha=0.25)
    want_variant = 2 if cfg.K <= 24 else ((4 if dtype == torch.float32 else 5) if cfg.K >= 64 else 0)
    assert out["variant"] == want_variant, \
        "training-step kernel (K <= 24) / tile kernel / K >= 64: two streaming passes (fp32), register column (16-bit)"
    N = cfg.B * cfg.H * cfg.W
    acc = out["acc"].cpu()
    close(acc[_cabi.ACC_LOSS] / N, want, what="loss")
    close(acc[_cabi.ACC_FOCAL] / kept, want_f, what="focal loss")
    close(out["gz"] / kept, z.grad[:, t], atol=2e-5 * float(z.grad.abs().max()), what="gz")
    assert torch.equal(out["preds"].cpu(), O.argmax_first(inp.logits.float()))
    want_g = (x.grad * scale).to(dtype).float()
    tol = RTOL if dtype == torch.float32 else 2.0 ** (-7 if dtype == torch.bfloat16 else -10)
    close(out["dlogits"].float(), want_g, atol=tol * float(want_g.abs().max()), what="dlogits")
    want_m = (clean == 0) & (smax > 0.5)
    diff = out["distill_mask"].cpu().bool() != want_m
    assert int((diff & ((smax - 0.5).abs() > 1e-6)).sum()) == 0
    assert int(diff.sum()) <= 2
    assert int(acc[_cabi.ACC_KEPT]) == kept
    assert int(acc[_cabi.ACC_VALID]) == kept
    assert int(acc[_cabi.ACC_BG]) == int((clean == 0).sum())
    assert int(acc[_cabi.ACC_INVALID]) == 7
    assert int(acc[_cabi.ACC_DISTILL_PIX]) == int(out["distill_mask"].sum())
    # forward only (evaluation): same loss, no gradient tensors
    out2 = ops.pixel_loss(inp.logits.cuda(), mask.cuda(), _cabi.PIX_WEIGHTED_CE, want_grad=False, z=z.detach().cuda(),
                          old_cl=cfg.old_cl, ukd=ukd)
    close(out2["acc"][_cabi.ACC_LOSS] / N, want, what="loss (no grad)")
    assert torch.equal(out2["preds"], out["preds"])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pixel_training_step_kernel_large_logits(ops, synth, dtype):
    """Confident logits (|x| up to ~40, far beyond random init): every log-sum-exp stays finite and matches the oracle;
    pixels whose foreground mass underflows completely in fp32 are excluded (the closed form has no fp32 value there)."""
    from bacs_b200 import _cabi
    cfg = synth.CONFIGS["row512"]
    inp = synth.make_step_inputs(cfg, seed=9, dtype=dtype)
    logits = (inp.logits.float() * 10.0).to(dtype)
    g = torch.Generator().manual_seed(5)
    mask = synth.make_labels(cfg, g, classes=list(range(1, cfg.K)))
    z = _seen_z(inp)
    smax = torch.sigmoid(O.bilinear_upsample(z, (cfg.H, cfg.W), True)).max(1)[0]
    x = logits.float().clone().requires_grad_(True)
    want = O.weighted_ce(x, mask, smax, cfg.old_cl, 2.0, 0.5, True)
    want.backward()
    out = ops.pixel_loss(logits.cuda(), mask.cuda(), _cabi.PIX_WEIGHTED_CE, want_grad=True, z=z.cuda(),
                         want_distill_mask=True, old_cl=cfg.old_cl, ukd=True)
    assert out["variant"] == 2
    assert bool(torch.isfinite(out["acc"]).all()) and bool(torch.isfinite(out["dlogits"].float()).all())
    N = cfg.B * cfg.H * cfg.W
    close(out["acc"][_cabi.ACC_LOSS] / N, want, rtol=2e-5, what="loss")
    assert torch.equal(out["preds"].cpu(), O.argmax_first(logits.float()))
    want_g = x.grad.to(dtype).float()
    tol = 2e-5 if dtype == torch.float32 else 2.0 ** -7
    close(out["dlogits"].float(), want_g, atol=tol * float(want_g.abs().max()), what="dlogits")


def test_pixel_focal_term(ops, synth):
    from bacs_b200 import _cabi
    cfg = synth.CONFIGS["tiny"]
    for alpha in (None, 0.25):
        inp = synth.make_step_inputs(cfg, seed=8)
        t = 1
        z = _seen_z(inp).clone().requires_grad_(True)
        zf = O.bilinear_upsample(z[:, t:t + 1], (cfg.H, cfg.W), True)
        kept = int((inp.mask != 255).sum())
        want = O.focal_seen_loss(zf, inp.mask, 2.0, alpha)
        want.backward()
        out = ops.pixel_loss(inp.logits.cuda(), inp.mask.cuda(), _cabi.PIX_WEIGHTED_CE, want_grad=False,
                             z=z.detach().cuda(), focal_head=t, old_cl=cfg.old_cl, focal_alpha=alpha)
        close(out["acc"][_cabi.ACC_FOCAL] / kept, want, what="focal loss")
        close(out["gz"] / kept, z.grad[:, t], atol=2e-5 * float(z.grad.abs().max()), what="gz")
        scale, out2 = ops.focal_scale(out["acc"], None, 0.75)
        close(out2[1], 0.75 * want, what="focal scaled")
        close(scale, 0.75 / kept, what="focal scale")
        # the one-launch variant that also assembles the loss scalar: main term + focal term
        scale2, loss = ops.focal_scale_loss(out["acc"], None, 0.75, 0.5, False)
        close(scale2, 0.75 / kept, what="focal scale (fused)")
        close(loss, 0.5 * float(out["acc"][_cabi.ACC_LOSS]) + 0.75 * float(want), what="fused loss scalar")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pixel_ce_modes(ops, synth, dtype):
    from bacs_b200 import _cabi
    cfg = synth.CONFIGS["small"]
    inp = synth.make_step_inputs(cfg, seed=4, dtype=dtype)
    g = torch.Generator().manual_seed(2)
    mask = synth.make_labels(cfg, g, classes=list(range(1, cfg.K)))
    w = torch.zeros(cfg.K)
    w[1:cfg.old_cl] = 1
    tol = RTOL if dtype == torch.float32 else 2.0 ** -7            # 1 bf16 ulp
    for weight in (None, w):
        x = inp.logits.float().clone().requires_grad_(True)
        want = O.cross_entropy(x, mask, weight)
        want.backward()
        out = ops.pixel_loss(inp.logits.cuda(), mask.cuda(), _cabi.PIX_CE, want_grad=True,
                             class_w=None if weight is None else weight.cuda(), grad_scale=0.2)
        acc = out["acc"]
        close(acc[_cabi.ACC_LOSS] / acc[_cabi.ACC_WSUM], want, what="ce")
        wg = (0.2 * x.grad).to(dtype).float()
        close(out["dlogits"].float(), wg, atol=tol * float(wg.abs().max()), what="ce grad")
        assert torch.equal(out["preds"].cpu(), O.argmax_first(inp.logits.float()))
    # eval: no gradient, no histogram
    out = ops.pixel_loss(inp.logits.cuda(), mask.cuda(), _cabi.PIX_CE, want_grad=False)
    close(out["acc"][_cabi.ACC_LOSS] / out["acc"][_cabi.ACC_WSUM], O.cross_entropy(inp.logits.float(), mask))
    # unbiased CE
    x = inp.logits.float().clone().requires_grad_(True)
    want = O.unbiased_ce(x, mask, cfg.old_cl)
    want.backward()
    out = ops.pixel_loss(inp.logits.cuda(), mask.cuda(), _cabi.PIX_UNBIASED_CE, want_grad=True, old_cl=cfg.old_cl)
    close(out["acc"][_cabi.ACC_LOSS] / out["acc"][_cabi.ACC_WSUM], want, what="uce")
    wg = x.grad.to(dtype).float()
    close(out["dlogits"].float(), wg, atol=tol * float(wg.abs().max()), what="uce grad")
    # per-image importance score
    w2 = torch.ones(cfg.K)
    w2[0] = 0
    out = ops.pixel_loss(inp.logits.cuda(), mask.cuda(), _cabi.PIX_SCORE, want_grad=False, class_w=w2.cuda(),
                         want_score=True, want_preds=False)
    close(out["score"], O.cross_entropy_per_image_score(inp.logits.float(), mask, w2), what="score")


@pytest.mark.parametrize("shape", [(1, 5, 33, 47), (2, 3, 16, 18), (1, 151, 32, 64), (1, 40, 31, 33)])
def test_pixel_ragged_shapes(ops, shape):
    """odd sizes / unaligned rows take the non-TMA loader and the 1-pixel-per-thread plan"""
    from bacs_b200 import _cabi
    B, K, H, W = shape
    g = torch.Generator().manual_seed(K)
    x = torch.randn(B, K, H, W, generator=g)
    y = torch.randint(0, K, (B, H, W), generator=g)
    y[torch.rand(B, H, W, generator=g) < 0.2] = 255
    xr = x.clone().requires_grad_(True)
    want = O.cross_entropy(xr, y)
    want.backward()
    out = ops.pixel_loss(x.cuda(), y.cuda(), _cabi.PIX_CE, want_grad=True)
    close(out["acc"][_cabi.ACC_LOSS] / out["acc"][_cabi.ACC_WSUM], want)
    close(out["dlogits"], xr.grad, what="grad")
    assert torch.equal(out["preds"].cpu(), x.argmax(1))


def test_pixel_all_ignored_and_invalid(ops):
    from bacs_b200 import _cabi
    x = torch.randn(1, 4, 16, 32)
    y = torch.full((1, 16, 32), 255)
    out = ops.pixel_loss(x.cuda(), y.cuda(), _cabi.PIX_CE, want_grad=True)
    assert float(out["acc"][_cabi.ACC_WSUM]) == 0 and float(out["dlogits"].abs().max()) == 0
    y[0, 0, :5] = 9            # outside [0,K) and not ignore: counted, treated as ignore
    out = ops.pixel_loss(x.cuda(), y.cuda(), _cabi.PIX_CE, want_grad=False)
    assert int(out["acc"][_cabi.ACC_INVALID]) == 5


# --------------------------------------------------------------------------------------
# teacher distill / DER
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,dtype,with_mask", [("tiny", torch.float32, True), ("tiny", torch.float32, False),
                                                  ("small", torch.float32, True), ("small", torch.bfloat16, True),
                                                  ("row512", torch.float32, True), ("row1024", torch.float32, True),
                                                  ("row1024", torch.bfloat16, False), ("wide1536", torch.float32, True)])
def test_teacher_distill(ops, synth, name, dtype, with_mask):
    cfg = synth.CONFIGS[name]
    inp = synth.make_step_inputs(cfg, seed=6, dtype=dtype)
    g = torch.Generator().manual_seed(1)
    m = (inp.mask == 0) & (torch.rand(cfg.B, cfg.H, cfg.W, generator=g) > 0.4) if with_mask else None
    new = inp.new_att.float().clone().requires_grad_(True)
    lab = inp.mask if m is None else torch.where(m, torch.zeros_like(inp.mask), torch.ones_like(inp.mask))
    if m is None:
        lab = torch.zeros_like(inp.mask)
    want = O.teacher_distill(inp.old_att.float(), new, lab, None, lkd=0.25)
    want.backward()
    coef = 0.25 / (cfg.B * cfg.A * cfg.H)
    mask_u8 = None if m is None else m.to(torch.uint8).cuda()
    s, dnew = ops.teacher_distill(inp.old_att.cuda(), inp.new_att.cuda(), mask_u8, (cfg.H, cfg.W), coef, True)
    close(s * coef, want, what="distill loss")
    tol = RTOL * 3 if dtype == torch.float32 else 2.0 ** -7
    wg = new.grad.to(dtype).float()
    close(dnew.float(), wg, atol=tol * float(wg.abs().max()), what="dnew")


def test_teacher_distill_identical_maps_zero(ops, synth):
    cfg = synth.CONFIGS["tiny"]
    inp = synth.make_step_inputs(cfg, seed=6)
    s, dnew = ops.teacher_distill(inp.new_att.cuda(), inp.new_att.cuda(), None, (cfg.H, cfg.W), 1.0, True)
    assert float(s) == 0.0 and float(dnew.abs().max()) == 0.0


@pytest.mark.parametrize("ncls", [[6, 6, 8, 7, 9], [9, 9, 9, 9, 9], [5, 6, 7, 8, 9], [8, 5, 5, 8, 6]])
@pytest.mark.parametrize("ignore_bg", [True, False])
def test_der_mse(ops, ncls, ignore_bg):
    K, Br = 9, 5
    g = torch.Generator().manual_seed(sum(ncls))
    s = torch.randn(Br, K, 4, 6, generator=g)
    m = torch.randn(Br, K, 4, 6, generator=g) * 2
    sr = s.clone().requires_grad_(True)
    want = O.der_mse(sr, m, np.array(ncls), ignore_rep_bg=ignore_bg)
    want.backward()
    cut = torch.from_numpy(ops.der_transplant_cut(ncls, K)).cuda()
    assert np.array_equal(ops.der_transplant_cut(ncls, K), O.der_transplant_cut(np.array(ncls), K))
    assert np.array_equal(ops.der_cut(torch.tensor(ncls, dtype=torch.uint8).cuda(), K).cpu().numpy(),
                          O.der_transplant_cut(np.array(ncls), K))
    coef = 0.8 / s.numel()
    for mem in (m.cuda(), m.long().cuda()):                      # float (truncated in-kernel) or int64 (Q4)
        tot, ds = ops.der_mse(s.cuda(), mem, cut, ignore_bg, True, coef, True)
        close(tot / s.numel(), want, what="der")
        close(ds, 0.8 * sr.grad, atol=1e-6 * float(sr.grad.abs().max()), what="dsem")


def test_unbiased_kd_module(synth):
    from bacs_b200.training.loss_utils import UnbiasedKnowledgeDistillationLoss
    cfg = synth.CONFIGS["small"]
    inp = synth.make_step_inputs(cfg, seed=4)
    g = torch.Generator().manual_seed(2)
    old = torch.randn(cfg.B, cfg.old_cl, cfg.H, cfg.W, generator=g)
    for alpha, use_mask in ((1.0, False), (0.5, True)):
        pm = (inp.mask == 0) if use_mask else None
        x = inp.logits.clone().requires_grad_(True)
        want = O.unbiased_kd(x, old, alpha=alpha, mask=pm)
        want.backward()
        xg = inp.logits.clone().cuda().requires_grad_(True)
        got = UnbiasedKnowledgeDistillationLoss(alpha=alpha)(xg, old.cuda(), None if pm is None else pm.cuda())
        got.backward()
        close(got, want, what="ukd")
        close(xg.grad, x.grad, atol=1e-5 * float(x.grad.abs().max()), what="ukd grad")


# --------------------------------------------------------------------------------------
# confusion matrix
# --------------------------------------------------------------------------------------
def test_confmat_known_answer(ops):
    # the reference's only known-answer vector: training/metrics.py:159-183
    label = torch.zeros((1, 4, 4), dtype=torch.long)
    pred = torch.zeros((1, 4, 4), dtype=torch.float32)
    label[:, :3, :3] = 1
    pred[:, -3:, -3:] = 1
    cm = torch.zeros(2, 2, dtype=torch.int64).cuda()
    ops.confmat_accumulate(pred.cuda(), label.cuda(), 2, cm)
    assert cm.cpu().tolist() == [[2, 5], [5, 4]]
    met = ops.confmat_metrics(cm).cpu()
    assert torch.allclose(met[0], torch.tensor([2.0 / 12, 4.0 / 14]), atol=1e-6)


@pytest.mark.parametrize("K,n", [(2, 100), (21, 512 * 512 * 2 + 3), (20, 1024 * 2048), (151, 300000), (256, 50000)])
def test_confmat_matches_oracle(ops, K, n):
    g = torch.Generator().manual_seed(K)
    t = torch.randint(0, K, (n,), generator=g)
    t[torch.rand(n, generator=g) < 0.1] = 255
    p = torch.randint(0, K, (n,), generator=g)
    # long runs like real segmentation maps
    t = t.view(-1)[torch.arange(n) // 7 * 7 % n]
    cm = torch.zeros(K, K, dtype=torch.int64).cuda()
    ops.confmat_accumulate(p.cuda(), t.cuda(), K, cm)
    ops.confmat_accumulate(p.cuda(), t.cuda(), K, cm)            # accumulates
    want = O.confusion_matrix(p.numpy(), t.numpy(), K)
    assert np.array_equal(cm.cpu().numpy(), 2 * want)
    met = ops.confmat_metrics(cm).cpu().numpy()
    wm = O.iou_metrics(2 * want)
    for row, key in enumerate(["iou_per_class", "accuracy", "precision", "recall", "specificity"]):
        assert np.allclose(met[row], wm[key], rtol=1e-6, atol=1e-7), key
    assert np.allclose(met[5], wm["miou"], rtol=1e-6)


def test_pack_unpack_and_scale(ops):
    T, D, K = 3, 8, 4
    sums = torch.randn(T, D, dtype=torch.float64).cuda()
    counts = torch.tensor([3.0, 0.0, 7.0], dtype=torch.float64).cuda()
    cm = torch.randint(0, 1 << 40, (K, K)).cuda()
    packed = ops.pack_state(sums, counts, cm, sums.device)
    assert packed.numel() == T * D + T + K * K
    s2, c2, cm2 = torch.empty_like(sums), torch.empty_like(counts), torch.empty_like(cm)
    ops.unpack_state(packed * 2, s2, c2, cm2)
    assert torch.equal(s2, sums * 2) and torch.equal(c2, counts * 2) and torch.equal(cm2, cm * 2)
    x = torch.randn(1000).cuda()
    y = x.clone()
    ops.scale_inplace(y, torch.tensor([1.0]).cuda())
    assert torch.equal(x, y)
    ops.scale_inplace(y, torch.tensor([0.5]).cuda())
    assert torch.equal(y, x * 0.5)
    # several tensors of mixed types in one launch; None entries are skipped
    t1, t2, t3 = torch.randn(5000).cuda(), torch.randn(3, 7).cuda().bfloat16(), torch.randn(1).cuda().half()
    w1, w2, w3 = t1.clone(), t2.clone(), t3.clone()
    ops.scale_inplace_multi([t1, None, t2, t3], torch.tensor(1.0).cuda())
    assert torch.equal(t1, w1) and torch.equal(t2, w2) and torch.equal(t3, w3)
    ops.scale_inplace_multi([t1, None, t2, t3], torch.tensor(0.25).cuda())
    assert torch.equal(t1, w1 * 0.25) and torch.equal(t2, w2 * 0.25) and torch.equal(t3, w3 * 0.25)
    a = torch.tensor([2.0, 8.0], dtype=torch.float64).cuda()
    r = ops.combine_scalars([(a, 0, 3.0), (a, 1, 1.0, a, 0)], a.device)
    assert math.isclose(float(r), 2 * 3 + 8 / 2)


def test_transform_label_class(ops):
    """labels.TransformLabel (reference constructor) incl. the background-shift map builder."""
    from bacs_b200 import labels as L
    class_order = [5, 3, 8, 1, 2, 9, 4, 7, 6]                     # shuffled order: aliasing chains (Q13)
    id2train = {i: i for i in range(10)}
    id2train[255] = 255
    rng = np.random.RandomState(1)
    lbl = rng.randint(0, 10, size=(4, 40, 56)).astype(np.int64)
    lbl[:, :3] = 255
    for train, test_bg, tasks in ((True, True, [8, 1]), (False, True, [5, 3, 8, 1]), (False, False, [5, 3])):
        inv, masking = L.build_inverted_order(class_order, tasks, train, test_bg)
        winv, wmask = O.build_inverted_order(class_order, tasks, train, test_bg)
        assert inv == winv and masking == wmask
        want = np.stack([O.transform_label(lbl[i], id2train, 255, inv, masking) for i in range(lbl.shape[0])])
        tl = L.TransformLabel(id2train, 255, inv, masking)
        got = tl(torch.from_numpy(lbl).cuda()).cpu().numpy()
        assert np.array_equal(got, want)
        assert np.array_equal(tl(torch.from_numpy(lbl[0]).cuda()).cpu().numpy(), want[0])


# --------------------------------------------------------------------------------------
# HBM-resident replay store
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,dtype", [((7, 3, 16, 24), torch.float32), ((5, 9), torch.uint8), ((4, 6, 2, 2), torch.float32),
                                         ((3, 1, 5, 7), torch.int64), ((6, 13), torch.bfloat16)])
def test_gather_rows(ops, shape, dtype):
    g = torch.Generator().manual_seed(shape[0])
    src = (torch.randn(shape, generator=g) * 50).to(dtype)
    idx = torch.randint(0, shape[0], (11,), generator=g)
    got = ops.gather_rows(src.cuda(), idx.cuda()).cpu()
    assert torch.equal(got, src[idx])
    assert ops.gather_rows(src.cuda(), idx[:0].cuda()).shape[0] == 0
    bad = torch.tensor([0, shape[0] + 3, -1])
    got = ops.gather_rows(src.cuda(), bad.cuda()).cpu()        # out-of-range rows come back as zeros, not a fault
    assert torch.equal(got[0], src[0]) and float(got[1:].float().abs().sum()) == 0.0


def test_device_replay_store_matches_buffer(tmp_path, monkeypatch):
    """DeviceReplayStore.get_data == Buffer.get_data(device=cuda) for the same np.random stream (bit-exact fields,
    same indices, same n_classes / task_id), incl. after more data arrived and only the touched slots were refreshed."""
    monkeypatch.setenv("BACS_BUFFER_ROOT", str(tmp_path))
    from bacs_b200.training.buffer import Buffer
    from bacs_b200.training.device_store import DeviceReplayStore
    rng = np.random.RandomState(3)

    def batch(B=3, K=5):
        return {"examples": torch.from_numpy(rng.rand(B, 3, 32, 48).astype(np.float32)),
                "logits": torch.from_numpy(rng.randn(B, K, 2, 3).astype(np.float32)),
                "labels": torch.from_numpy(rng.randint(0, K, size=(B, 32, 48)).astype(np.int64)),
                "seen": torch.from_numpy(rng.randn(B, 1, 32, 48).astype(np.float32)),
                "loss": torch.from_numpy(-rng.rand(B).astype(np.float32))}
    np.random.seed(0)
    buf = Buffer(6, "all_tasks")
    buf.update_task(task_num=0, new_class_size=5)
    for _ in range(3):
        buf.add_data(batch())
    buf.merge_scores()
    store = DeviceReplayStore(buf, "cuda")
    for trial in range(3):
        np.random.seed(10 + trial)
        want = buf.get_data(4, device="cuda")
        np.random.seed(10 + trial)
        got, choice = store.get_data(4, return_indexes=True)
        assert set(got) == set(want)
        for key in ("examples", "logits", "labels", "seen"):
            assert got[key].is_cuda and got[key].dtype == want[key].dtype and torch.equal(got[key], want[key]), key
        assert np.array_equal(got["n_classes"], want["n_classes"]) and got["task_id"] == want["task_id"]
    # reservoir replacement, then refresh only what changed
    before = np.array(buf.dataset_map["examples"][:])
    buf.add_data(batch())
    changed = np.where(np.abs(np.array(buf.dataset_map["examples"][:]) - before).reshape(6, -1).sum(1) > 0)[0]
    store.refresh(changed)
    np.random.seed(99)
    want = buf.get_data(5, device="cuda")
    np.random.seed(99)
    got = store.get_data(5)
    for key in ("examples", "logits", "labels", "seen"):
        assert torch.equal(got[key], want[key]), key
    # the "bufferlogits" loader: every slot once per epoch, rows intact
    seen_rows = []
    for ex, lg, ncl in store.logits_batches(4, length=6, generator=torch.Generator().manual_seed(1)):
        assert ex.shape[0] == lg.shape[0] == ncl.shape[0]
        seen_rows.append(lg.cpu())
    all_rows = torch.cat(seen_rows)
    ref = torch.from_numpy(np.array(buf.dataset_map["logits"][:]))
    assert all_rows.shape == ref.shape
    assert sorted(map(float, all_rows.reshape(6, -1).sum(1))) == sorted(map(float, ref.reshape(6, -1).sum(1)))


# --------------------------------------------------------------------------------------
# optional per-class prototype family: pixel x class-prototype distance on the tensor cores
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,D,h,w,Kc", [(2, 64, 8, 16, 10), (1, 72, 4, 6, 3), (3, 512, 32, 32, 150), (2, 256, 33, 40, 21)])
def test_class_distance_tensor_cores(ops, B, D, h, w, Kc):
    """bf16 operands are exact products in the fp32 accumulator: against the oracle fed the same bf16-rounded inputs the
    squared distances agree to fp32 rounding of the |f|^2 + |c|^2 - 2 f.c expansion (1e-5 of the largest distance)."""
    g = torch.Generator().manual_seed(Kc)
    f = torch.randn(B, D, h, w, generator=g).to(torch.bfloat16)
    c = torch.randn(Kc, D, generator=g).to(torch.bfloat16)
    c[0] = f[0, :, 1, 2]                                   # a pixel that coincides with a prototype: distance 0
    want, want_near = O.class_distance(f.float(), c.float())
    got, near = ops.class_distance(f.cuda(), c.cuda())
    assert tuple(got.shape) == (B, Kc, h, w) and got.dtype == torch.float32
    close(got, want, atol=1e-5 * float(want.max()), what="dist2")
    assert float(got.min()) >= 0.0 and float(got[0, 0, 1, 2]) <= 1e-5 * float(want.max())
    top2 = want.topk(2, dim=1, largest=False)[0]
    diff = near.cpu() != want_near
    assert int((diff & ((top2[:, 1] - top2[:, 0]) > 2e-5 * float(want.max()))).sum()) == 0
    d2, none = ops.class_distance(f.cuda(), c.cuda(), want_nearest=False)
    assert none is None and torch.equal(d2, got)


def test_class_distance_limits(ops):
    """shapes whose prototypes do not fit one SM's shared memory are refused, never computed some other way"""
    from bacs_b200._cabi import BacsError
    f = torch.randn(1, 1024, 4, 8).to(torch.bfloat16).cuda()
    with pytest.raises(BacsError):
        ops.class_distance(f, torch.randn(10, 1024).to(torch.bfloat16).cuda())
    f = torch.randn(2, 256, 8, 8, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
    c = torch.randn(200, 256, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16)
    want, _ = O.class_distance(f.float(), c.float())
    got, _ = ops.class_distance(f.cuda(), c.cuda())                    # 200 classes x 256 channels: fits
    close(got, want, atol=1e-5 * float(want.max()), what="dist2 200 classes")


@pytest.mark.parametrize("B,D,h,w,K,dtype", [(3, 64, 8, 16, 21, torch.float32), (2, 40, 5, 7, 151, torch.bfloat16),
                                            (4, 512, 32, 32, 151, torch.bfloat16), (1, 8, 3, 3, 256, torch.float16),
                                            (2, 32, 8, 8, 21, torch.float16), (3, 48, 4, 16, 100, torch.bfloat16)])
def test_class_sums(ops, B, D, h, w, K, dtype):
    """per-class segmented sums (G = K) and counts against index_add / bincount"""
    g = torch.Generator().manual_seed(K + D)
    f = torch.randn(B, D, h, w, generator=g).to(dtype)
    lab = torch.randint(0, K, (B, h, w), generator=g)
    lab[torch.rand(B, h, w, generator=g) < 0.2] = 255 if K <= 255 else 300      # ignore / out of range
    lab[0, 0, 0] = 0
    want_s, want_n = O.class_sums(f.float(), lab, K)
    got_s, got_n = ops.class_sums(f.cuda(), lab.cuda(), K)
    assert got_s.dtype == torch.float64 and tuple(got_s.shape) == (K, D)
    if K <= 255:
        assert torch.equal(got_n.cpu(), want_n)
    close(got_s, want_s, atol=1e-5 * float(want_s.abs().max()), what="class sums")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pixel_modes_on_the_streaming_path(ops, dtype):
    """K >= 64 takes the two streaming passes (fp32, variant 4) or the one-pass register-column kernel (16-bit, variant
    5): plain / class-weighted CE, unbiased CE and the per-image score"""
    from bacs_b200 import _cabi
    B, K, H, W, old_cl = 3, 70, 16, 64, 41
    g = torch.Generator().manual_seed(70)
    x0 = (torch.randn(B, K, H, W, generator=g) * 3).to(dtype)
    y = torch.randint(0, K, (B, H, W), generator=g)
    y[torch.rand(B, H, W, generator=g) < 0.15] = 255
    y[1, 3, 5:9] = K + 7                                   # invalid, not ignore
    clean = torch.where((y >= K) & (y != 255), torch.full_like(y, 255), y)
    w = torch.rand(K, generator=g)
    tol = 1e-5 if dtype == torch.float32 else 2.0 ** -7
    variant = 4 if dtype == torch.float32 else 5
    for weight in (None, w):
        x = x0.float().clone().requires_grad_(True)
        want = O.cross_entropy(x, clean, weight)
        want.backward()
        out = ops.pixel_loss(x0.cuda(), y.cuda(), _cabi.PIX_CE, want_grad=True,
                             class_w=None if weight is None else weight.cuda())
        assert out["variant"] == variant
        acc = out["acc"]
        close(acc[_cabi.ACC_LOSS] / acc[_cabi.ACC_WSUM], want, what="ce")
        wg = x.grad.to(dtype).float()
        close(out["dlogits"].float(), wg, atol=tol * float(wg.abs().max()), what="ce grad")
        assert torch.equal(out["preds"].cpu(), O.argmax_first(x0.float()))
        assert int(acc[_cabi.ACC_INVALID]) == 4
    x = x0.float().clone().requires_grad_(True)
    want = O.unbiased_ce(x, clean, old_cl)
    want.backward()
    out = ops.pixel_loss(x0.cuda(), y.cuda(), _cabi.PIX_UNBIASED_CE, want_grad=True, old_cl=old_cl)
    close(out["acc"][_cabi.ACC_LOSS] / out["acc"][_cabi.ACC_WSUM], want, what="uce")
    wg = x.grad.to(dtype).float()
    close(out["dlogits"].float(), wg, atol=tol * float(wg.abs().max()), what="uce grad")
    w2 = torch.ones(K)
    w2[0] = 0
    out = ops.pixel_loss(x0.cuda(), y.cuda(), _cabi.PIX_SCORE, want_grad=False, class_w=w2.cuda(), want_score=True,
                         want_preds=False)
    assert out["variant"] == 4, "per-image scores stay on the streaming path"
    close(out["score"], O.cross_entropy_per_image_score(x0.float(), clean, w2), what="score")
    # confident logits: the exponent reference is lifted when the running max runs ahead (|x| up to ~60)
    xb = (x0.float() * 6).to(dtype)
    x = xb.float().clone().requires_grad_(True)
    want = O.cross_entropy(x, clean, None)
    want.backward()
    out = ops.pixel_loss(xb.cuda(), y.cuda(), _cabi.PIX_CE, want_grad=True)
    close(out["acc"][_cabi.ACC_LOSS] / out["acc"][_cabi.ACC_WSUM], want, rtol=2e-5, what="ce (large logits)")
    wg = x.grad.to(dtype).float()
    close(out["dlogits"].float(), wg, atol=tol * float(wg.abs().max()), what="ce grad (large logits)")


@pytest.mark.parametrize("K,old_cl,H,W,dtype", [
    (151, 101, 16, 64, torch.bfloat16),    # ADE20K 100-50: the boundary falls into the upper lane's channel half
    (151, 151, 16, 32, torch.bfloat16),    # no new classes (old_cl = K)
    (151, 1, 16, 32, torch.float16),       # only the background is old; one tile of 256 pixels + one of 256
    (152, 76, 16, 64, torch.bfloat16),     # every channel register used; boundary exactly between the two lanes
    (152, 80, 16, 16, torch.bfloat16),     # a single 256-pixel tile; boundary on a block edge of the upper half
    (101, 75, 32, 32, torch.bfloat16),     # ADE20K step 0 size: 51 of the upper lane's 76 registers hold -inf
    (77, 5, 16, 48, torch.float16),        # one channel in the upper half; boundary inside block 0 (next to channel 0)
    (64, 33, 16, 32, torch.bfloat16)])     # smallest K of the path: the upper lanes hold no channel at all
def test_pixel_register_column_edges(ops, K, old_cl, H, W, dtype):
    """the one-pass kernel for 64 <= K <= 152 in 16-bit storage (pixel_regs.cuh): channel halves of a pixel pair in two
    lanes, block-wise old / new sums with a rolled boundary fix-up, -inf padding above K"""
    from bacs_b200 import _cabi
    from bacs_b200.synth import StepConfig, make_labels
    B, T = 2, 3
    cfg = StepConfig("edge", B=B, K=K, old_cl=old_cl, T=T, H=H, W=W)
    g = torch.Generator().manual_seed(K * 1000 + old_cl)
    logits = (torch.randn(B, K, H, W, generator=g) * 2).to(dtype)
    mask = make_labels(cfg, g, classes=list(range(1, K)), border=1, block=4)
    mask[0, 3, 5:9] = K + 1                                         # invalid, not ignore
    clean = torch.where((mask >= K) & (mask != 255), torch.full_like(mask, 255), mask)
    z = torch.randn(B, T, H // 16, W // 16, generator=g).requires_grad_(True)
    up = O.bilinear_upsample(z, (H, W), True)
    smax = torch.sigmoid(up).max(1)[0].detach()
    x = logits.float().clone().requires_grad_(True)
    want = O.weighted_ce(x, clean, smax, old_cl, 2.0, 0.5, True)
    want.backward()
    kept = int((clean != 255).sum())
    want_f = O.focal_seen_loss(up[:, T - 1:T], clean, 2.0, None)
    want_f.backward()
    scale = 1024.0 if dtype == torch.float16 else 1.0
    out = ops.pixel_loss(logits.cuda(), mask.cuda(), _cabi.PIX_WEIGHTED_CE, want_grad=True, z=z.detach().cuda(),
                         want_distill_mask=True, old_cl=old_cl, ukd=True, grad_scale=scale, focal_head=T - 1)
    assert out["variant"] == 5
    acc = out["acc"].cpu()
    close(acc[_cabi.ACC_LOSS] / (B * H * W), want, what="loss")
    close(acc[_cabi.ACC_FOCAL] / kept, want_f, what="focal loss")
    close(out["gz"] / kept, z.grad[:, T - 1], atol=2e-5 * float(z.grad.abs().max()), what="gz")
    assert torch.equal(out["preds"].cpu(), O.argmax_first(logits.float()))
    want_g = (x.grad * scale).to(dtype).float()
    tol = 2.0 ** (-7 if dtype == torch.bfloat16 else -10)
    close(out["dlogits"].float(), want_g, atol=tol * float(want_g.abs().max()), what="dlogits")
    # element-wise: no entry further than one unit in the last place of the storage type from the oracle's rounding
    ulp = torch.maximum(want_g.abs(), torch.full_like(want_g, 2.0 ** -100)) * 2.0 ** (-7 if dtype == torch.bfloat16 else -10)
    big = want_g.abs() > 1e-3 * float(want_g.abs().max())
    assert bool((((out["dlogits"].float().cpu() - want_g).abs() <= 1.01 * ulp) | ~big).all())
    want_m = (clean == 0) & (smax > 0.5)
    diff = out["distill_mask"].cpu().bool() != want_m
    assert int((diff & ((smax - 0.5).abs() > 1e-6)).sum()) == 0 and int(diff.sum()) <= 2
    assert int(acc[_cabi.ACC_KEPT]) == kept and int(acc[_cabi.ACC_INVALID]) == 4
    assert int(acc[_cabi.ACC_BG]) == int((clean == 0).sum())
    # the same inputs through the two streaming passes (BACS_NO_REGS is read per call): same loss, same arg-max
    os.environ["BACS_NO_REGS"] = "1"
    try:
        ref = ops.pixel_loss(logits.cuda(), mask.cuda(), _cabi.PIX_WEIGHTED_CE, want_grad=True, z=z.detach().cuda(),
                             want_distill_mask=True, old_cl=old_cl, ukd=True, grad_scale=scale, focal_head=T - 1)
    finally:
        del os.environ["BACS_NO_REGS"]
    assert ref["variant"] == 4
    assert torch.equal(ref["preds"], out["preds"]) and torch.equal(ref["distill_mask"], out["distill_mask"])
    close(out["acc"][_cabi.ACC_LOSS], ref["acc"][_cabi.ACC_LOSS], what="loss against the streaming passes")


@pytest.mark.parametrize("dtype,T,D,B", [(torch.bfloat16, 11, 48, 5), (torch.float16, 6, 32, 2), (torch.bfloat16, 3, 16, 1)])
def test_proto_accumulate_tensor_core_path(ops, dtype, T, D, B):
    """16-bit features on 32-pixel chunks take the one-hot GEMM (mma.sync) with direct fp64 sums: both modes against the
    oracle, more than eight tasks (two task tiles), very unequal images (many runs split across rows of the reference's
    view), and agreement with the shared-memory-table kernel"""
    from bacs_b200.synth import StepConfig, make_labels
    K = 21
    initial = K - (T - 1)                                   # T tasks: classes 1..initial-1, then one class per task
    cfg = StepConfig("mma", B=B, K=K, old_cl=K - 1, T=T, H=128, W=256, D=D, A=16, initial_classes=initial, increment=1)
    g = torch.Generator().manual_seed(100 * T + B)
    mask = make_labels(cfg, g, classes=list(range(1, K)), border=2, block=16)
    if B > 1:
        mask[0, :, : cfg.W // 2] = 255                     # image 0 contributes far fewer pixels than the others
    pen = torch.randn(B, D, cfg.h, cfg.w, generator=g).to(dtype)
    lut = torch.from_numpy(O.class_task_lut(cfg.initial_classes, cfg.increment)).int()
    lut[lut >= cfg.T] = -1
    task, rank, n_bt, _ = ops.label_downsample_task(mask.cuda(), cfg.h, cfg.w, lut.cuda(), cfg.T)
    for mode, mname in [(0, "exact"), (1, "channel")]:
        sums, counts = ops.proto_accumulate(pen.cuda(), task, rank, n_bt, cfg.T, mode)
        want_s, want_n = O.proto_accumulate(pen.float(), mask, cfg.initial_classes, cfg.increment, cfg.T, mode=mname)
        assert torch.equal(counts.cpu().long(), want_n)
        close(sums, want_s, atol=2e-5 * float(want_s.abs().max()), what="sums " + mname)
        os.environ["BACS_NO_PROTO_MMA"] = "1"
        try:
            s_tab, n_tab = ops.proto_accumulate(pen.cuda(), task, rank, n_bt, cfg.T, mode)
        finally:
            del os.environ["BACS_NO_PROTO_MMA"]
        assert torch.equal(n_tab, counts)
        close(sums, s_tab, atol=2e-6 * float(want_s.abs().max()), what="tensor-core sums against the table kernel, " + mname)
