"""Whole-step parity at BASELINE.json's GPU configs: ``BACSLoss.compute_loss`` + backward (the call the training loop
makes) against the oracle's ``bacs_step`` evaluated on the SAME full-size inputs -- the oracle's torch formulas run on
device tensors in fp32 (the CPU oracle needs minutes and tens of GB of host memory at B=24 512^2; its formulas are the
same file, oracle/bacs_oracle.py).  The teacher-distill term of the oracle is evaluated four images at a time (it is a
mean over images: its [B,A,H,W] up-sampled maps are 6.4 GB each at B=24).

Tolerances: loss, prototypes 1e-5 relative; arg-max and counts bit-exact; fp32-stored gradients 1e-5 (3e-5 for the
attention / head gradients, sums of ~1e5 terms) of the largest entry; gradients stored in bf16 are compared with the
oracle's gradient rounded to bf16 -- by max-norm AND element by element on a sampled 1 % (1 bf16 ulp of the value,
plus the north_star's 1e-5 of the largest entry for values that cancel to ~0)."""
import pytest
import torch

from oracle import bacs_oracle as O

pytestmark = pytest.mark.gpu


def close(got, want, rtol=1e-5, atol=None, what=""):
    got = torch.as_tensor(got).detach().double().cpu()
    want = torch.as_tensor(want).detach().double().cpu()
    if atol is None:
        atol = rtol * max(1e-30, float(want.abs().max()))
    err = float((got - want).abs().max())
    assert torch.allclose(got, want, rtol=rtol, atol=atol), "%s max abs err %.3e (atol %.3e)" % (what, err, atol)


def ulp_check(got, want_f32, dtype, what, frac=0.01, seed=0):
    """element-wise on a random sample: |got - round(want)| <= 1 ulp of the storage type (relative to the value)"""
    ref = want_f32.to(dtype).float().reshape(-1)
    g = got.float().reshape(-1)
    n = ref.numel()
    gen = torch.Generator(device=ref.device).manual_seed(seed)
    idx = torch.randint(0, n, (max(1, int(n * frac)),), device=ref.device, generator=gen)
    r, x = ref[idx], g[idx]
    ulp = 2.0 ** (-7 if dtype == torch.bfloat16 else -10)
    tol = ulp * r.abs() + 1e-5 * float(ref.abs().max())
    bad = int(((x - r).abs() > tol).sum())
    assert bad == 0, "%s: %d of %d sampled elements differ by more than 1 ulp" % (what, bad, idx.numel())


def device_oracle_step(cfg, inp, dev, chunk=4):
    leaf = lambda t: t.float().to(dev).clone().requires_grad_(True)
    lg, pn, na = leaf(inp.logits), leaf(inp.pen), leaf(inp.new_att)
    hw, hb = leaf(inp.head_w), leaf(inp.head_b)
    mask = inp.mask.to(dev)
    leaves = {"logits": lg, "pen": pn, "new_att": na, "head_w": hw, "head_b": hb}
    rp2 = None
    if inp.replay is not None:
        rp = inp.replay
        rlg, rsem = leaf(rp["logits"]), leaf(rp["sem_logits"])
        rp2 = dict(rp, logits=rlg, sem_logits=rsem, pen=rp["pen"].float().to(dev), mask=rp["mask"].to(dev),
                   n_classes=rp["n_classes"].numpy(), memory_logits=rp["memory_logits"].float().to(dev))
        leaves.update(replay_logits=rlg, replay_sem=rsem)
    out = O.bacs_step(lg, pn, inp.old_att.float().to(dev), na, mask, inp.protos.to(dev), inp.counts.to(dev), hw, hb,
                      initial_classes=cfg.initial_classes, increment=cfg.increment, old_cl=cfg.old_cl,
                      task_num=cfg.T - 1, first_task=False, epoch=3, max_epochs=30, replay=rp2,
                      nb_current_classes=cfg.K, lkd=0.0)
    out["loss"].backward()
    total = out["loss"].detach().double()
    old = inp.old_att.float().to(dev)
    for b0 in range(0, cfg.B, chunk):                   # teacher distill: mean over images, a few images at a time
        b1 = min(cfg.B, b0 + chunk)
        part = O.teacher_distill(old[b0:b1], na[b0:b1], mask[b0:b1], out["seen_max"][b0:b1], 0.25, 0.5) * ((b1 - b0) / cfg.B)
        part.backward()
        total += part.detach().double()
    out["loss"] = total
    return out, leaves


@pytest.mark.parametrize("name,dtype", [("voc15-1_b24", torch.bfloat16), ("voc10-1_der", torch.bfloat16),
                                        ("cityscapes", torch.bfloat16), ("ade100-50", torch.bfloat16),
                                        ("voc15-1_b24", torch.float32)])
def test_full_size_step_matches_device_oracle(name, dtype):
    from bacs_b200 import synth
    dev = torch.device("cuda")
    torch.cuda.empty_cache()
    cfg = synth.CONFIGS[name]
    inp = synth.make_step_inputs(cfg, seed=21, dtype=dtype)
    loss_fn, net, batch, leaves = synth.build_bacs_step(cfg, inp)
    loss, preds = loss_fn.compute_loss(batch, net, train=True)
    loss.backward()
    torch.cuda.synchronize()
    got = {"loss": float(loss), "preds": preds.clone(), "protos": loss_fn.prototypes.clone(),
           "counts": loss_fn._prototypes._count_features.clone(),
           "grads": {k: (v.grad.clone() if v.grad is not None else None) for k, v in leaves.items()}}
    del loss_fn, net, batch, leaves, loss, preds
    torch.cuda.empty_cache()
    want, wl = device_oracle_step(cfg, inp, dev)
    assert abs(got["loss"] - float(want["loss"])) <= 1e-5 * abs(float(want["loss"])), (got["loss"], float(want["loss"]))
    assert torch.equal(got["preds"], want["preds"])
    close(got["protos"], want["protos"], what="prototypes")
    assert torch.equal(got["counts"].double(), want["counts"].double())
    tol = 1e-5 if dtype == torch.float32 else 2.0 ** -7
    # pixels whose seen probability sits within fp32 rounding of the 0.5 threshold may take the other branch of the
    # weighted CE (s > thr -> 1): a few in 10^5 (the synthetic heads put the seen probability close to 0.5); they are excluded from the logit-gradient comparison
    band = (want["seen_max"] - 0.5).abs() < 2e-6
    assert int(band.sum()) <= 1e-4 * band.numel()
    for key in ("logits", "new_att", "replay_logits", "replay_sem"):
        if key not in wl:
            continue
        g, w = got["grads"][key], wl[key].grad
        if key == "logits":
            keep = (~band).unsqueeze(1)
            g, w = g * keep, w * keep
        if key == "new_att":   # the distill mask takes the same threshold: low-res cells under a band pixel (+ neighbours)
            r = cfg.H // cfg.h
            cells = torch.nn.functional.max_pool2d(band.float().unsqueeze(1), r)
            cells = torch.nn.functional.max_pool2d(cells, 3, stride=1, padding=1) > 0
            g, w = g * ~cells, w * ~cells
        wr = w.to(dtype).float()
        close(g.float(), wr, atol=(3 * tol if key == "new_att" else tol) * float(wr.abs().max()), what="d" + key)
        if dtype != torch.float32:
            ulp_check(g, w, dtype, "d" + key)
    t = cfg.T - 1
    close(got["grads"]["head_w"].reshape(-1), wl["head_w"].grad[t], atol=3e-5 * float(wl["head_w"].grad[t].abs().max()),
          what="dhead_w")
    close(got["grads"]["head_b"].reshape(()), wl["head_b"].grad[t], atol=3e-5 * float(wl["head_b"].grad[t].abs()),
          what="dhead_b")
