"""GPU parity of the whole BACS step through the reference-facing class interface
(BACSLoss.compute_loss + backward) against the CPU oracle's bacs_step on the same seeded
inputs: loss, arg-max, prototypes / counts, and every gradient."""
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
    err = float((got - want).abs().max())
    assert torch.allclose(got, want, rtol=rtol, atol=atol), "%s max abs err %.3e (atol %.3e)" % (what, err, atol)


def _oracle_step(cfg, inp, first_task, proto_mode="exact"):
    leaf = lambda t: t.float().clone().requires_grad_(True)
    lg, pn, na = leaf(inp.logits), leaf(inp.pen), leaf(inp.new_att)
    hw, hb = leaf(inp.head_w), leaf(inp.head_b)
    rp2 = None
    leaves = {"logits": lg, "pen": pn, "new_att": na, "head_w": hw, "head_b": hb}
    if inp.replay is not None:
        rp = inp.replay
        rlg, rsem = leaf(rp["logits"]), leaf(rp["sem_logits"])
        rp2 = dict(rp, logits=rlg, sem_logits=rsem, pen=rp["pen"].float(), n_classes=rp["n_classes"].numpy(),
                   memory_logits=rp["memory_logits"].float())
        leaves.update(replay_logits=rlg, replay_sem=rsem)
    out = O.bacs_step(lg, pn, inp.old_att.float(), na, inp.mask, inp.protos, inp.counts, hw, hb,
                      initial_classes=cfg.initial_classes, increment=cfg.increment, old_cl=cfg.old_cl,
                      task_num=cfg.T - 1, first_task=first_task, epoch=3, max_epochs=30, replay=rp2,
                      nb_current_classes=cfg.K, proto_mode=proto_mode)
    out["loss"].backward()
    return out, leaves


@pytest.mark.parametrize("name,first_task,dtype", [("tiny", False, torch.float32), ("tiny", True, torch.float32),
                                                   ("small", False, torch.float32), ("small", False, torch.bfloat16),
                                                   ("row512", False, torch.float32), ("row512", False, torch.bfloat16),
                                                   ("row512", True, torch.float32)])
def test_full_step_matches_oracle(name, first_task, dtype):
    from bacs_b200 import synth
    cfg = synth.CONFIGS[name]
    inp = synth.make_step_inputs(cfg, seed=11, dtype=dtype)
    want, wl = _oracle_step(cfg, inp, first_task)
    loss_fn, net, batch, leaves = synth.build_bacs_step(cfg, inp, first_task=first_task)
    loss, preds = loss_fn.compute_loss(batch, net, train=True)
    loss.backward()
    close(loss, want["loss"], what="loss")
    assert torch.equal(preds.cpu(), want["preds"])
    close(loss_fn.prototypes, want["protos"], what="prototypes")
    close(loss_fn._prototypes._count_features, want["counts"], what="counts")
    tol = 1e-5 if dtype == torch.float32 else 2.0 ** -7
    for key in ("logits", "new_att", "replay_logits", "replay_sem"):
        g, w = leaves[key].grad, wl[key].grad.to(dtype).float()
        close(g.float(), w, atol=(3 * tol if key == "new_att" else tol) * float(w.abs().max()), what="d" + key)
    t = cfg.T - 1
    close(leaves["head_w"].grad.reshape(-1), wl["head_w"].grad[t], atol=3e-5 * float(wl["head_w"].grad[t].abs().max()),
          what="dhead_w")
    close(leaves["head_b"].grad.reshape(()), wl["head_b"].grad[t], atol=3e-5 * float(wl["head_b"].grad[t].abs()),
          what="dhead_b")
    if first_task:
        w = wl["pen"].grad
        close(leaves["pen"].grad.float(), w, atol=3e-5 * float(w.abs().max()), what="dpen")
    else:
        assert leaves["pen"].grad is None or float(leaves["pen"].grad.abs().max()) == 0.0


def test_eval_step_and_confusion_matrix():
    from bacs_b200 import synth
    from bacs_b200.training.metrics import IoU
    cfg = synth.CONFIGS["small"]
    inp = synth.make_step_inputs(cfg, seed=5)
    loss_fn, net, batch, leaves = synth.build_bacs_step(cfg, inp)
    with torch.no_grad():
        loss, preds = loss_fn.compute_loss(batch, net, train=False)
    close(loss, O.cross_entropy(inp.logits, inp.mask), what="eval CE")
    assert torch.equal(preds.cpu(), O.argmax_first(inp.logits))
    iou = IoU(num_classes=cfg.K).cuda()
    iou(preds, batch["main"][1])
    want = O.confusion_matrix(preds.cpu().numpy(), inp.mask.numpy(), cfg.K)
    assert np.array_equal(iou.confmat.cpu().numpy(), want)
    met = iou.compute()
    wm = O.iou_metrics(want)
    assert np.allclose(met.iou_per_class.cpu().numpy(), wm["iou_per_class"], rtol=1e-6, atol=1e-7)
    assert np.allclose(float(met.miou), wm["miou"], rtol=1e-6)
    iou.reset()
    assert int(iou.confmat.sum()) == 0


def test_step_is_free_of_host_sync():
    """The training step must be capturable in a CUDA graph (no host sync, no allocation
    outside the caching allocator): capture it, replay it, compare with the eager result."""
    from bacs_b200 import synth
    cfg = synth.CONFIGS["small"]
    inp = synth.make_step_inputs(cfg, seed=2)
    loss_fn, net, batch, leaves = synth.build_bacs_step(cfg, inp)
    protos0 = loss_fn._prototypes._prototypes_tensors.clone()
    counts0 = loss_fn._prototypes._count_features.clone()

    def step():
        for v in leaves.values():
            v.grad = None
        loss, preds = loss_fn.compute_loss(batch, net, train=True)
        loss.backward()
        return loss, preds

    # PyTorch's capture recipe: warm up on a side stream, capture, replay -- and only then run
    # the eager reference on the default stream (an eager backward on the default stream before
    # the capture ties the leaves' AccumulateGrad nodes to it and invalidates the capture)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        gl, gp = step()
    loss_fn._prototypes._prototypes_tensors.copy_(protos0)
    loss_fn._prototypes._count_features.copy_(counts0)
    g.replay()
    torch.cuda.synchronize()
    graph_loss, graph_preds, graph_grad = float(gl), gp.clone(), leaves["logits"].grad.clone()
    loss_fn._prototypes._prototypes_tensors.copy_(protos0)
    loss_fn._prototypes._count_features.copy_(counts0)
    eager_loss, eager_preds = step()
    eager_grad = leaves["logits"].grad.clone()
    assert graph_loss == float(eager_loss)
    assert torch.equal(graph_preds, eager_preds)
    assert torch.equal(graph_grad, eager_grad)
