"""One BACS training step of the UNMODIFIED reference (loss/bacs_loss.py:212-256 + backward) on the CPU, fed the same
synthetic network outputs as bench.py's own arm.  TEST INFRASTRUCTURE: used by `bench.py --impl reference` /
`cpu_baseline` and by the parity tests; never imported by the product package.

The reference modules come from oracle/_ref/ (staged by oracle/make_ref.py) or from /root/reference; third-party
imports the hot path does not execute (pytorch_lightning, hydra) and segmentation_models_pytorch's FocalLoss are
stubbed by tests/golden/ref_shim.py exactly as for the golden fixtures."""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def reference_root():
    for cand in (os.environ.get("BACS_REFERENCE_ROOT"), os.path.join(HERE, "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "loss")):
            return cand
    return None


def available() -> bool:
    return reference_root() is not None


def load_reference():
    root = reference_root()
    if root is None:
        raise RuntimeError("no reference tree: run oracle/make_ref.py in the build container")
    os.environ["BACS_REFERENCE_ROOT"] = root
    for p in (os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tests"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ref_shim
    ref_shim.REFERENCE_ROOT = root
    return ref_shim.install()


class _Accel:
    def __init__(self, torch):
        self.root_device = torch.device("cpu")


def build_step(cfg, inp, first_task=False, **loss_kwargs):
    """-> (step, leaves): step() runs BACSLoss.compute_loss(batch, net, train=True) + backward of the reference."""
    import warnings
    import torch
    warnings.filterwarnings("ignore")
    ref = load_reference()
    from fake_net import FakeNet
    task_num = cfg.T - 1
    loss_kwargs.setdefault("bg_weighted_ce", True)
    L = ref["loss.bacs_loss"].BACSLoss(name="ref", **loss_kwargs)
    L.init_prototype_compute()
    L.set_continual_task_size(cfg.initial_classes, cfg.increment)
    for t in range(cfg.T):
        L._prototypes._init_prototypes(t, _Accel(torch), cfg.D)
    L._update_task(task_num)
    L.old_classes, L.nb_current_classes = cfg.old_cl, cfg.K
    L.first_task = first_task
    L._use_der_loss = True
    L.set_device(torch.device("cpu"))
    L._init_dark_criterion(torch.device("cpu"))
    L.logit_transforms = lambda x: x
    L.on_train_batch_start(epoch=3, max_epochs=30, batch_idx=0)
    bg = ref["networks.bg_detector"].BgDetector(cfg.D * 4)
    heads = torch.nn.ModuleList([bg.get_classification_head(1) for _ in range(cfg.T)])
    with torch.no_grad():
        for t, head in enumerate(heads):
            head.conv.weight.copy_(inp.head_w[t].view(1, cfg.D, 1, 1))
            head.conv.bias.copy_(inp.head_b[t].view(1))
    bg.seen_not_seen_clf = heads
    net, prev = FakeNet(bg), FakeNet(bg)
    img = torch.zeros(cfg.B, 3, 2, 2)
    prev.register(img, inp.logits.float(), inp.pen.float(), [inp.old_att.float()])
    L.prev_model = prev
    state = {}

    def step():
        leaf = lambda t: t.float().clone().requires_grad_(True)
        lg, pn, na = leaf(inp.logits), leaf(inp.pen), leaf(inp.new_att)
        L._prototypes._prototypes_tensors = inp.protos.clone()
        L._prototypes._count_features = inp.counts.clone()
        net.register(img, lg, pn, [na])
        if inp.replay is not None:
            rp = inp.replay
            rimg, limg = torch.zeros(cfg.Br, 3, 2, 2), torch.zeros(cfg.Br, 3, 2, 2)
            rlg, rsem = leaf(rp["logits"]), leaf(rp["sem_logits"])
            net.register(rimg, rlg, rp["pen"].float(), [na])
            net.register_sem(limg, rsem)
            batch = {"main": [img, inp.mask.clone()], "buffer": [rimg, rp["mask"].clone()],
                     "bufferlogits": [limg, rp["memory_logits"].clone(), rp["n_classes"]]}
            batch = L.preprocess_batch(batch)
            batch["main"][0], batch["buffer"][0], batch["bufferlogits"][0] = img, rimg, limg
        else:
            L.alpha = L.beta = 0.0
            batch = L.preprocess_batch([img, inp.mask.clone()])
            batch[0] = img
        loss, preds = L.compute_loss(batch, net, train=True)
        loss.backward()
        state.update(loss=loss.detach(), preds=preds, dlogits=lg.grad, dnew_att=na.grad,
                     protos=L.prototypes.detach().clone(), counts=L._prototypes._count_features.clone())
        return state

    return step, state


def time_step(cfg, inp, steps=1, warmup=0):
    """-> (seconds per step, last state)"""
    step, state = build_step(cfg, inp)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / max(1, steps), state
