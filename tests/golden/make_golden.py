"""Generates tests/golden/*.npz from the UPSTREAM REFERENCE code (imported through
ref_shim from /root/reference, CPU, fp32).  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures carry the reference's own outputs to machines without the reference tree
(the GPU box): seeded inputs + loss, gradients, prototypes, counts, arg-max, seen maps,
confusion matrix, label remaps.  Nothing from the reference is copied but numbers."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_shim                      # noqa: E402
from fake_net import FakeNet         # noqa: E402
from bacs_b200 import synth          # noqa: E402

warnings.filterwarnings("ignore")


class _Accel:
    root_device = torch.device("cpu")


def ref_seen_net(ref, inp):
    cfg = inp.cfg
    bg = ref["networks.bg_detector"].BgDetector(cfg.D * 4)
    heads = torch.nn.ModuleList([bg.get_classification_head(1) for _ in range(cfg.T)])
    with torch.no_grad():
        for t, head in enumerate(heads):
            head.conv.weight.copy_(inp.head_w[t].view(1, cfg.D, 1, 1))
            head.conv.bias.copy_(inp.head_b[t].view(1))
    bg.seen_not_seen_clf = heads
    return bg


def full_step(ref, name, first_task, seed, **loss_kwargs):
    cfg = synth.CONFIGS[name]
    inp = synth.make_step_inputs(cfg, seed=seed)
    task_num = cfg.T - 1
    loss_kwargs.setdefault("bg_weighted_ce", True)
    L = ref["loss.bacs_loss"].BACSLoss(name="ref", **loss_kwargs)
    L.init_prototype_compute()
    L.set_continual_task_size(cfg.initial_classes, cfg.increment)
    for t in range(cfg.T):
        L._prototypes._init_prototypes(t, _Accel(), cfg.D)
    L._prototypes._prototypes_tensors = inp.protos.clone()
    L._prototypes._count_features = inp.counts.clone()
    L._update_task(task_num)
    L.old_classes, L.nb_current_classes = cfg.old_cl, cfg.K
    L.first_task = first_task
    L._use_der_loss = True
    L.set_device(torch.device("cpu"))
    L._init_dark_criterion(torch.device("cpu"))
    L.logit_transforms = lambda x: x
    L.on_train_batch_start(epoch=3, max_epochs=30, batch_idx=0)
    bg = ref_seen_net(ref, inp)
    net, prev = FakeNet(bg), FakeNet(bg)
    leaf = lambda t: t.clone().requires_grad_(True)
    lg, pn, na = leaf(inp.logits), leaf(inp.pen), leaf(inp.new_att)
    rp = inp.replay
    rlg, rsem = leaf(rp["logits"]), leaf(rp["sem_logits"])
    img, rimg, limg = torch.zeros(cfg.B, 3, 2, 2), torch.zeros(cfg.Br, 3, 2, 2), torch.zeros(cfg.Br, 3, 2, 2)
    net.register(img, lg, pn, [na])
    net.register(rimg, rlg, rp["pen"], [na])
    net.register_sem(limg, rsem)
    prev.register(img, inp.logits, inp.pen, [inp.old_att])
    L.prev_model = prev
    batch = {"main": [img, inp.mask.clone()], "buffer": [rimg, rp["mask"].clone()],
             "bufferlogits": [limg, rp["memory_logits"].clone(), rp["n_classes"]]}
    batch = L.preprocess_batch(batch)
    batch["main"][0], batch["buffer"][0], batch["bufferlogits"][0] = img, rimg, limg
    loss, preds = L.compute_loss(batch, net, train=True)
    loss.backward()
    head = bg.seen_not_seen_clf[task_num]
    seen = bg.get_seen_probs(inp.pen, inp.protos, bg_detect=True).detach()
    out = {
        "loss": loss.detach().numpy(), "preds": preds.numpy().astype(np.uint8),
        "protos": L.prototypes.numpy(), "counts": L._prototypes._count_features.numpy(),
        "dlogits": lg.grad.numpy(), "dnew_att": na.grad.numpy(), "dreplay_logits": rlg.grad.numpy(),
        "dreplay_sem": rsem.grad.numpy(), "dhead_w": head.conv.weight.grad.view(-1).numpy(),
        "dhead_b": head.conv.bias.grad.view(()).numpy(),
        "dpen": pn.grad.numpy() if pn.grad is not None else np.zeros(0, np.float32),
        "seen_max_before_update": seen.max(1)[0].numpy().astype(np.float32),
        "seed": np.int64(seed), "first_task": np.bool_(first_task),
    }
    return out


def train_end(ref, backfill):
    """BACSLoss.on_train_end (loss/bacs_loss.py:133-203): prototype back-fill when a task has no sample yet
    (loss/prototypes.py:92-125), then one pass over the train loader that fills the replay buffer."""
    import tempfile
    import types
    from fake_net import EndOfTaskAccelerator, EndOfTaskLoader, EndOfTaskNet, end_of_task_case
    if not hasattr(np, "Inf"):
        np.Inf = np.inf                                   # the reference predates numpy 2
    cfg, inps, images, sems, paths, tpaths = end_of_task_case(synth)
    inp = inps[0]
    task_num = cfg.T - 1
    root = tempfile.mkdtemp()
    os.environ["BACS_REF_CWD"] = root
    L = ref["loss.bacs_loss"].BACSLoss(name="ref", bg_weighted_ce=True, buffer_size=4)
    L.init_prototype_compute()
    L.set_continual_task_size(cfg.initial_classes, cfg.increment)
    for t in range(cfg.T):
        L._prototypes._init_prototypes(t, _Accel(), cfg.D)
    counts = inp.counts.clone()
    if backfill:
        counts[task_num] = 0
    L._prototypes._prototypes_tensors = inp.protos.clone()
    L._prototypes._count_features = counts
    L._update_task(task_num)
    L.old_classes, L.nb_current_classes = cfg.old_cl, cfg.K
    L.set_device(torch.device("cpu"))
    L.accelerator = _Accel()
    bg = ref_seen_net(ref, inp)
    net = EndOfTaskNet(bg, [i.logits for i in inps], sems, [i.pen for i in inps])
    loader = EndOfTaskLoader([(im.clone(), i.mask.clone()) for im, i in zip(images, inps)], paths, tpaths)
    trainer = types.SimpleNamespace(datamodule=types.SimpleNamespace(_sweep=False, debug=False))
    np.random.seed(0)
    L.on_train_end(pre_last_tasks=True, model=net, train_dataloader=loader, accelerator=EndOfTaskAccelerator("cpu"),
                   trainer=trainer)
    buf = L.buffer
    out = {"protos": L.prototypes.numpy(), "counts": np.asarray(L._prototypes._count_features.numpy(), np.float64),
           "examples": np.array(buf.dataset_map["examples"][:]), "logits": np.array(buf.dataset_map["logits"][:]),
           "labels": np.array(buf.dataset_map["labels"][:]).astype(np.uint8),
           "seen": np.array(buf.dataset_map["seen"][:]),
           "importance": np.asarray(buf.importance_score, np.float64), "scores": np.asarray(buf.scores, np.float64),
           "existing": np.asarray(buf._existing_indices), "n_classes": np.asarray(buf._logits_n_classes),
           "img_paths": np.asarray([str(buf.img_paths.get(i, "")) for i in range(4)]),
           "num_seen": np.int64(buf.num_seen_examples),
           "backfill": np.bool_(backfill)}
    return out


def label_cases(ref):
    TL = __import__("training.utils", fromlist=["TransformLabel"]).TransformLabel
    rng = np.random.RandomState(0)
    d1s, d2s, lbls, outs = [], [], [], []
    for trial in range(6):
        keys = rng.choice(35, size=14, replace=False) - 1
        d1 = {int(k): int(rng.randint(0, 20)) for k in keys}
        d2 = {int(k): int(v) for k, v in zip(rng.choice(20, 7, replace=False), rng.randint(0, 9, 7))}
        d2[255] = 255
        lbl = rng.randint(-1, 34, size=(24, 31)).astype(np.int64)
        out = TL(d1, 255, d2, 0)(torch.from_numpy(lbl.copy())).numpy()
        m1 = np.full(257, 255, np.int32)
        m2 = np.full(257, 0, np.int32)
        for k, v in d1.items():
            m1[k + 1] = v
        for k, v in d2.items():
            m2[k + 1] = v
        d1s.append(m1), d2s.append(m2), lbls.append(lbl), outs.append(out)
    return {"map1": np.stack(d1s), "map2": np.stack(d2s), "labels": np.stack(lbls), "remapped": np.stack(outs)}


def task_and_downsample(ref):
    out = {}
    for init, inc in [(16, 1), (16, 5), (11, 1), (101, 50), (15, 2)]:
        L = ref["loss.base_loss"].BaseLoss("x")
        L.set_continual_task_size(init, inc)
        out["task_%d_%d" % (init, inc)] = np.asarray(L.label_to_task_num(torch.arange(0, 255))).astype(np.int64)
    # inputs are regenerated from the seed by the tests (torch CPU generator), only outputs are stored
    for (H, W, h, w) in [(512, 512, 32, 32), (528, 528, 33, 33), (513, 513, 33, 33), (100, 75, 7, 5)]:
        g = torch.Generator().manual_seed(H * 1000 + w)
        t = torch.randint(0, 256, (1, H, W), generator=g)
        want = torch.nn.functional.interpolate(t.unsqueeze(1).double(), size=(h, w), mode="nearest").long()[:, 0]
        out["down_out_%d_%d_%d_%d" % (H, W, h, w)] = want.numpy().astype(np.uint8)
    return out


def main():
    ref = ref_shim.install()
    np.savez_compressed(os.path.join(HERE, "step_tiny.npz"), **full_step(ref, "tiny", False, 11))
    np.savez_compressed(os.path.join(HERE, "step_tiny_first_task.npz"), **full_step(ref, "tiny", True, 11))
    # plain CE + pseudo-labelled background (bacs_loss.py:65,205-210) + distill on mask == 0 alone (282-285)
    np.savez_compressed(os.path.join(HERE, "step_tiny_pseudo.npz"),
                        **full_step(ref, "tiny", False, 11, bg_weighted_ce=False, pseudo_label=True))
    np.savez_compressed(os.path.join(HERE, "labels.npz"), **label_cases(ref), **task_and_downsample(ref))
    np.savez_compressed(os.path.join(HERE, "train_end.npz"), **train_end(ref, False))
    np.savez_compressed(os.path.join(HERE, "train_end_backfill.npz"), **train_end(ref, True))
    # the reference's only known-answer vector (training/metrics.py:159-183)
    label = np.zeros((1, 4, 4), np.int64)
    pred = np.zeros((1, 4, 4), np.float32)
    label[:, :3, :3] = 1
    pred[:, -3:, -3:] = 1
    np.savez_compressed(os.path.join(HERE, "iou_known_answer.npz"), label=label, pred=pred,
                        confmat=np.array([[2, 5], [5, 4]], np.int64), iou=np.array([2.0 / 12, 4.0 / 14], np.float32))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
