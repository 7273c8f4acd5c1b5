"""GPU parity against the committed golden fixtures -- the upstream reference's own outputs
(tests/golden/make_golden.py) -- through the reference-facing classes and the C ABI."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def close(got, want, rtol=1e-5, atol=None, what=""):
    got = torch.as_tensor(np.asarray(got.detach().cpu() if isinstance(got, torch.Tensor) else got)).double()
    want = torch.as_tensor(np.asarray(want)).double()
    if atol is None:
        atol = rtol * max(1e-30, float(want.abs().max()))
    assert torch.allclose(got, want, rtol=rtol, atol=atol), "%s %.3e" % (what, float((got - want).abs().max()))


@pytest.mark.parametrize("fixture,first_task,kwargs", [
    ("step_tiny.npz", False, {}), ("step_tiny_first_task.npz", True, {}),
    ("step_tiny_pseudo.npz", False, {"bg_weighted_ce": False, "pseudo_label": True})])
def test_step_matches_reference_fixture(fixture, first_task, kwargs):
    from bacs_b200 import synth
    gold = np.load(os.path.join(GOLD, fixture))
    cfg = synth.CONFIGS["tiny"]
    inp = synth.make_step_inputs(cfg, seed=int(gold["seed"]))
    loss_fn, net, batch, leaves = synth.build_bacs_step(cfg, inp, first_task=first_task, **kwargs)
    loss, preds = loss_fn.compute_loss(batch, net, train=True)
    loss.backward()
    close(loss, gold["loss"], what="loss")
    assert np.array_equal(preds.cpu().numpy(), gold["preds"].astype(np.int64))
    close(loss_fn.prototypes, gold["protos"], what="prototypes")
    close(loss_fn._prototypes._count_features, gold["counts"], what="counts")
    close(leaves["logits"].grad, gold["dlogits"], what="dlogits")
    close(leaves["new_att"].grad, gold["dnew_att"], rtol=3e-5, what="dnew_att")
    close(leaves["replay_logits"].grad, gold["dreplay_logits"], what="dreplay_logits")
    close(leaves["replay_sem"].grad, gold["dreplay_sem"], what="dreplay_sem")
    close(leaves["head_w"].grad.reshape(-1), gold["dhead_w"], rtol=3e-5, what="dhead_w")
    close(leaves["head_b"].grad.reshape(()), gold["dhead_b"], rtol=3e-5, what="dhead_b")
    if first_task:
        close(leaves["pen"].grad, gold["dpen"], rtol=3e-5, what="dpen")


def test_labels_match_reference_fixture():
    from bacs_b200 import ops
    gold = np.load(os.path.join(GOLD, "labels.npz"))
    for i in range(gold["labels"].shape[0]):
        got = ops.label_remap(torch.from_numpy(gold["labels"][i][None]).cuda(), torch.from_numpy(gold["map1"][i]).cuda(),
                              255, torch.from_numpy(gold["map2"][i]).cuda(), 0, lo=-1)
        assert np.array_equal(got.cpu().numpy()[0], gold["remapped"][i])
    for key in gold.files:
        if key.startswith("down_out_"):
            H, W, h, w = (int(v) for v in key.split("_")[2:])
            g = torch.Generator().manual_seed(H * 1000 + w)
            t = torch.randint(0, 256, (1, H, W), generator=g)
            lut = torch.full((256,), -1, dtype=torch.int32).cuda()
            _, _, _, down = ops.label_downsample_task(t.cuda(), h, w, lut, 1, want_labels_down=True)
            assert np.array_equal(down.cpu().numpy(), gold[key].astype(np.int64))


def test_iou_known_answer_vector():
    from bacs_b200.training.metrics import IoU
    gold = np.load(os.path.join(GOLD, "iou_known_answer.npz"))
    iou = IoU(num_classes=2).cuda()
    iou(torch.from_numpy(gold["pred"]).cuda(), torch.from_numpy(gold["label"]).cuda())
    assert np.array_equal(iou.confmat.cpu().numpy(), gold["confmat"])
    assert np.allclose(iou.compute().iou_per_class.cpu().numpy(), gold["iou"], atol=1e-6)


@pytest.mark.parametrize("fixture,backfill", [("train_end.npz", False), ("train_end_backfill.npz", True)])
def test_end_of_task_matches_reference_fixture(fixture, backfill, tmp_path, monkeypatch):
    """SURVEY 8a row 16 (+ 8f-3): BACSLoss.on_train_end -- the prototype back-fill over the train loader when a task has
    no sample yet (loss/prototypes.py:92-125), the per-image importance scores (bacs_loss.py:183-189), the seen map of
    the current head and the reservoir insert into the replay buffer (training/buffer.py:205-270) -- against the buffer
    the upstream code produced from the same seeded loader; run twice: reference data flow and fused_logit_upsample."""
    import sys
    import types
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fake_net import EndOfTaskAccelerator, EndOfTaskLoader, EndOfTaskNet, end_of_task_case
    from bacs_b200 import synth
    from bacs_b200.loss import BACSLoss
    gold = np.load(os.path.join(GOLD, fixture))
    cfg, inps, images, sems, paths, tpaths = end_of_task_case(synth)
    inp = inps[0]
    task_num = cfg.T - 1
    dev = torch.device("cuda")
    for fused in (False, True):
        root = tmp_path / ("fused" if fused else "plain")
        root.mkdir()
        monkeypatch.setenv("BACS_BUFFER_ROOT", str(root))
        L = BACSLoss(name="bacs", bg_weighted_ce=True, buffer_size=4, fused_logit_upsample=fused)
        L.init_prototype_compute()
        L.set_continual_task_size(cfg.initial_classes, cfg.increment)
        L._update_task(task_num)
        L.old_classes, L.nb_current_classes = cfg.old_cl, cfg.K
        L.set_device(dev)
        L.accelerator = EndOfTaskAccelerator(dev)
        counts = inp.counts.clone()
        if backfill:
            counts[task_num] = 0
        L._prototypes._prototypes_tensors = inp.protos.clone().to(dev)
        L._prototypes._count_features = counts.to(dev)
        L._prototypes.refresh_ready()
        heads = synth.SeenHeads(inp.head_w, inp.head_b).to(dev)
        # the fused data flow scores the up-sampled sem_logits; give it the low-res logits whose up-sample the plain
        # flow sees (the reference's model(images)), i.e. keep `logits` consistent with `sems` for that run only
        if fused:
            from oracle import bacs_oracle as O
            logits = [O.upsample_sem_logits(s, (cfg.H, cfg.W)) for s in sems]
        else:
            logits = [i.logits for i in inps]
        net = EndOfTaskNet(heads, logits, sems, [i.pen for i in inps])
        loader = EndOfTaskLoader([(im.clone(), i.mask.clone()) for im, i in zip(images, inps)], paths, tpaths)
        trainer = types.SimpleNamespace(datamodule=types.SimpleNamespace(_sweep=False, debug=False))
        np.random.seed(0)
        L.on_train_end(pre_last_tasks=True, model=net, train_dataloader=loader, accelerator=EndOfTaskAccelerator(dev),
                       trainer=trainer)
        buf = L.buffer
        assert int(buf.num_seen_examples) == int(gold["num_seen"])
        assert np.array_equal(np.asarray(buf._existing_indices), gold["existing"])
        assert np.array_equal(np.asarray(buf._logits_n_classes), gold["n_classes"])
        assert [str(buf.img_paths.get(i, "")) for i in range(4)] == list(gold["img_paths"])
        close(L.prototypes, gold["protos"], what="prototypes")
        close(L._prototypes._count_features, gold["counts"], what="counts")
        assert np.array_equal(np.array(buf.dataset_map["examples"][:]), gold["examples"])
        assert np.array_equal(np.array(buf.dataset_map["logits"][:]), gold["logits"])
        assert np.array_equal(np.array(buf.dataset_map["labels"][:]).astype(np.uint8), gold["labels"])
        close(np.array(buf.dataset_map["seen"][:]), gold["seen"], what="seen map")
        if not fused:       # (the fused run scores different logits by construction: checked against the oracle below)
            close(np.asarray(buf.importance_score), gold["importance"], what="importance")
            close(np.asarray(buf.scores), gold["scores"], rtol=1e-4, what="scores")
        else:
            from oracle import bacs_oracle as O
            w = torch.ones(cfg.K)
            w[0] = 0
            # slots hold samples 0, 1, 5, 3 of the loader (reservoir with np.random.seed(0), see img_paths)
            order = [int(p.split("_")[1].split(".")[0]) for p in gold["img_paths"]]
            want = []
            for idx in order:
                bi, bj = divmod(idx, cfg.B)
                sc = O.cross_entropy_per_image_score(logits[bi][bj:bj + 1], inps[bi].mask[bj:bj + 1], w)
                want.append(float(sc[0]))
            close(np.asarray(buf.importance_score), np.asarray(want), what="importance (fused)")
