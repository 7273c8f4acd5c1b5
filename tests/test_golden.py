"""CPU tests: the oracle against the committed golden fixtures (generated from the upstream
reference by tests/golden/make_golden.py) and the reference's only known-answer vector."""
import os

import numpy as np
import torch

from oracle import bacs_oracle as O
from bacs_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def close(got, want, rtol=1e-5, atol=None):
    got = torch.as_tensor(np.asarray(got)).double()
    want = torch.as_tensor(np.asarray(want)).double()
    if atol is None:
        atol = rtol * max(1e-30, float(want.abs().max()))
    assert torch.allclose(got, want, rtol=rtol, atol=atol), float((got - want).abs().max())


def oracle_step(name, first_task, seed):
    cfg = synth.CONFIGS[name]
    inp = synth.make_step_inputs(cfg, seed=seed)
    leaf = lambda t: t.clone().requires_grad_(True)
    lv = {"logits": leaf(inp.logits), "pen": leaf(inp.pen), "new_att": leaf(inp.new_att),
          "head_w": leaf(inp.head_w), "head_b": leaf(inp.head_b)}
    rp = inp.replay
    lv["replay_logits"], lv["replay_sem"] = leaf(rp["logits"]), leaf(rp["sem_logits"])
    rp2 = dict(rp, logits=lv["replay_logits"], sem_logits=lv["replay_sem"], n_classes=rp["n_classes"].numpy())
    out = O.bacs_step(lv["logits"], lv["pen"], inp.old_att, lv["new_att"], inp.mask, inp.protos, inp.counts,
                      lv["head_w"], lv["head_b"], initial_classes=cfg.initial_classes, increment=cfg.increment,
                      old_cl=cfg.old_cl, task_num=cfg.T - 1, first_task=first_task, epoch=3, max_epochs=30,
                      replay=rp2, nb_current_classes=cfg.K)
    out["loss"].backward()
    return cfg, inp, out, lv


def check_step(fixture, first_task):
    gold = load(fixture)
    cfg, inp, out, lv = oracle_step("tiny", first_task, int(gold["seed"]))
    close(out["loss"].detach(), gold["loss"])
    assert np.array_equal(out["preds"].numpy(), gold["preds"].astype(np.int64))
    close(out["protos"], gold["protos"])
    close(out["counts"], gold["counts"])
    close(lv["logits"].grad, gold["dlogits"])
    close(lv["new_att"].grad, gold["dnew_att"])
    close(lv["replay_logits"].grad, gold["dreplay_logits"])
    close(lv["replay_sem"].grad, gold["dreplay_sem"])
    t = cfg.T - 1
    close(lv["head_w"].grad[t], gold["dhead_w"])
    close(lv["head_b"].grad[t], gold["dhead_b"], atol=1e-7)
    if first_task:
        close(lv["pen"].grad, gold["dpen"])
    close(O.seen_max(inp.pen, inp.protos, inp.head_w, inp.head_b), gold["seen_max_before_update"])


def test_oracle_step_matches_reference_fixture():
    check_step("step_tiny.npz", False)


def test_oracle_first_task_step_matches_reference_fixture():
    check_step("step_tiny_first_task.npz", True)


def test_oracle_labels_match_reference_fixture():
    gold = load("labels.npz")
    for i in range(gold["labels"].shape[0]):
        d1 = {v - 1: int(gold["map1"][i][v]) for v in range(257)}
        d2 = {v - 1: int(gold["map2"][i][v]) for v in range(257)}
        got = O.transform_label(gold["labels"][i], d1, 255, d2, 0)
        assert np.array_equal(got, gold["remapped"][i])
    for key in gold.files:
        if key.startswith("task_"):
            _, init, inc = key.split("_")
            assert np.array_equal(O.class_to_task(np.arange(255), int(init), int(inc)), gold[key])
        if key.startswith("down_out_"):
            H, W, h, w = (int(v) for v in key.split("_")[2:])
            g = torch.Generator().manual_seed(H * 1000 + w)
            t = torch.randint(0, 256, (1, H, W), generator=g)
            assert np.array_equal(O.downsample_labels(t, h, w).numpy(), gold[key].astype(np.int64))


def test_iou_known_answer_vector():
    gold = load("iou_known_answer.npz")
    C = O.confusion_matrix(gold["pred"], gold["label"], 2)
    assert np.array_equal(C, gold["confmat"])
    assert np.allclose(O.iou_metrics(C)["iou_per_class"], gold["iou"], atol=1e-6)
