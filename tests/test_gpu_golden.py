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


@pytest.mark.parametrize("fixture,first_task", [("step_tiny.npz", False), ("step_tiny_first_task.npz", True)])
def test_step_matches_reference_fixture(fixture, first_task):
    from bacs_b200 import synth
    gold = np.load(os.path.join(GOLD, fixture))
    cfg = synth.CONFIGS["tiny"]
    inp = synth.make_step_inputs(cfg, seed=int(gold["seed"]))
    loss_fn, net, batch, leaves = synth.build_bacs_step(cfg, inp, first_task=first_task)
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
