"""ExperienceReplay (reference loss/experience_replay.py:153-272): the replay loss, the buffer
sampling around it and the end-of-task buffer fill, against the oracle's class-weighted CE / per-image score.

The network is a fixed linear map of the image, so replayed images (new tensors every call) get well-defined logits on
both sides.  Replay sampling draws from numpy's global RNG exactly like the reference: the expected minibatch is
obtained by re-seeding and asking the same Buffer."""
import numpy as np
import pytest
import torch

from oracle import bacs_oracle as O

pytestmark = pytest.mark.gpu

K, HW = 7, (32, 48)


class LinearNet(torch.nn.Module):
    def __init__(self, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.weight = torch.nn.Parameter(torch.randn(K, 3, generator=g))
        self.bias = torch.nn.Parameter(torch.randn(K, generator=g))

    def forward(self, x, return_attentions=False, return_penultimate=False, return_sem_logits=False, only_attentions=False):
        logits = torch.einsum("bchw,kc->bkhw", x, self.weight) + self.bias.view(1, -1, 1, 1)
        pen = torch.nn.functional.avg_pool2d(x, 16)
        if return_penultimate and return_attentions:
            return logits, pen, [pen]
        return logits


class Accel:
    root_device = torch.device("cuda")

    def process_dataloader(self, loader):
        return loader

    def to_device(self, batch):
        return [t.cuda() for t in batch]


def images_labels(n, seed):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(n, 3, *HW, generator=g)
    lab = torch.randint(0, K, (n, *HW), generator=g)
    lab[:, :2] = 255
    return img, lab


def make_loss(tmp_path, monkeypatch, **kw):
    monkeypatch.setenv("BACS_BUFFER_ROOT", str(tmp_path))
    from bacs_b200.loss import ExperienceReplay
    L = ExperienceReplay(alpha=0.7, buffer_size=6, replay_minibatch_size=3, **kw)
    L.set_continual_task_size(3, 2)
    L.nb_current_classes, L.old_classes = K, 5
    L.set_device(torch.device("cuda"))
    L.accelerator = Accel()
    L.on_train_batch_start(epoch=1, max_epochs=30, batch_idx=0)
    return L


def fill(L, seed, n=8):
    img, lab = images_labels(n, seed)
    L._add_to_buffer(img, lab, -torch.rand(n, generator=torch.Generator().manual_seed(seed)))


def test_er_loss_is_ce_plus_replayed_old_class_ce(tmp_path, monkeypatch):
    L = make_loss(tmp_path, monkeypatch)
    L.on_train_start(0)
    assert isinstance(L.buffer, list) and len(L.buffer) == 1 and not L._use_er_loss
    fill(L, seed=1)
    L.update_buffer_scores()
    net = LinearNet().cuda()
    img, lab = images_labels(4, seed=2)
    loss0, preds0 = L.compute_loss([img.cuda(), lab.cuda()], net, train=True)      # first task: plain CE
    want0 = O.cross_entropy(net(img.cuda()).detach().cpu(), lab)
    assert abs(float(loss0) - float(want0)) <= 1e-5 * float(want0)
    assert torch.equal(preds0.cpu(), O.argmax_first(net(img.cuda()).detach().cpu()))

    L.on_train_start(1)
    assert len(L.buffer) == 2 and L._use_er_loss
    np.random.seed(11)
    loss, preds = L.compute_loss([img.cuda(), lab.cuda()], net, train=True)
    net.zero_grad()
    loss.backward()
    got_gw, got_gb = net.weight.grad.clone().cpu(), net.bias.grad.clone().cpu()
    # the same draw from the only earlier buffer
    np.random.seed(11)
    mem = L.buffer[0].get_data(3, device=None)
    cpu = LinearNet()
    class_w = torch.zeros(K)
    assert L.buffer[0].task_num == 0
    class_w[1:L.get_n_old_classes(0 + 1)] = 1                  # the classes known after the replayed buffer's task
    want = O.cross_entropy(cpu(img), lab) + 0.7 * 0.7 * O.cross_entropy(cpu(mem["examples"]), mem["labels"].long(), class_w)
    want.backward()
    assert abs(float(loss) - float(want)) <= 1e-5 * float(want), (float(loss), float(want))
    assert float((got_gw - cpu.weight.grad).abs().max()) <= 3e-5 * float(cpu.weight.grad.abs().max())
    assert float((got_gb - cpu.bias.grad).abs().max()) <= 3e-5 * float(cpu.bias.grad.abs().max())
    # evaluation: no replay term
    with torch.no_grad():
        ev, _ = L.compute_loss([img.cuda(), lab.cuda()], net, train=False)
    assert abs(float(ev) - float(want0)) <= 1e-5 * float(want0)


def test_replay_source_and_minibatch(tmp_path, monkeypatch):
    L = make_loss(tmp_path, monkeypatch)
    for t in range(4):
        L.on_train_start(t)
        if t < 3:
            fill(L, seed=10 + t, n=5 + t)
            L.update_buffer_scores()
    # three earlier buffers: softmax(importance / max importance), numpy's global RNG
    imp = np.array([b.get_importance() for b in L.buffer[:-1]], dtype=np.float64)
    s = imp / imp.max()
    p = np.exp(s - s.max())
    p /= p.sum()
    np.random.seed(3)
    want = [int(np.random.choice(range(3), p=p, size=1)[0]) for _ in range(20)]
    np.random.seed(3)
    got = [L.buffer.index(L._get_random_buffer()) for _ in range(20)]
    assert got == want and len(set(got)) > 1
    # a minibatch: the reference's 6-tuple; the ER buffer stores no logits
    np.random.seed(5)
    d, ex, logits, labels, ncls, task = L._sample_buffer(L.buffer[1])
    assert ex.shape == (3, 3, *HW) and ex.is_cuda and labels.shape == (3, *HW) and logits is None
    assert task == L.buffer[1].task_num == 1 and len(ncls) == 3
    assert L._sample_buffer(L.buffer[3]) is None                                  # the current task's buffer is empty
    assert list(L.get_available_tasks()) == [0, 1, 2, 3]


def test_end_of_task_fills_the_buffer_with_scored_images(tmp_path, monkeypatch):
    L = make_loss(tmp_path, monkeypatch)
    L.on_train_start(0)
    net = LinearNet().cuda()
    batches = [images_labels(4, seed=20 + i) for i in range(4)]
    seen = []
    orig = L._add_to_buffer

    def spy(examples, labels, losses):
        seen.append((examples.cpu(), labels.cpu(), losses.cpu()))
        orig(examples, labels, losses)
    monkeypatch.setattr(L, "_add_to_buffer", spy)
    L.on_train_end(model=net, train_dataloader=batches, accelerator=Accel(), pre_last_tasks=True)
    assert len(seen) == 3                                   # stops once index * batch >= buffer_size (6): batches 0, 1, 2
    cpu = LinearNet()
    class_w = torch.ones(K)
    class_w[0] = 0
    for (img, lab), (ex, lb, score) in zip(batches, seen):
        want = O.cross_entropy_per_image_score(cpu(img).detach(), lab, class_w)
        assert torch.equal(ex, img) and torch.equal(lb, lab)
        assert float((score - want).abs().max()) <= 1e-5 * float(want.abs().max())
    buf = L._get_current_buffer()
    assert buf.num_seen_examples == 12 and np.isclose(buf.scores.sum(), 1.0)
    # without pre_last_tasks nothing is touched
    L2 = make_loss(tmp_path / "b", monkeypatch)
    (tmp_path / "b").mkdir(exist_ok=True)
    L2.on_train_start(0)
    L2.on_train_end(model=net, train_dataloader=batches, accelerator=Accel(), pre_last_tasks=False)
    assert L2._get_current_buffer().is_empty()
