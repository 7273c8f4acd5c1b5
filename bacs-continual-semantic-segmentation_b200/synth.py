"""Synthetic VOC / ADE / Cityscapes-shaped inputs for the BACS loss path.

Shapes follow BASELINE.json's configs; values are seeded.  There is no network access,
so no dataset or checkpoint is ever read: network outputs (logits, penultimate
features, attentions) are random tensors of the architecture's shapes."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch


@dataclass
class StepConfig:
    name: str
    B: int
    K: int                 # nb_current_classes (logit channels)
    old_cl: int            # classes known before this task (incl. background)
    T: int                 # task prototypes so far (incl. the current task)
    H: int
    W: int
    D: int = 512           # penultimate (bg-detector) width
    A: int = 256           # attention channels
    initial_classes: int = 16
    increment: int = 1
    Br: int = 0            # replay batch (0 = no replay in the step)
    stride: int = 16

    @property
    def h(self):
        return self.H // self.stride

    @property
    def w(self):
        return self.W // self.stride

    @property
    def pixels(self):
        return self.B * self.H * self.W


CONFIGS = {
    # BASELINE.json configs[0]: the reference's CPU-runnable case
    "voc15-1_cpu": StepConfig("voc15-1_cpu", B=8, K=21, old_cl=20, T=6, H=528, W=528, D=256, A=256),
    # configs[1]: the headline single-GPU workload
    "voc15-1_b24": StepConfig("voc15-1_b24", B=24, K=21, old_cl=20, T=6, H=512, W=512, D=512, A=256),
    # configs[2]: VOC 10-1 with DER replay
    "voc10-1_der": StepConfig("voc10-1_der", B=24, K=21, old_cl=20, T=11, H=512, W=512, D=512, A=256,
                              initial_classes=11, increment=1, Br=24),
    # configs[3]: ADE20K 100-50
    "ade100-50": StepConfig("ade100-50", B=24, K=151, old_cl=101, T=2, H=512, W=512, D=512, A=256,
                            initial_classes=101, increment=50),
    # configs[4]: Cityscapes
    "cityscapes": StepConfig("cityscapes", B=12, K=20, old_cl=18, T=3, H=512, W=1024, D=512, A=256,
                             initial_classes=15, increment=2),
    # small cases for parity tests / smoke
    "tiny": StepConfig("tiny", B=2, K=7, old_cl=5, T=3, H=64, W=96, D=32, A=16, initial_classes=3,
                       increment=2, Br=2),
    "small": StepConfig("small", B=3, K=21, old_cl=20, T=6, H=128, W=160, D=64, A=32, Br=3),
    # 512-pixel row tiles (the training-step kernel): whole rows, two tiles per row, padded class counts
    "row512": StepConfig("row512", B=2, K=21, old_cl=20, T=6, H=48, W=512, D=32, A=16, Br=2),
    "row1024": StepConfig("row1024", B=1, K=20, old_cl=18, T=3, H=32, W=1024, D=32, A=16, initial_classes=15,
                          increment=2),
    "row512_k17": StepConfig("row512_k17", B=2, K=17, old_cl=16, T=2, H=32, W=512, D=32, A=16),
    "row512_k11": StepConfig("row512_k11", B=1, K=11, old_cl=6, T=11, H=32, W=512, D=32, A=16, initial_classes=6,
                             increment=5),
    # 3-stage ring of the training-step kernel: many seen heads / 24 register rows
    "row512_t11": StepConfig("row512_t11", B=1, K=21, old_cl=20, T=11, H=32, W=512, D=32, A=16, initial_classes=11,
                             increment=1),
    "row512_k24": StepConfig("row512_k24", B=1, K=24, old_cl=20, T=3, H=32, W=512, D=32, A=16, initial_classes=20,
                             increment=2),
    "wide1536": StepConfig("wide1536", B=1, K=5, old_cl=3, T=2, H=32, W=1536, D=16, A=12, initial_classes=3, increment=2),
    # more classes than fit in registers: the shared-memory kernel with cooperating lane pairs (ADE20K-like)
    "row512_k40": StepConfig("row512_k40", B=2, K=40, old_cl=31, T=2, H=32, W=512, D=32, A=16, initial_classes=31,
                             increment=9),
    # ADE20K-sized class count: 128-pixel tiles in 128-thread CTAs
    "row512_k151": StepConfig("row512_k151", B=1, K=151, old_cl=101, T=2, H=32, W=512, D=32, A=16, initial_classes=101,
                              increment=50),
    "row512_k7": StepConfig("row512_k7", B=1, K=7, old_cl=5, T=3, H=32, W=512, D=32, A=16, initial_classes=3,
                            increment=2),
}


def make_labels(cfg: StepConfig, gen: torch.Generator, device="cpu", classes=None, B: Optional[int] = None,
                p_bg: float = 0.6, border: int = 8, block: int = 16) -> torch.Tensor:
    """[B,H,W] int64: ~60 % background, an ignore-255 border (+ a few ignore blobs), the
    rest uniform over ``classes`` (default: the current task's classes old_cl..K-1)."""
    B = cfg.B if B is None else B
    if classes is None:
        classes = list(range(cfg.old_cl, cfg.K)) or [1]
    classes = torch.tensor(classes, dtype=torch.int64)
    gh, gw = (cfg.H + block - 1) // block, (cfg.W + block - 1) // block
    u = torch.rand(B, gh, gw, generator=gen)
    pick = classes[torch.randint(len(classes), (B, gh, gw), generator=gen)]
    coarse = torch.where(u < p_bg, torch.zeros_like(pick), pick)
    coarse = torch.where(u > 0.985, torch.full_like(pick, 255), coarse)      # ignore blobs
    lab = coarse.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :cfg.H, :cfg.W].clone()
    # per-pixel speckle so neighbouring pixels differ inside blocks
    sp = torch.rand(B, cfg.H, cfg.W, generator=gen)
    lab = torch.where(sp < 0.02, torch.zeros_like(lab), lab)
    if border > 0:
        lab[:, :border] = 255
        lab[:, -border:] = 255
        lab[:, :, :border] = 255
        lab[:, :, -border:] = 255
    return lab.to(device)


@dataclass
class StepInputs:
    cfg: StepConfig
    logits: torch.Tensor
    pen: torch.Tensor
    old_att: torch.Tensor
    new_att: torch.Tensor
    mask: torch.Tensor
    protos: torch.Tensor
    counts: torch.Tensor
    head_w: torch.Tensor           # [T,D]
    head_b: torch.Tensor           # [T]
    replay: Optional[dict] = field(default=None)


def make_step_inputs(cfg: StepConfig, seed: int = 0, dtype=torch.float32, device="cpu",
                     with_replay: Optional[bool] = None) -> StepInputs:
    """Seeded inputs; features are drawn in fp32 then rounded to ``dtype`` so the fp32
    oracle can be fed exactly the same (e.g. bf16-rounded) values."""
    gen = torch.Generator().manual_seed(seed)

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=gen) * scale).to(dtype).to(device)

    logits = rnd(cfg.B, cfg.K, cfg.H, cfg.W)
    pen = rnd(cfg.B, cfg.D, cfg.h, cfg.w)
    old_att = rnd(cfg.B, cfg.A, cfg.h, cfg.w)
    new_att = rnd(cfg.B, cfg.A, cfg.h, cfg.w)
    mask = make_labels(cfg, gen, device)
    protos = torch.randn(cfg.T, cfg.D, generator=gen).to(device)
    counts = torch.full((cfg.T,), 1000.0).to(device)
    head_w = (torch.randn(cfg.T, cfg.D, generator=gen) * (2.0 / cfg.D ** 0.5)).to(device)
    head_b = (torch.randn(cfg.T, generator=gen) * 0.1).to(device)
    replay = None
    if with_replay is None:
        with_replay = cfg.Br > 0
    if with_replay and cfg.Br > 0:
        old_classes = list(range(1, cfg.old_cl))
        replay = {
            "logits": rnd(cfg.Br, cfg.K, cfg.H, cfg.W),
            "pen": rnd(cfg.Br, cfg.D, cfg.h, cfg.w),
            "mask": make_labels(cfg, gen, device, classes=old_classes, B=cfg.Br),
            "sem_logits": rnd(cfg.Br, cfg.K, cfg.h, cfg.w),
            "memory_logits": rnd(cfg.Br, cfg.K, cfg.h, cfg.w, scale=2.0),
            "n_classes": torch.randint(cfg.initial_classes, cfg.K + 1, (cfg.Br,), generator=gen,
                                       dtype=torch.int64).to(torch.uint8).to(device),
        }
    return StepInputs(cfg, logits, pen, old_att, new_att, mask, protos, counts, head_w, head_b, replay)


# --------------------------------------------------------------------------------------
# A stand-in network that satisfies the contract the loss relies on by
# returning fixed tensors, so that only the loss path runs (the DeepLabV3 forward /
# backward is out of scope for this path).
# --------------------------------------------------------------------------------------
class HeadHolder(torch.nn.Module):
    """Parameter layout of the reference's classification_head: conv.weight [1,D,1,1], conv.bias [1]."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor):
        super().__init__()
        D = weight.numel()
        self.conv = torch.nn.Conv2d(D, 1, 1)
        with torch.no_grad():
            self.conv.weight.copy_(weight.reshape(1, D, 1, 1))
            self.conv.bias.copy_(bias.reshape(1))
        self.stop_gradients = False


class SeenHeads(torch.nn.Module):
    def __init__(self, head_w: torch.Tensor, head_b: torch.Tensor):
        super().__init__()
        self.inter_channels = head_w.shape[1]
        self.seen_not_seen_clf = torch.nn.ModuleList([HeadHolder(head_w[t], head_b[t]) for t in range(head_w.shape[0])])
        self.stop_gradients = False

    def set_stop_gradients(self, stop):
        if stop == self.stop_gradients:
            return
        self.stop_gradients = stop
        for h in self.seen_not_seen_clf:
            h.stop_gradients = stop


class FixedOutputNetwork(torch.nn.Module):
    """model(img, return_penultimate=True, return_attentions=True) -> (logits, pen, [att]);
    model(img, return_sem_logits=True) -> low-res logits; keyed by the image tensor's id."""

    def __init__(self, seen_fg_network=None):
        super().__init__()
        self.seen_fg_network = seen_fg_network
        self._full, self._sem = {}, {}

    def register(self, img, logits, pen, atts):
        self._full[id(img)] = (logits, pen, atts)

    def register_sem(self, img, sem):
        self._sem[id(img)] = sem

    def forward(self, x, return_attentions=False, return_penultimate=False, return_sem_logits=False,
                only_attentions=False):
        if return_sem_logits:
            sem = self._sem[id(x)]
            if return_penultimate and return_attentions:     # networks/deeplab_v3.py:155-172 with all three flags
                _, pen, atts = self._full[id(x)]
                return sem, pen, atts
            return sem
        logits, pen, atts = self._full[id(x)]
        if return_penultimate and return_attentions:
            return logits, pen, atts
        if return_penultimate:
            return logits, pen
        if return_attentions:
            return logits, atts
        return logits

    def get_penultimate_layer_dim(self):
        return self.seen_fg_network.inter_channels


class _Accelerator:
    def __init__(self, device):
        self.root_device = torch.device(device)


def build_bacs_step(cfg: StepConfig, inp: StepInputs, device="cuda", first_task: bool = False, epoch: int = 3,
                    max_epochs: int = 30, exact_prototypes: bool = True, **loss_kwargs):
    """Wires a BACSLoss + FixedOutputNetwork for one training step of the given config.
    Returns (loss_fn, network, batch, leaves) where ``leaves`` are the tensors that receive
    gradients (logits, new_att, replay logits, replay sem logits, focal head params)."""
    from .loss import BACSLoss

    dev = torch.device(device)
    task_num = cfg.T - 1
    loss_kwargs.setdefault("bg_weighted_ce", True)
    loss_fn = BACSLoss(name="bacs", **loss_kwargs)
    loss_fn.init_prototype_compute()
    loss_fn._prototypes.exact = exact_prototypes
    loss_fn.set_continual_task_size(cfg.initial_classes, cfg.increment)
    loss_fn._update_task(task_num)
    loss_fn.old_classes, loss_fn.nb_current_classes = cfg.old_cl, cfg.K
    loss_fn.first_task = first_task
    loss_fn._use_der_loss = True
    loss_fn.set_device(dev)
    loss_fn.accelerator = _Accelerator(dev)
    loss_fn.on_train_batch_start(epoch=epoch, max_epochs=max_epochs, batch_idx=0)
    P = loss_fn._prototypes
    P._prototypes_tensors = inp.protos.clone().to(dev)
    P._count_features = inp.counts.clone().to(dev)
    P.refresh_ready()
    heads = SeenHeads(inp.head_w, inp.head_b).to(dev)
    net, prev = FixedOutputNetwork(heads), FixedOutputNetwork(heads)

    def leaf(t):
        return t.clone().to(dev).requires_grad_(True)

    logits, pen, new_att = leaf(inp.logits), leaf(inp.pen), leaf(inp.new_att)
    img = torch.zeros(cfg.B, 3, 2, 2, device=dev)
    net.register(img, logits, pen, [new_att])
    prev.register(img, inp.logits.to(dev), inp.pen.to(dev), [inp.old_att.to(dev)])
    loss_fn.prev_model = prev
    leaves = {"logits": logits, "pen": pen, "new_att": new_att, "head_w": heads.seen_not_seen_clf[task_num].conv.weight,
              "head_b": heads.seen_not_seen_clf[task_num].conv.bias}
    mask = inp.mask.clone().to(dev)
    if inp.replay is not None:
        rp = inp.replay
        rimg, limg = torch.zeros(cfg.Br, 3, 2, 2, device=dev), torch.zeros(cfg.Br, 3, 2, 2, device=dev)
        rlogits, rsem = leaf(rp["logits"]), leaf(rp["sem_logits"])
        net.register(rimg, rlogits, rp["pen"].to(dev), [new_att])
        net.register_sem(limg, rsem)
        batch = {"main": [img, mask], "buffer": [rimg, rp["mask"].clone().to(dev)],
                 "bufferlogits": [limg, rp["memory_logits"].float().clone().to(dev), rp["n_classes"].to(dev)]}
        batch = loss_fn.preprocess_batch(batch)
        # preprocess_batch makes new image tensors (.float()); keep the registered ones
        batch["main"][0], batch["buffer"][0], batch["bufferlogits"][0] = img, rimg, limg
        leaves.update(replay_logits=rlogits, replay_sem=rsem)
    else:
        loss_fn.alpha = loss_fn.beta = 0.0
        batch = [img, mask]
    return loss_fn, net, batch, leaves
