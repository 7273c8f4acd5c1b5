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
