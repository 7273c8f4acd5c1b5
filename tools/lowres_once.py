"""Runs the fused low-res pixel kernel a few times (ncu target)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from bacs_b200 import synth, ops, _cabi
cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "voc15-1_b24"]
inp = synth.make_step_inputs(cfg, seed=0, dtype=torch.bfloat16, device="cuda")
z = torch.randn(cfg.B, cfg.T, cfg.h, cfg.w, device="cuda")
sem = torch.randn(cfg.B, cfg.K, cfg.h, cfg.w, device="cuda").to(torch.bfloat16)
for _ in range(4):
    ops.pixel_loss(sem, inp.mask, _cabi.PIX_WEIGHTED_CE, lowres=True, want_grad=True, z=z, want_distill_mask=True,
                   old_cl=cfg.old_cl, focal_head=cfg.T - 1)
torch.cuda.synchronize()
