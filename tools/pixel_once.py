"""Runs the full-resolution fused pixel kernel a few times at a named config (ncu target)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from bacs_b200 import synth, ops, _cabi
cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "voc15-1_b24"]
dtype = {"bf16": torch.bfloat16, "f32": torch.float32}[sys.argv[2] if len(sys.argv) > 2 else "bf16"]
B = int(sys.argv[3]) if len(sys.argv) > 3 else cfg.B
g = torch.Generator(device="cuda").manual_seed(0)
logits = torch.randn(B, cfg.K, cfg.H, cfg.W, device="cuda", generator=g).to(dtype)
mask = synth.make_labels(cfg, torch.Generator().manual_seed(1), classes=list(range(1, cfg.K)), B=B).cuda()
z = torch.randn(B, cfg.T, cfg.h, cfg.w, device="cuda")
for _ in range(4):
    ops.pixel_loss(logits, mask, _cabi.PIX_WEIGHTED_CE, want_grad=True, z=z, want_distill_mask=True,
                   old_cl=cfg.old_cl, focal_head=cfg.T - 1)
torch.cuda.synchronize()
