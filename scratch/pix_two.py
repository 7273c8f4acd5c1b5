import sys, torch
sys.path.insert(0, '/root/repo')
from bacs_b200 import synth, ops, _cabi
cfg = synth.CONFIGS["voc15-1_b24"]
g = torch.Generator().manual_seed(0)
logits = torch.randn(cfg.B, cfg.K, cfg.H, cfg.W, generator=g).bfloat16().cuda()
mask = synth.make_labels(cfg, g).cuda()
hist = ops.label_hist(mask)
W = _cabi.PIX_WEIGHTED_CE; C = _cabi.PIX_CE
smax = torch.rand(cfg.B, cfg.H, cfg.W, generator=g).cuda()
z1 = torch.randn(cfg.B, 1, cfg.h, cfg.w, generator=g).cuda()
z6 = torch.randn(cfg.B, 6, cfg.h, cfg.w, generator=g).cuda()
for _ in range(2):
    ops.pixel_loss(logits, mask, C, want_grad=True, hist=hist)                                             # 0: ce+grad
    ops.pixel_loss(logits, mask, W, want_grad=True, seen_max=smax, want_distill_mask=True, old_cl=cfg.old_cl)  # 1: wce seen_max
    ops.pixel_loss(logits, mask, W, want_grad=True, z=z1, want_distill_mask=True, old_cl=cfg.old_cl)        # 2: wce T=1
    ops.pixel_loss(logits, mask, W, want_grad=True, z=z6, want_distill_mask=True, old_cl=cfg.old_cl)        # 3: wce T=6
    ops.pixel_loss(logits, mask, W, want_grad=True, z=z6, want_distill_mask=True, old_cl=cfg.old_cl, focal_head=5)  # 4: full
torch.cuda.synchronize()
