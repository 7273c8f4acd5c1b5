import sys, torch
sys.path.insert(0, '/root/repo')
from bacs_b200 import synth, ops, _cabi
cfg = synth.CONFIGS["voc15-1_b24"]
dt = torch.bfloat16
g = torch.Generator().manual_seed(0)
logits = torch.randn(cfg.B, cfg.K, cfg.H, cfg.W, generator=g).to(dt).cuda()
mask = synth.make_labels(cfg, g).cuda()
hist = ops.label_hist(mask)
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
W = _cabi.PIX_WEIGHTED_CE; C = _cabi.PIX_CE
smax = torch.rand(cfg.B, cfg.H, cfg.W, generator=g).cuda()
for T in (1, 2, 6, 11):
    z = torch.randn(cfg.B, T, cfg.h, cfg.w, generator=g).cuda()
    print("wce+grad T=%d: %.1f us" % (T, timeit(lambda: ops.pixel_loss(logits, mask, W, want_grad=True, z=z, want_distill_mask=True, old_cl=cfg.old_cl))))
print("wce+grad seen_max tensor: %.1f us" % timeit(lambda: ops.pixel_loss(logits, mask, W, want_grad=True, seen_max=smax, want_distill_mask=True, old_cl=cfg.old_cl)))
z = torch.randn(cfg.B, 6, cfg.h, cfg.w, generator=g).cuda()
print("ce+grad with z (focal only, T=6): %.1f us" % timeit(lambda: ops.pixel_loss(logits, mask, C, want_grad=True, z=z, focal_head=5, hist=hist)))
print("ce+grad no z: %.1f us" % timeit(lambda: ops.pixel_loss(logits, mask, C, want_grad=True, hist=hist)))
print("uce+grad: %.1f us" % timeit(lambda: ops.pixel_loss(logits, mask, _cabi.PIX_UNBIASED_CE, want_grad=True, hist=hist, old_cl=cfg.old_cl)))
