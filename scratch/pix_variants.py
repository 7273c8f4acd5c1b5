import sys, torch
sys.path.insert(0, '/root/repo')
from bacs_b200 import synth, ops, _cabi
cfg = synth.CONFIGS["voc15-1_b24"]
dt = torch.bfloat16 if len(sys.argv) < 2 else {"bf16": torch.bfloat16, "fp32": torch.float32}[sys.argv[1]]
g = torch.Generator().manual_seed(0)
logits = torch.randn(cfg.B, cfg.K, cfg.H, cfg.W, generator=g).to(dt).cuda()
mask = synth.make_labels(cfg, g).cuda()
z = torch.randn(cfg.B, cfg.T, cfg.h, cfg.w, generator=g).cuda()
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
variants = {
 "full (wce+focal+mask+grad)": lambda: ops.pixel_loss(logits, mask, W, want_grad=True, z=z, want_distill_mask=True, focal_head=cfg.T-1, old_cl=cfg.old_cl),
 "wce+mask+grad (no focal)": lambda: ops.pixel_loss(logits, mask, W, want_grad=True, z=z, want_distill_mask=True, old_cl=cfg.old_cl),
 "wce no grad": lambda: ops.pixel_loss(logits, mask, W, want_grad=False, z=z, want_distill_mask=True, focal_head=cfg.T-1, old_cl=cfg.old_cl),
 "ce+grad (no z)": lambda: ops.pixel_loss(logits, mask, C, want_grad=True, hist=hist),
 "ce no grad (read+argmax)": lambda: ops.pixel_loss(logits, mask, C, want_grad=False),
 "ce no grad no preds": lambda: ops.pixel_loss(logits, mask, C, want_grad=False, want_preds=False),
}
px = cfg.B * cfg.H * cfg.W
es = logits.element_size()
for name, fn in variants.items():
    us = timeit(fn)
    print("%-32s %8.1f us   %.2f Gpx/s" % (name, us, px / us / 1e3))
# plain copy for reference
dst = torch.empty_like(logits)
us = timeit(lambda: dst.copy_(logits))
print("%-32s %8.1f us   %.0f GB/s" % ("torch copy logits", us, 2 * logits.numel() * es / us / 1e3))
