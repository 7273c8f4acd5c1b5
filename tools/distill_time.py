import sys, torch
sys.path.insert(0, '/root/repo')
from bacs_b200 import synth, ops
cfg = synth.CONFIGS["voc15-1_b24"]
g = torch.Generator().manual_seed(0)
old = torch.randn(cfg.B, cfg.A, cfg.h, cfg.w, generator=g).bfloat16().cuda()
new = torch.randn(cfg.B, cfg.A, cfg.h, cfg.w, generator=g).bfloat16().cuda()
mask = ((synth.make_labels(cfg, g) == 0) & (torch.rand(cfg.B, cfg.H, cfg.W, generator=g) > 0.4)).to(torch.uint8).cuda()
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
print("distill fwd+bwd %.1f us" % timeit(lambda: ops.teacher_distill(old, new, mask, (cfg.H, cfg.W), 1e-3, True)))
print("distill fwd only %.1f us" % timeit(lambda: ops.teacher_distill(old, new, mask, (cfg.H, cfg.W), 1e-3, False)))
