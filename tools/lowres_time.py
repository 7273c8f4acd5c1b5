"""Times the fused low-res pixel kernel (bacs_pixel_loss_lowres) next to the full-resolution one."""
import sys, torch
sys.path.insert(0, '/root/repo')
from bacs_b200 import synth, ops, _cabi
from oracle import bacs_oracle as O
name = sys.argv[1] if len(sys.argv) > 1 else "voc15-1_b24"
dtype = {"bf16": torch.bfloat16, "f32": torch.float32}[sys.argv[2] if len(sys.argv) > 2 else "bf16"]
cfg = synth.CONFIGS[name]
inp = synth.make_step_inputs(cfg, seed=0, dtype=dtype, device="cuda")
z = torch.randn(cfg.B, cfg.T, cfg.h, cfg.w, device="cuda")
sem = torch.randn(cfg.B, cfg.K, cfg.h, cfg.w, device="cuda").to(dtype)
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1000
kw = dict(want_grad=True, z=z, want_distill_mask=True, old_cl=cfg.old_cl, focal_head=cfg.T - 1)
t_full = timeit(lambda: ops.pixel_loss(inp.logits, inp.mask, _cabi.PIX_WEIGHTED_CE, **kw))
t_low = timeit(lambda: ops.pixel_loss(sem, inp.mask, _cabi.PIX_WEIGHTED_CE, lowres=True, **kw))
t_low_fwd = timeit(lambda: ops.pixel_loss(sem, inp.mask, _cabi.PIX_WEIGHTED_CE, lowres=True, want_grad=False, z=z, old_cl=cfg.old_cl))
def torch_path():
    s = sem.detach().requires_grad_(True)
    full = torch.nn.functional.interpolate(s, size=(cfg.H, cfg.W), mode="bilinear", align_corners=False)
    out = ops.pixel_loss(full.detach(), inp.mask, _cabi.PIX_WEIGHTED_CE, **kw)
    full.backward(out["dlogits"])
t_torch = timeit(torch_path, 10)
px = cfg.B * cfg.H * cfg.W
print("%s %s: full-res kernel %.1f us | torch up-sample fwd+bwd + full-res kernel %.1f us | fused low-res %.1f us "
      "(fwd only %.1f us) = %.2f Gpx/s" % (name, dtype, t_full, t_torch, t_low, t_low_fwd, px / t_low / 1e3))
