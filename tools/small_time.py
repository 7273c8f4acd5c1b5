"""Times the small kernels of the step one by one at the headline config (bf16)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from bacs_b200 import synth, ops
cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "voc15-1_b24"]
inp = synth.make_step_inputs(cfg, seed=0, dtype=torch.bfloat16, device="cuda")
pen, protos, hw, hb, mask = inp.pen, inp.protos.float(), inp.head_w.float(), inp.head_b.float(), inp.mask
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(n):
        big.zero_()                                  # flush L2 between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1000
lut = torch.zeros(256, dtype=torch.int32, device="cuda")
ncls = cfg.K
for c in range(1, ncls):
    lut[c] = 0 if c < cfg.initial_classes else 1 + (c - cfg.initial_classes) // cfg.increment
lut[0] = -1; lut[255] = -1
task, rank, n_bt, _ = ops.label_downsample_task(mask, cfg.h, cfg.w, lut, cfg.T)
print("downsample_task   %.1f us" % timeit(lambda: ops.label_downsample_task(mask, cfg.h, cfg.w, lut, cfg.T)))
print("proto_accumulate  %.1f us (accumulate + finalize)" % timeit(lambda: ops.proto_accumulate(pen, task, rank, n_bt, cfg.T)))
print("seen_logits       %.1f us" % timeit(lambda: ops.seen_logits(pen, protos, hw, hb)))
gz = torch.randn(cfg.B, cfg.h, cfg.w, device="cuda")
sc = torch.ones(1, device="cuda")
print("seen_head_bwd     %.1f us" % timeit(lambda: ops.seen_head_backward(pen, protos[-1], hw[-1], gz, sc, True)))
print("label_hist        %.1f us" % timeit(lambda: ops.label_hist(mask)))
