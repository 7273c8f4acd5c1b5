"""Time of ops.class_distance at the ADE20K shape (B=24, D=512, 32x32, 150 classes) against torch's fp32 cdist^2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bacs_b200 import ops


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


B, D, h, w, Kc = 24, 512, 32, 32, 150
g = torch.Generator().manual_seed(0)
f = torch.randn(B, D, h, w, generator=g).to(torch.bfloat16).cuda()
c = torch.randn(Kc, D, generator=g).to(torch.bfloat16).cuda()
def graph_time(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    g_ = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_):
        for _ in range(reps):
            fn()
    return timeit(g_.replay, 20) / reps


print("class_distance (dist2 + nearest) %.1f us eager, %.1f us per launch in a CUDA graph"
      % (timeit(lambda: ops.class_distance(f, c)), graph_time(lambda: ops.class_distance(f, c))))
print("class_distance (dist2 only)      %.1f us" % timeit(lambda: ops.class_distance(f, c, want_nearest=False)))
ff, cf = f.float().permute(0, 2, 3, 1).reshape(B, h * w, D), c.float()
print("torch fp32 cdist^2               %.1f us" % timeit(lambda: torch.cdist(ff, cf.unsqueeze(0).expand(B, -1, -1)) ** 2))
flops = 2.0 * B * h * w * 160 * D
t = timeit(lambda: ops.class_distance(f, c)) * 1e-6
print("%.1f TFLOP/s on the padded 160-class GEMM; bytes: features %.1f MB, dist2 %.1f MB" % (flops / t / 1e12, f.numel() * 2 / 1e6, B * Kc * h * w * 4 / 1e6))
