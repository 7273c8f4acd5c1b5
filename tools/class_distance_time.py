"""Times the tensor-core pixel x class-prototype distance at the ADE20K shape (B=24, D=512, 32x32, 150 classes)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from bacs_b200 import ops
B, D, h, w, Kc = 24, 512, 32, 32, 150
f = torch.randn(B, D, h, w, device="cuda").to(torch.bfloat16)
c = torch.randn(Kc, D, device="cuda").to(torch.bfloat16)
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
t = timeit(lambda: ops.class_distance(f, c))
def torch_ref():
    x = f.float().permute(0, 2, 3, 1).reshape(-1, D)
    return torch.cdist(x, c.float()) ** 2
t2 = timeit(torch_ref)
flop = 2.0 * B * h * w * Kc * D
print("class_distance B=%d D=%d %dx%d Kc=%d: %.1f us (%.1f TFLOP/s on the dot products; bytes %.1f MB) | torch fp32 cdist**2 %.1f us"
      % (B, D, h, w, Kc, t, flop / t / 1e6, (f.numel() * 2 * 2 + B * Kc * h * w * 4 * 3) / 1e6, t2))
