"""GPU check of the tensor-core teacher-distill kernel against the FMA kernel (and the fp64 formula on small shapes).
usage: python tools/distill_tc_check.py [quick|full|time]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from bacs_b200 import ops

def fp64_ref(old, new, m, H, W):
    from oracle import bacs_oracle as O
    h, w = old.shape[-2:]
    def up(x):
        y0, y1, wy = O._lerp_table(H, h, False); x0, x1, wx = O._lerp_table(W, w, False)
        wy = wy.double().view(-1, 1); wx = wx.double()
        rows = x[..., y0, :] * (1 - wy) + x[..., y1, :] * wy
        return rows[..., x0] * (1 - wx) + rows[..., x1] * wx
    n = new.double().cpu().requires_grad_(True)
    e = (up(old.double().cpu()) ** 2 - up(n) ** 2)
    if m is not None:
        e = e * m.cpu().bool().unsqueeze(1)
    tot = torch.linalg.vector_norm(e, 2.0, dim=-1).sum()
    tot.backward()
    return float(tot), n.grad

def run(B, A, h, w, ratio, dtype, with_mask=True, near=None, ref64=False, seed=0, ratio_x=None):
    g = torch.Generator().manual_seed(seed)
    H, W = h * ratio, w * (ratio_x or ratio)
    old = torch.randn(B, A, h, w, generator=g)
    new = torch.randn(B, A, h, w, generator=g) if near is None else old + near * torch.randn(B, A, h, w, generator=g)
    old, new = old.to(dtype).cuda(), new.to(dtype).cuda()
    m = (torch.rand(B, H, W, generator=g) > 0.4).to(torch.uint8).cuda() if with_mask else None
    var = ops.distill_kernel_variant(new, (H, W))
    ops.distill_set_mode(1)
    s1, d1 = ops.teacher_distill(old, new, m, (H, W), 1.0, True)
    torch.cuda.synchronize()
    ops.distill_set_mode(2 if var else 0)
    s2, d2 = ops.teacher_distill(old, new, m, (H, W), 1.0, True)
    torch.cuda.synchronize()
    s3, _ = ops.teacher_distill(old, new, m, (H, W), 1.0, False)      # forward only
    torch.cuda.synchronize()
    ops.distill_set_mode(0)
    rl = abs(float(s1) - float(s2)) / max(abs(float(s1)), 1e-30)
    gmax = float(d1.float().abs().max())
    rg = float((d1.float() - d2.float()).abs().max()) / max(gmax, 1e-30)
    msg = "B%d A%d h%d w%d r%d %s mask%d near%s: variant %d  loss fma %.8g tc %.8g rel %.2e | grad rel-to-max %.2e | fwd-only rel %.1e" % (
        B, A, h, w, ratio, str(dtype).split(".")[-1], with_mask, near, var, float(s1), float(s2), rl, rg,
        abs(float(s3) - float(s2)) / max(abs(float(s2)), 1e-30))
    if ref64:
        t64, g64 = fp64_ref(old, new, m, H, W)
        gm = float(g64.abs().max())
        msg += " | vs fp64: loss fma %.1e tc %.1e, grad fma %.1e tc %.1e" % (
            abs(float(s1) - t64) / t64, abs(float(s2) - t64) / t64,
            float((d1.double().cpu() - g64).abs().max()) / gm, float((d2.double().cpu() - g64).abs().max()) / gm)
    print(msg, flush=True)
    return rl, rg

def timeit(B, A, h, w, ratio, dtype, mode, iters=20):
    g = torch.Generator().manual_seed(0)
    H, W = h * ratio, w * ratio
    old = torch.randn(B, A, h, w, generator=g).to(dtype).cuda()
    new = torch.randn(B, A, h, w, generator=g).to(dtype).cuda()
    m = (torch.rand(B, H, W, generator=g) > 0.4).to(torch.uint8).cuda()
    ops.distill_set_mode(mode)
    for _ in range(3):
        ops.teacher_distill(old, new, m, (H, W), 1.0, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.teacher_distill(old, new, m, (H, W), 1.0, True)
    e1.record(); torch.cuda.synchronize()
    ops.distill_set_mode(0)
    return e0.elapsed_time(e1) / iters * 1e3

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "quick"
    bf, f32, f16 = torch.bfloat16, torch.float32, torch.float16
    if what == "one":
        B, A, h, w, r = [int(v) for v in sys.argv[2:7]]
        run(B, A, h, w, r, {"bf16": bf, "f32": f32, "f16": f16}[sys.argv[7]])
        sys.exit(0)
    run(1, 16, 3, 16, 16, bf, ref64=True)
    run(1, 16, 3, 8, 16, f32, ref64=True)
    run(2, 40, 4, 16, 16, f32, ref64=True)
    run(2, 300, 5, 32, 16, bf, ref64=False)
    if what in ("full", "time"):
        run(2, 256, 8, 32, 16, f32, ref64=True)
        run(2, 128, 6, 16, 8, f16, ref64=True)
        run(1, 130, 4, 32, 16, bf, with_mask=False, ref64=True)
        run(2, 64, 5, 32, 16, f32, near=1e-3, ref64=True)
        run(3, 256, 32, 32, 16, bf)
        run(24, 256, 32, 32, 16, bf)
        run(1, 256, 7, 16, 40, bf, ref64=True)          # several row blocks per interval
        run(1, 64, 5, 32, 24, f32, ref64=True, ratio_x=16)
    if what == "time":
        for mode in (1, 2):
            print("mode %d: headline bf16 %.1f us, fp32 %.1f us" % (mode, timeit(24, 256, 32, 32, 16, bf, mode),
                                                                   timeit(24, 256, 32, 32, 16, f32, mode)), flush=True)
