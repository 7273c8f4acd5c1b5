"""GPU parity tests of the tensor-core teacher-distill kernel (csrc/distill_tc.cu, loss/bacs_loss.py:258-294).

The kernel is checked against (i) the CPU oracle (fp32 restatement of the reference), (ii) the oracle's formula
evaluated in fp64 -- the yard-stick for the split-tf32 error budget: 1e-5 relative on the loss, 3e-5 of the largest
gradient entry (north_star tolerance; measured ~1e-6) -- and (iii) the packed-fp32 kernel of csrc/distill.cu on the
same inputs.  Structural properties: old == new gives exactly zero, results do not depend on how the work is cut
into CTA ranges (per-image calls reproduce the batched gradient bit for bit), forward-only == forward of fwd+bwd."""
import pytest
import torch

from oracle import bacs_oracle as O

pytestmark = pytest.mark.gpu

F32, BF16, F16 = torch.float32, torch.bfloat16, torch.float16


@pytest.fixture(scope="module")
def ops():
    from bacs_b200 import ops as _ops
    yield _ops
    _ops.distill_set_mode(0)


def fp64_reference(old, new, m, H, W):
    """Row norms of m * (U(old)^2 - U(new)^2) with the oracle's fp32 interpolation tables, evaluated in fp64."""
    h, w = old.shape[-2:]

    def up(x):
        y0, y1, wy = O._lerp_table(H, h, False)
        x0, x1, wx = O._lerp_table(W, w, False)
        wy, wx = wy.double().view(-1, 1), wx.double()
        rows = x[..., y0, :] * (1 - wy) + x[..., y1, :] * wy
        return rows[..., x0] * (1 - wx) + rows[..., x1] * wx
    n = new.double().cpu().requires_grad_(True)
    e = up(old.double().cpu()) ** 2 - up(n) ** 2
    if m is not None:
        e = e * m.cpu().bool().unsqueeze(1)
    tot = torch.linalg.vector_norm(e, 2.0, dim=-1).sum()
    tot.backward()
    return float(tot.detach()), n.grad


def make(B, A, h, w, ry, rx, dtype, with_mask, near=None, seed=0):
    g = torch.Generator().manual_seed(seed)
    H, W = h * ry, w * rx
    old = torch.randn(B, A, h, w, generator=g)
    new = torch.randn(B, A, h, w, generator=g) if near is None else old + near * torch.randn(B, A, h, w, generator=g)
    m = (torch.rand(B, H, W, generator=g) > 0.4).to(torch.uint8).cuda() if with_mask else None
    return old.to(dtype).cuda(), new.to(dtype).cuda(), m, H, W


CASES = [
    # B, A, h, w, ry, rx, dtype, mask
    (1, 16, 3, 16, 16, 16, BF16, True),      # one channel block, mostly idle lanes
    (1, 16, 3, 8, 16, 16, F32, True),        # 32-byte rows, two moment buffers
    (2, 40, 4, 16, 16, 16, F32, True),
    (2, 300, 5, 32, 16, 16, BF16, True),     # two channel blocks, the second one partial
    (2, 256, 8, 32, 16, 16, F32, True),      # 128-byte rows: two attention-row slots
    (2, 128, 6, 16, 8, 8, F16, True),        # x8 up-sampling
    (1, 130, 4, 32, 16, 16, BF16, False),    # no mask
    (1, 256, 7, 16, 40, 16, BF16, True),     # three row blocks per source-row interval
    (1, 64, 5, 32, 24, 16, F32, True),       # two row blocks per interval, anisotropic
    (3, 256, 32, 32, 16, 16, BF16, True),    # VOC-sized maps: CTA ranges start inside images (boundary rows)
]


@pytest.mark.parametrize("B,A,h,w,ry,rx,dtype,with_mask", CASES)
def test_tensor_core_distill_matches_fp64_and_fma(ops, B, A, h, w, ry, rx, dtype, with_mask):
    old, new, m, H, W = make(B, A, h, w, ry, rx, dtype, with_mask)
    assert ops.distill_kernel_variant(new, (H, W)) == 1, "shape must be served by the tensor-core kernel"
    ops.distill_set_mode(2)
    s_tc, d_tc = ops.teacher_distill(old, new, m, (H, W), 1.0, True)
    s_fwd, _ = ops.teacher_distill(old, new, m, (H, W), 1.0, False)
    ops.distill_set_mode(1)
    s_fma, d_fma = ops.teacher_distill(old, new, m, (H, W), 1.0, True)
    ops.distill_set_mode(0)
    assert float(s_fwd) == float(s_tc), "forward-only must reproduce the forward of fwd+bwd"
    want, wgrad = fp64_reference(old, new, m, H, W)
    assert abs(float(s_tc) - want) <= 1e-5 * want, (float(s_tc), want)
    assert abs(float(s_tc) - float(s_fma)) <= 1e-5 * want
    gmax = float(wgrad.abs().max())
    if dtype == F32:
        assert float((d_tc.double().cpu() - wgrad).abs().max()) <= 3e-5 * gmax
    else:
        # 16-bit gradients: the fp64 gradient rounded to the storage type, element by element within 1 ulp of that type
        # (ulp of a value x: 2^-7 |x| for bf16, 2^-10 |x| for fp16 -- two roundings of nearby fp32 values can land on
        # adjacent representable numbers)
        ref = wgrad.to(dtype).float()
        got = d_tc.float().cpu()
        ulp = 2.0 ** (-7 if dtype == BF16 else -10)
        assert bool(((got - ref).abs() <= ulp * ref.abs() + 3e-5 * gmax).all())
    assert float((d_tc.float() - d_fma.float()).abs().max()) <= (3e-5 if dtype == F32 else 2.0 ** -7) * gmax


def test_tensor_core_distill_against_cpu_oracle(ops):
    """the fp32 oracle (torch on the CPU, autograd gradient) on a VOC-shaped slice"""
    old, new, m, H, W = make(2, 128, 8, 32, 16, 16, F32, True, seed=3)
    lab = torch.where(m.cpu().bool(), 0, 1)
    n = new.cpu().clone().requires_grad_(True)
    want = O.teacher_distill(old.cpu(), n, lab, None, lkd=0.25)
    want.backward()
    coef = 0.25 / (2 * 128 * H)
    ops.distill_set_mode(2)
    s, d = ops.teacher_distill(old, new, m, (H, W), coef, True)
    ops.distill_set_mode(0)
    assert abs(float(s) * coef - float(want)) <= 1e-5 * float(want)
    assert float((d.cpu() - n.grad).abs().max()) <= 3e-5 * float(n.grad.abs().max())


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_tensor_core_distill_identical_and_near_identical_maps(ops, dtype):
    old, new, m, H, W = make(2, 64, 5, 32, 16, 16, dtype, True, near=1e-3, seed=5)
    ops.distill_set_mode(2)
    s0, d0 = ops.teacher_distill(old, old, m, (H, W), 1.0, True)
    assert float(s0) == 0.0 and float(d0.float().abs().max()) == 0.0     # start of every task: exact zero
    s, d = ops.teacher_distill(old, new, m, (H, W), 1.0, True)
    ops.distill_set_mode(0)
    want, wgrad = fp64_reference(old, new, m, H, W)
    assert abs(float(s) - want) <= 1e-5 * want
    if dtype == F32:   # scale invariant: the relative accuracy does not degrade when old ~ new
        assert float((d.double().cpu() - wgrad).abs().max()) <= 3e-5 * float(wgrad.abs().max())


def test_tensor_core_distill_is_partition_independent(ops):
    """per-image calls (different CTA ranges, boundary rows completed by the finish kernel) reproduce the batched
    gradient bit for bit, two runs are identical, and a power-of-two scaling is exact (degree-2 homogeneity)"""
    old, new, m, H, W = make(5, 256, 32, 32, 16, 16, BF16, True, seed=9)
    ops.distill_set_mode(2)
    s, d = ops.teacher_distill(old, new, m, (H, W), 1.0, True)
    s_again, d_again = ops.teacher_distill(old, new, m, (H, W), 1.0, True)
    assert float(s) == float(s_again) and torch.equal(d, d_again)
    parts = [ops.teacher_distill(old[b:b + 1], new[b:b + 1], m[b:b + 1], (H, W), 1.0, True) for b in range(5)]
    assert torch.equal(torch.cat([p[1] for p in parts]), d)
    assert abs(sum(float(p[0]) for p in parts) - float(s)) <= 1e-6 * float(s)
    s2, d2 = ops.teacher_distill(old * 2, new * 2, m, (H, W), 1.0, True)
    ops.distill_set_mode(0)
    assert float(s2) == 4.0 * float(s) and torch.equal(d2.float(), 2.0 * d.float())


def test_distill_mode_switch_and_fallback(ops):
    from bacs_b200 import _cabi
    old, new, m, H, W = make(1, 16, 3, 6, 16, 16, F32, True)           # 24-byte rows: not a tensor-core shape
    assert ops.distill_kernel_variant(new, (H, W)) == 0
    s, d = ops.teacher_distill(old, new, m, (H, W), 1.0, True)            # auto: served by the FMA kernel
    ops.distill_set_mode(2)
    with pytest.raises(_cabi.BacsError):
        ops.teacher_distill(old, new, m, (H, W), 1.0, True)
    ops.distill_set_mode(0)
    want, _ = fp64_reference(old, new, m, H, W)
    assert abs(float(s) - want) <= 1e-5 * want
