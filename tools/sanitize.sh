#!/bin/bash
# compute-sanitizer targets for the kernels that synchronise through mbarriers / TMA rings / tensor memory
# (race detection).  One tool per gpurun call (B200_PROFILING.md), on small shapes:
#   gpurun -- 'bash tools/sanitize.sh racecheck'      (or memcheck / synccheck)
# The pixel kernel (pixel_wce: 4-stage TMA ring, rotating tile ownership), the tensor-core distill kernel
# (distill_tc: operand ring, tcgen05 commit barriers, bulk-copied mask rows), the register-column pixel kernel
# (pixel_regs: one TMA landing buffer per CTA) and the label / prototype chain.
set -e
TOOL=${1:-racecheck}
cd "$(dirname "$0")/.."
compute-sanitizer --tool "$TOOL" --print-limit 20 python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from bacs_b200 import ops, synth, _cabi
gen = torch.Generator().manual_seed(0)
# tensor-core distill, two row blocks per interval, boundary rows between CTA ranges
old = torch.randn(2, 130, 5, 16, generator=gen).bfloat16().cuda()
new = torch.randn(2, 130, 5, 16, generator=gen).bfloat16().cuda()
m = (torch.rand(2, 120, 256, generator=gen) > 0.4).to(torch.uint8).cuda()
assert ops.distill_kernel_variant(new, (120, 256)) == 1
s, d = ops.teacher_distill(old, new, m, (120, 256), 1.0, True)
# the training-step pixel kernel on 512-pixel row tiles, with the seen heads
cfg = synth.CONFIGS["row512"]
inp = synth.make_step_inputs(cfg, seed=1, dtype=torch.bfloat16)
loss_fn, net, batch, leaves = synth.build_bacs_step(cfg, inp)
loss, preds = loss_fn.compute_loss(batch, net, train=True)
loss.backward()
# the one-pass register-column kernel (K = 151, bf16): TMA landing buffer re-armed per tile, lane-pair exchange
cfg = synth.CONFIGS["row512_k151"]
inp = synth.make_step_inputs(cfg, seed=5, dtype=torch.bfloat16)
lab = synth.make_labels(cfg, torch.Generator().manual_seed(3), classes=list(range(1, cfg.K)))
out = ops.pixel_loss(inp.logits.cuda(), lab.cuda(), _cabi.PIX_WEIGHTED_CE, want_grad=True,
                     z=torch.randn(cfg.B, cfg.T, cfg.h, cfg.w).cuda(), want_distill_mask=True, old_cl=cfg.old_cl,
                     focal_head=cfg.T - 1)
assert out["variant"] == 5
torch.cuda.synchronize()
print("sanitize target ran: distill %.4f, step loss %.5f" % (float(s), float(loss)))
PY
