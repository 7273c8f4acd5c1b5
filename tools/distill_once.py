"""Runs the teacher-distill kernels a few times at the headline shape (ncu target)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from bacs_b200 import synth, ops
cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "voc15-1_b24"]
inp = synth.make_step_inputs(cfg, seed=0, dtype=torch.bfloat16, device="cuda")
g = torch.Generator(device="cuda").manual_seed(1)
mask = ((inp.mask == 0) & (torch.rand(cfg.B, cfg.H, cfg.W, device="cuda", generator=g) > 0.3)).to(torch.uint8)
for _ in range(4):
    ops.teacher_distill(inp.old_att, inp.new_att, mask, (cfg.H, cfg.W), 1e-3, True, want_scaled=True)
torch.cuda.synchronize()
