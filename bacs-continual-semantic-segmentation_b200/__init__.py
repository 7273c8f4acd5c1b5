"""bacs_b200 -- B200-native (sm_100a) implementation of BACS's per-pixel continual
learning loss path, behind the reference's loss-class interface.

Layout (mirrors the reference's module names for the path only):
  csrc/                 hand-written CUDA kernels + the C-ABI (include/bacs_b200.h)
  _cabi.py              ctypes binding of libbacs_b200.so (no torch types in the ABI)
  ops.py                tensor-level wrappers: validation, workspaces, stream plumbing
  autograd.py           torch.autograd.Function wrappers (fused fwd+bwd kernels)
  loss/                 BaseLoss, Prototypes, ExperienceReplay, BACSLoss, loss_utils
  networks/             BgDetector / classification_head mirror
  training/             IoU metric (confusion matrix), replay Buffer
  labels.py             continual-learning label remap
  distributed.py        packed single all-reduce of prototype sums/counts + confmat
  synth.py              synthetic VOC/ADE/Cityscapes-shaped inputs

There is no CPU fallback: every op raises if the CUDA library is missing."""
__version__ = "0.1.0"
