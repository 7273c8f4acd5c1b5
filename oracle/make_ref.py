"""Stages the reference's own hot-path modules under oracle/_ref/ (TEST INFRASTRUCTURE, build container only).

    python oracle/make_ref.py            # called by __graft_entry__.build() when /root/reference exists

The reference is pure Python (no setup.py / pyproject.toml): "building" it means placing the modules of the
BACS loss path where the GPU box can import them -- /root/reference does not exist there.  oracle/_ref/ is listed in
.gitignore (the files never enter the history) but not in .gpurunignore (they travel like the built .so).  The files
are byte-for-byte copies; nothing is edited, and the product package never imports them.  tests/golden/ref_shim.py
(BACS_REFERENCE_ROOT) imports them without the package __init__ files, oracle/ref_step.py drives one training step."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
MODULES = ("loss/base_loss.py", "loss/bacs_loss.py", "loss/prototypes.py", "loss/experience_replay.py",
           "training/loss_utils.py", "training/buffer.py", "training/metrics.py", "training/utils.py",
           "networks/bg_detector.py")


def stage(reference_root: str = "/root/reference") -> bool:
    if not os.path.isdir(os.path.join(reference_root, "loss")):
        return False
    for rel in MODULES:
        src, dst = os.path.join(reference_root, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    with open(os.path.join(DEST, "STAGED_FROM"), "w") as f:
        f.write(reference_root + "\n")
    return True


def staged() -> bool:
    return all(os.path.exists(os.path.join(DEST, rel)) for rel in MODULES)


if __name__ == "__main__":
    ok = stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("staged %d reference modules under %s" % (len(MODULES), DEST) if ok else "reference tree not found")
