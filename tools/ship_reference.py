#!/usr/bin/env python
"""Copy the UNMODIFIED reference (its Python sources only) into the git-ignored baseline/_ref/ so that it travels to
the GPU box with the gpurun snapshot (/root/reference itself does not exist there).

    python tools/ship_reference.py

Used by tests/test_gpu_zz_dropin.py (the byol_b200 classes dropped into the reference's own main.execute_graph) and
tools/ref_gpu_baseline.py (the reference's stock PyTorch path timed on the same GPUs).  Nothing is modified; the
missing `helpers` / `datasets` / `tree` submodules come from oracle/ref_shims at import time.  baseline/_ref/ is
never committed (.gitignore) — reference sources stay out of this repository's history.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["main.py", "objective.py", "optimizers/__init__.py", "optimizers/lars.py", "optimizers/scheduler.py"]


def main():
    if not os.path.isdir(SRC):
        print("no %s here: nothing to ship" % SRC)
        return 1
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        print("shipped", rel)
    return 0


if __name__ == "__main__":
    sys.exit(main())
