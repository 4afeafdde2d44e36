"""Copy the files of the reference that the benchmarked path needs into ``baseline/_ref/VisFly`` (git-ignored).

The reference is pure Python, so "installing" it is copying: the package root ``__init__.py``, ``envs/**.py``, the
top-level ``utils/*.py`` (maths, type, randomization, common ...), ``utils/algorithms/*.py`` (the reference's own
BPTT / SHAC trainers, run on the new env by tests/test_gpu_reference_callers.py) and ``configs/drone/*.json``.  The copy is verbatim —
every fix the reference needs to run is a runtime monkeypatch in ``baseline/ref_loader.py`` (see REF_PATCHES.md).
Run by ``__graft_entry__.build()`` where ``/root/reference`` exists; the GPU box receives the result with the snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref", "VisFly")
PATTERNS = ["__init__.py", "LICENSE", "envs/*.py", "envs/base/*.py", "utils/*.py", "utils/algorithms/*.py",
            "configs/drone/*.json"]


def install(src: str = "/root/reference") -> int:
    """Returns the number of files copied (0 if the source tree is absent)."""
    if not os.path.isfile(os.path.join(src, "envs", "base", "dynamics.py")):
        return 0
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    count = 0
    for pat in PATTERNS:
        for path in glob.glob(os.path.join(src, pat)):
            rel = os.path.relpath(path, src)
            out = os.path.join(DEST, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(path, out)
            count += 1
    # packages the copy needs to be importable as `VisFly.*`
    for pkg in ("", "envs", "envs/base", "utils", "utils/algorithms", "configs"):
        init = os.path.join(DEST, pkg, "__init__.py")
        if os.path.isdir(os.path.dirname(init)) and not os.path.isfile(init):
            open(init, "w").close()
    return count


if __name__ == "__main__":
    print(f"copied {install()} reference files into {DEST}")
