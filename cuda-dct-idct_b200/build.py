"""In-tree build of the CUDA libraries (nvcc, sm_100a only)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(jobs: int | None = None, verbose: bool = False) -> None:
    """make -C csrc: libb200dct.so + libb200dct_compat.so next to this file."""
    jobs = jobs or min(8, os.cpu_count() or 1)
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), f"-j{jobs}"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
