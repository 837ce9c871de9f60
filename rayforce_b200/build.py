"""In-tree build of librfb200.so (nvcc, sm_100a only) and librfb200_ops.so (gcc).  No JIT cache: the outputs sit
next to this file so they travel to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(jobs: int = 8, verbose: bool = False) -> str:
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), "-j%d" % jobs, "all"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return os.path.join(HERE, "librfb200.so")


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
