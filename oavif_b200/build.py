"""Build the CUDA library in-tree: ``python -m oavif_b200.build``.

One nvcc invocation, sm_100a only (no multi-arch, no JIT cache): the resulting
``oavif_b200/lib/liboavif_ssimu2.so`` is git-ignored but travels with the repo snapshot.
``-fmad=false`` is part of the numerical contract (see csrc/ssimu2_common.cuh).
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(ROOT, "csrc", "ssimu2_api.cu")
OUT = os.path.join(ROOT, "lib", "liboavif_ssimu2.so")
DEPS = [os.path.join(ROOT, "csrc", f) for f in os.listdir(os.path.join(ROOT, "csrc"))] + [
    os.path.join(ROOT, "..", "include", "oavif_ssimu2.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    with open(os.path.join(os.path.dirname(OUT), "ptxas.log"), "w") as f:
        f.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
