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
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-Xptxas", "-v",
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


HOST_SRC = [os.path.join(ROOT, "host", "cpp", f) for f in ("oavif_host.cpp", "host_capi.cpp")]
HOST_OUT = os.path.join(ROOT, "lib", "liboavif_host.so")
HOST_CLI = os.path.join(ROOT, "lib", "oavif-b200")


def build_host(force: bool = False) -> str:
    """The C++ harness above the C ABI (g++ only; links the CUDA library by rpath)."""
    build()
    deps = [os.path.join(ROOT, "host", "cpp", f) for f in os.listdir(os.path.join(ROOT, "host", "cpp"))]
    deps += [OUT, os.path.join(ROOT, "..", "include", "oavif_host.h")]
    cli_src = os.path.join(ROOT, "host", "cpp", "main.cpp")
    stale = force or not os.path.exists(HOST_OUT) or any(os.path.getmtime(HOST_OUT) < os.path.getmtime(d) for d in deps)
    if stale:
        cmd = [os.environ.get("CXX", "g++"), "-O2", "-ffp-contract=off", "-std=c++17", "-Wall", "-Wextra", "-fPIC", "-shared", "-o", HOST_OUT,
               *HOST_SRC, "-L" + os.path.dirname(OUT), "-loavif_ssimu2", "-Wl,-rpath,$ORIGIN", "-ldl", "-lpthread"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("g++ failed: " + " ".join(cmd))
    if os.path.exists(cli_src) and (stale or not os.path.exists(HOST_CLI)):
        cmd = [os.environ.get("CXX", "g++"), "-O2", "-ffp-contract=off", "-std=c++17", "-Wall", "-Wextra", "-o", HOST_CLI, cli_src,
               "-L" + os.path.dirname(OUT), "-loavif_host", "-loavif_ssimu2", "-Wl,-rpath,$ORIGIN", "-ldl", "-lpthread"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("g++ failed: " + " ".join(cmd))
    return HOST_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_host(force="--force" in sys.argv))
