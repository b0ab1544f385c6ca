"""The C-ABI library loads on a CPU-only box and exports every symbol include/oavif_ssimu2.h declares
(no compute calls here: without a GPU they must fail loudly, never fall back)."""
import os
import re

import numpy as np
import pytest

from oavif_b200.host import ssimu2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "oavif_ssimu2.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oavif_ssimu2_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_and_exports_the_header():
    from oavif_b200 import build
    build.build()
    L = ssimu2.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert sorted(ssimu2.SYMBOLS) == names, "python binding and header disagree"
    assert L.oavif_ssimu2_abi_version() == 1


def test_built_for_sm_100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", ssimu2.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "oavif_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".zig")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and '#include "../../oracle' not in text, f


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ssimu2.Ssimu2Error) as e:
        ssimu2.Scorer(64, 64)
    assert e.value.code == ssimu2.E_CUDA
    with pytest.raises(ssimu2.Ssimu2Error):
        ssimu2.compute_ssimu2(np.zeros((16, 16, 3), np.uint8), np.zeros((16, 16, 3), np.uint8))


def test_argument_errors_without_a_context():
    L = ssimu2.load()
    import ctypes as C
    out = C.c_double()
    z = np.zeros(48, np.uint8)
    assert L.oavif_ssimu2_compute_rgb8(None, z.ctypes.data, 4, 4, 3, C.byref(out)) == ssimu2.E_ARG
    assert L.oavif_ssimu2_compute_rgb8(z.ctypes.data, z.ctypes.data, 4, 4, 4, C.byref(out)) == ssimu2.E_UNSUPPORTED
    assert b"channels" in L.oavif_ssimu2_last_error(None)
    assert L.oavif_ssimu2_compute_rgb8(z.ctypes.data, z.ctypes.data, 0, 4, 3, C.byref(out)) == ssimu2.E_ARG
    ctx = C.c_void_p()
    assert L.oavif_ssimu2_ctx_create(0, 0, 10, 1, C.byref(ctx)) == ssimu2.E_ARG
    assert L.oavif_ssimu2_score_rgb8(None, z.ctypes.data, 12, C.byref(out)) == ssimu2.E_ARG
    L.oavif_ssimu2_ctx_destroy(None)  # no-op
