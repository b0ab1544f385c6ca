"""The Zig side is shipped as source (no zig toolchain here): what can be checked without one."""
import glob
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ZIG = os.path.join(HERE, "..", "oavif_b200", "zig")
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")) or shutil.which("patch") is None,
                    reason="reference tree (or patch) not on this machine")
def test_patches_apply_to_the_reference_tree(tmp_path):
    for name in ("src", "build.zig", "build.zig.zon"):
        src = os.path.join(REF, name)
        (shutil.copytree if os.path.isdir(src) else shutil.copy)(src, tmp_path / name)
    patches = sorted(glob.glob(os.path.join(ZIG, "patches", "*.patch")))
    assert len(patches) == 6
    for p in patches:
        r = subprocess.run(["patch", "-p1", "--no-backup-if-mismatch", "-i", p], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0 and "FAILED" not in r.stdout and "fuzz" not in r.stdout, (p, r.stdout, r.stderr)
    patched = (tmp_path / "src" / "tq.zig").read_text()
    assert "scoreYuv444" in patched and "computeSsimu2(" not in patched.split("fn computeScoreAtQuality")[1].split("\n}\n")[0]
    assert "sourceSamples" in (tmp_path / "src" / "main.zig").read_text()


def test_shim_binds_every_entry_point_it_names():
    """Every c.oavif_ssimu2_* the shim calls is declared in the header it @cImports."""
    import re
    shim = open(os.path.join(ZIG, "fssimu2.zig")).read()
    header = open(os.path.join(HERE, "..", "include", "oavif_ssimu2.h")).read()
    used = set(re.findall(r"c\.(oavif_ssimu2_[a-z0-9_]+)\(", shim))
    assert used and all(re.search(r"\b%s\(" % u, header) for u in used), used
