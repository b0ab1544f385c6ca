"""The decoded-YUV -> RGB8 oracle is REFERENCE-PINNED: bit-exact against libavif 1.4.1's
avifImageYUVToRGB (the call at /root/reference/src/io.zig:478), live when Pillow's bundled
libavif is present and always against the committed fixtures generated from it."""
import os

import numpy as np
import pytest

import avif_ctypes as A


def test_against_committed_libavif_vectors(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "yuv2rgb_libavif.npz"))
    n = 0
    for depth in (8, 10):
        y, u, v = g[f"y{depth}"], g[f"u{depth}"], g[f"v{depth}"]
        for m in (1, 2, 5, 6, 9):
            for al in (0, 1):
                got = oracle.yuv444_to_rgb8(y, u, v, depth, m, bool(al))
                np.testing.assert_array_equal(got, g[f"rgb_d{depth}_m{m}_a{al}"], err_msg=f"d{depth} m{m} a{al}")
                n += 1
    assert n == 20
    # the two 10-bit conversions really are different functions
    assert (g["rgb_d10_m2_a0"] != g["rgb_d10_m2_a1"]).any()
    assert (g["rgb_d8_m2_a0"] == g["rgb_d8_m2_a1"]).all()


def test_against_committed_decoder_output(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "avif_roundtrip.npz"))
    for tag in ("65", "40"):
        got = oracle.yuv444_to_rgb8(g[f"y{tag}"], g[f"u{tag}"], g[f"v{tag}"], 8, int(g[f"matrix{tag}"]))
        np.testing.assert_array_equal(got, g[f"rgb{tag}"])
    np.testing.assert_array_equal(oracle.yuv444_to_rgb8(g["y10"], g["u10"], g["v10"], 10, 2, False), g["rgb10"])
    np.testing.assert_array_equal(oracle.yuv444_to_rgb8(g["y10"], g["u10"], g["v10"], 10, 2, True), g["rgb10a"])


@pytest.mark.skipif(A.find_libavif() is None, reason="Pillow's bundled libavif not present")
def test_live_against_libavif(oracle):
    rng = np.random.default_rng(99)
    for depth in (8, 10):
        dt = np.uint8 if depth == 8 else np.uint16
        y, u, v = (rng.integers(0, 1 << depth, (37, 61)).astype(dt) for _ in range(3))
        for m in (1, 2, 9):
            for al in (False, True):
                np.testing.assert_array_equal(oracle.yuv444_to_rgb8(y, u, v, depth, m, al),
                                              A.yuv444_to_rgb8(y, u, v, depth, m, al))


def test_unsupported_matrix_and_depth(oracle):
    y = np.zeros((4, 4), np.uint8)
    with pytest.raises(RuntimeError):
        oracle.yuv444_to_rgb8(y, y, y, 8, matrix=0)
    y16 = np.zeros((4, 4), np.uint16)
    import ctypes as C
    out = np.zeros((4, 4, 3), np.uint8)
    rc = oracle.lib().oracle_yuv444_to_rgb8(y16.ctypes.data, y16.ctypes.data, y16.ctypes.data, 8, 8, 8, 4, 4, 12, 2, 0,
                                            out.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert rc < 0
