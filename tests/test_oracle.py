"""CPU tests of the oracle itself: analytic known answers + committed fixtures.

The reference has no tests or golden vectors for this path, so these pin the
restatement against closed-form facts of the published algorithm and against its own committed
outputs (regression).  Parity versus fssimu2 0.1.1 stays unpinned."""
import json
import os

import numpy as np
import pytest

from oavif_b200.host import synth


def test_weights_and_final_map(oracle):
    w = np.ctypeslib.as_array(oracle.lib().oracle_weights(), shape=(108,))
    assert w.shape == (108,) and int((w != 0).sum()) == 52
    assert abs(w.sum() - 888.3148365876135) < 1e-9  # checksum of the published list
    import ctypes as C
    z6 = ((C.c_double * 6) * 6)()
    z12 = ((C.c_double * 12) * 6)()
    assert oracle.lib().oracle_final_score(6, z6, z12) == 100.0
    z6[0][0] = 1.0  # weight index 0 is zero -> still 100
    assert oracle.lib().oracle_final_score(6, z6, z12) == 100.0


def test_srgb_lut(oracle):
    lut = oracle.srgb_lut()
    assert lut[0] == 0.0 and lut[255] == 1.0
    assert np.all(np.diff(lut) > 0)
    assert abs(lut[10] - (10 / 255) / 12.92) < 1e-9           # linear segment
    assert abs(lut[128] - ((128 / 255 + 0.055) / 1.055) ** 2.4) < 1e-7


def test_recursive_gaussian_coefficients(oracle):
    n2, d1, radius = oracle.rg_coeffs(1.5)
    assert radius == 5
    # the published constants are the binary32 roundings of the double solution
    np.testing.assert_array_equal(n2.astype(np.float32),
                                  np.float32([0.05529523640871048, -0.0588366873562336, 0.012955819256603718]))
    np.testing.assert_array_equal(d1[:2].astype(np.float32), np.float32([-1.9021130800247192, -1.1755704879760742]))
    np.testing.assert_allclose(d1[:2], [-2 * np.cos(np.pi / 10), -2 * np.cos(3 * np.pi / 10)], rtol=1e-14)
    assert abs(d1[2]) < 1e-15


def test_fir_taps_are_the_impulse_response(oracle):
    taps = oracle.fir_taps(1.5)
    want = [0.00941436780965, 0.0360111146584, 0.109335372777, 0.21292859201, 0.264621105488,
            0.21292859201, 0.109335372777, 0.0360111146584, 0.00941436780965]
    np.testing.assert_allclose(taps, want, rtol=1e-9)
    assert abs(taps.sum() - 1.0) < 1e-12
    # the f32 recursion on an impulse gives those taps (and nothing else) to round-off
    img = np.zeros((33, 33), np.float32)
    img[16, 16] = 1.0
    out = oracle.blur(img, oracle.BLUR_IIR).astype(np.float64)
    np.testing.assert_allclose(out[12:21, 12:21], np.outer(taps, taps), atol=2e-7)
    mask = np.ones_like(out, bool)
    mask[12:21, 12:21] = False
    assert np.abs(out[mask]).max() < 2e-7


def test_blur_of_ones_edge_profile(oracle):
    ones = np.ones((40, 40), np.float32)
    for mode in (oracle.BLUR_IIR, oracle.BLUR_FIR, oracle.BLUR_FIR64):
        out = oracle.blur(ones, mode)
        prof = out[20, :6].astype(np.float64)
        np.testing.assert_allclose(prof, [0.632310553, 0.845239145, 0.954574518, 0.990585632, 1.0, 1.0], atol=2e-6)
        np.testing.assert_allclose(out, out.T, atol=2e-6)


def test_blur_modes_agree_to_roundoff(oracle):
    rng = np.random.default_rng(0)
    p = rng.random((97, 131), dtype=np.float32)
    ref = oracle.blur(p, oracle.BLUR_FIR64)
    assert np.abs(oracle.blur(p, oracle.BLUR_FIR) - ref).max() < 5e-7
    assert np.abs(oracle.blur(p, oracle.BLUR_IIR) - ref).max() < 2e-5


def test_downsample_clamps_odd_edges(oracle):
    import ctypes as C
    a = np.arange(15, dtype=np.float32).reshape(3, 5)
    out = np.zeros((2, 3), np.float32)
    oracle.lib().oracle_downsample2x(a.ctypes.data_as(C.POINTER(C.c_float)), 5, 3,
                                     out.ctypes.data_as(C.POINTER(C.c_float)))
    want = np.array([[(0 + 1 + 5 + 6) / 4, (2 + 3 + 7 + 8) / 4, (4 + 4 + 9 + 9) / 4],
                     [(10 + 11 + 10 + 11) / 4, (12 + 13 + 12 + 13) / 4, (14 + 14 + 14 + 14) / 4]], np.float32)
    np.testing.assert_array_equal(out, want)


def test_cbrt_fixed_sequence_accuracy(oracle):
    import ctypes as C
    f = oracle.lib().oracle_cbrtf
    f.restype, f.argtypes = C.c_float, [C.c_float]
    xs = np.concatenate([np.linspace(0.0037, 1.2, 20001), [0.0037930732552754493, 1.0, 0.125]]).astype(np.float32)
    got = np.array([f(float(x)) for x in xs], np.float32)
    want = np.cbrt(xs.astype(np.float64))
    ulp = np.abs(got - want) / np.spacing(want.astype(np.float32))
    assert ulp.max() < 1.0
    assert f(0.125) == 0.5 and f(1.0) == 1.0


def test_xyb_of_white_black_and_range(oracle):
    img = np.zeros((8, 8, 3), np.uint8)
    img[:, 4:] = 255
    x = oracle.xyb_at_scale(img, 0)
    # gray pixels: L == M so X' == 0.42 exactly; B - Y offset == 0.55 for neutral colours to round-off
    assert np.all(x[0] == np.float32(0.42))
    assert abs(x[1][0, 0] - 0.01) < 1e-7                      # black: Y = 0 + 0.01
    assert abs(x[1][0, 7] - (np.cbrt(1.0037930732552754) - 0.15595420054924863 + 0.01)) < 1e-6
    np.testing.assert_allclose(x[2], 0.55, atol=1e-6)


@pytest.mark.parametrize("size,want", [((7, 64), 0), ((8, 8), 2), ((15, 40), 3), ((100, 75), 5), ((64, 64), 5),
                                        ((255, 255), 6), ((640, 360), 6)])
def test_scale_schedule(oracle, size, want):
    w, h = size
    img = synth.synth(w, h, "noise", 3)
    _, det = oracle.ssimu2_rgb8(img, synth.distort(img, 0.3), detail=True)
    assert det.n_scales == want


def test_identical_is_exactly_100_and_monotone(oracle):
    src = synth.synth(160, 120, "mixture", 11)
    for mode in (oracle.BLUR_IIR, oracle.BLUR_FIR):
        assert oracle.ssimu2_rgb8(src, src, mode) == 100.0
        scores = [oracle.ssimu2_rgb8(src, synth.distort(src, s), mode) for s in (0.1, 0.3, 0.6, 1.0)]
        assert all(a > b for a, b in zip(scores, scores[1:])), scores
        assert scores[0] < 100.0


def test_strided_input_equals_tight(oracle):
    import ctypes as C
    src = synth.synth(50, 40, "mixture", 2)
    dist = synth.distort(src, 0.4)
    tight = oracle.ssimu2_rgb8(src, dist)
    pad = np.zeros((40, 50 * 3 + 13), np.uint8)
    pad[:, :150] = src.reshape(40, 150)
    sc = C.c_double()
    rc = oracle.lib().oracle_ssimu2_rgb8(pad.ctypes.data_as(C.POINTER(C.c_uint8)), pad.strides[0],
                                         dist.ctypes.data_as(C.POINTER(C.c_uint8)), 150, 50, 40, 0,
                                         C.byref(sc), None)
    assert rc == 0 and sc.value == tight


def test_bad_arguments(oracle):
    import ctypes as C
    sc = C.c_double()
    buf = np.zeros(48, np.uint8).ctypes.data_as(C.POINTER(C.c_uint8))
    assert oracle.lib().oracle_ssimu2_rgb8(None, 12, buf, 12, 4, 4, 0, C.byref(sc), None) < 0
    assert oracle.lib().oracle_ssimu2_rgb8(buf, 11, buf, 12, 4, 4, 0, C.byref(sc), None) < 0
    assert oracle.lib().oracle_ssimu2_rgb8(buf, 12, buf, 12, 0, 4, 0, C.byref(sc), None) < 0


def test_committed_scores_regression(oracle, golden_dir):
    with open(os.path.join(golden_dir, "oracle_scores.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 7
    for c in cases:
        src = synth.synth(c["w"], c["h"], c["kind"], c["seed"])
        dist = synth.distort(src, c["strength"], seed=c["seed"] + 100)
        for name, mode in (("iir", oracle.BLUR_IIR), ("fir", oracle.BLUR_FIR)):
            sc, det = oracle.ssimu2_rgb8(src, dist, mode, detail=True)
            assert det.n_scales == c[name]["n_scales"]
            assert sc == pytest.approx(c[name]["score"], abs=1e-9), (c["w"], c["h"], name)
            np.testing.assert_allclose(oracle.detail_sums(det), np.array(c[name]["sums"]), rtol=1e-12, atol=0)


def test_fast_build_matches_reproducible_build(oracle):
    src = synth.synth(200, 150, "mixture", 5)
    dist = synth.distort(src, 0.35)
    for mode in (oracle.BLUR_IIR, oracle.BLUR_FIR):
        assert oracle.ssimu2_rgb8(src, dist, mode, fast=True) == oracle.ssimu2_rgb8(src, dist, mode, fast=False)


def test_to_rgb8_semantics(oracle):
    rng = np.random.default_rng(4)
    g16 = rng.integers(0, 65536, (6, 5, 2)).astype(np.uint16)        # gray + alpha, 16-bit
    out = oracle.to_rgb8(g16, 2, True)
    assert np.all(out[..., 0] == (g16[..., 0] >> 8)) and np.all(out[..., 1] == out[..., 0]) and np.all(out[..., 2] == out[..., 0])
    rgba = rng.integers(0, 256, (6, 5, 4)).astype(np.uint8)
    np.testing.assert_array_equal(oracle.to_rgb8(rgba, 4, False), rgba[..., :3])
    rgb16 = rng.integers(0, 65536, (6, 5, 3)).astype(np.uint16)
    np.testing.assert_array_equal(oracle.to_rgb8(rgb16, 3, True), (rgb16 >> 8).astype(np.uint8))


def test_source_samples_are_the_reference_formulas(oracle):
    """encodeAvifToBuffer's loops (io.zig:566-609): exhaustive over the sample values."""
    v8 = np.arange(256, dtype=np.uint8).reshape(16, 16, 1)
    want = [(int(v) * 1023 + 127) // 255 for v in range(256)]
    got = oracle.source_samples(v8, 10)
    assert got.dtype == np.uint16 and got.reshape(-1).tolist() == want
    assert got.min() == 0 and got.max() == 1023
    v16 = np.arange(65536, dtype=np.uint16).reshape(256, 64, 4)
    np.testing.assert_array_equal(oracle.source_samples(v16, 10), v16 >> 6)
    out8 = oracle.source_samples(v16, 8)
    assert out8.dtype == np.uint8
    np.testing.assert_array_equal(out8, (v16 >> 8).astype(np.uint8))
    with pytest.raises(RuntimeError):
        oracle.source_samples(v8, 8)          # passed through by the reference (io.zig:611-613)


# ---- fssimu2 golden vectors (scripts/pin_fssimu2.md) ---------------------------------------------------------------
FSSIMU2_VECTORS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fssimu2_scores.json")


def fssimu2_vectors():
    """Records produced by scripts/pin_fssimu2 on a machine with zig + the 0.1.1 tarball; absent here: the
    reference's scorer can be neither fetched nor built in this image, so parity against it stays UNPINNED."""
    if not os.path.exists(FSSIMU2_VECTORS):
        pytest.skip("parity unpinned: tests/golden/fssimu2_scores.json absent (see scripts/pin_fssimu2.md)")
    with open(FSSIMU2_VECTORS) as f:
        return json.load(f)


def test_which_variant_identifies_the_generating_reading(oracle):
    """scripts/pin_fssimu2/which_variant.py: vectors made under a non-default combination of the variant switches
    rank exactly that combination first, at zero distance — the tool a maintainer with zig runs when pinned vectors
    disagree with the default reading."""
    import importlib.util
    from oavif_b200.host import synth
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("which_variant", os.path.join(here, "..", "scripts", "pin_fssimu2", "which_variant.py"))
    W = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(W)
    recs = []
    try:
        for i, (w, h) in enumerate(((96, 80), (150, 130))):
            src = synth.synth(w, h, i % 4, i)
            dst = synth.distort(src, 0.4 + 0.3 * i, seed=i + 100)
            oracle.set_variant(oracle.VARIANT_VERTICAL_ORDER | oracle.VARIANT_F32_TRANSFER, fast=True)
            recs.append({"w": w, "h": h, "kind": i % 4, "seed": i, "strength": 0.4 + 0.3 * i,
                         "score": oracle.ssimu2_rgb8(src, dst, oracle.BLUR_IIR, fast=True)})
    finally:
        oracle.set_variant(0, fast=True)
    ranked = W.rank(recs)
    assert ranked[0][0] == 0.0 and ranked[0][2] == "recursive + vertical_order + f32_transfer", ranked[:3]
    assert ranked[1][0] > 0.0
    assert oracle.lib(True).oracle_get_variant() == 0 if hasattr(oracle.lib(True), "oracle_get_variant") else True


def test_oracle_against_fssimu2_vectors(oracle):
    from oavif_b200.host import synth
    for c in fssimu2_vectors():
        src = synth.synth(c["w"], c["h"], c["kind"], c["seed"])
        dst = synth.distort(src, c["strength"], seed=c["seed"] + 100)
        assert abs(oracle.ssimu2_rgb8(src, dst, oracle.BLUR_IIR, fast=True) - c["score"]) <= 0.05, c   # north_star's bar
