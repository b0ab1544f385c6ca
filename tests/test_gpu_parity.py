"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Run with -m gpu on a B200.

Bars (stated here, enforced below):
  * integer / byte work (decoded YUV -> RGB8): bit-exact against libavif-generated fixtures;
  * XYB pyramid planes and blurred planes: bit-identical binary32 against the oracle (both sides
    execute one fixed sequence of IEEE operations);
  * pooled sums: relative 1e-5 (the maps run in binary32 on the GPU where the published code widens
    a few operations to double; sums are reduced in a different but fixed order);
  * score: |delta| <= 1e-4 against the oracle in the same blur mode — 500x inside north_star's 0.05.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oavif_b200.host import ssimu2, synth

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-4
SUM_RTOL = 1e-5
MODES = [(ssimu2.BLUR_RECURSIVE, 0, "recursive"), (ssimu2.BLUR_FIR, 1, "fir")]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def scorer():
    with ssimu2.Scorer(1100, 800, 4) as sc:
        yield sc


def test_cuda_extension_is_what_runs():
    L = ssimu2.load()
    assert L.oavif_ssimu2_abi_version() == 1
    with open("/proc/self/maps") as f:
        assert "liboavif_ssimu2.so" in f.read()


# ---- K0: decoded YUV -> RGB8 --------------------------------------------------------------------
def test_yuv_to_rgb8_bit_exact_vs_libavif_vectors(scorer, golden_dir):
    g = np.load(os.path.join(golden_dir, "yuv2rgb_libavif.npz"))
    for depth in (8, 10):
        y, u, v = g[f"y{depth}"], g[f"u{depth}"], g[f"v{depth}"]
        for m in (1, 2, 5, 6, 9):
            for al in (0, 1):
                got = scorer.yuv444_to_rgb8(y, u, v, depth, m, bool(al))
                np.testing.assert_array_equal(got, g[f"rgb_d{depth}_m{m}_a{al}"], err_msg=f"d{depth} m{m} a{al}")


def test_yuv_to_rgb8_exhaustive_luma_chroma_grid(scorer, oracle):
    # every 10-bit luma against a chroma lattice that includes the extremes, both conversions
    yy = np.arange(1024, dtype=np.uint16)
    cc = np.array([0, 1, 2, 3, 255, 256, 511, 512, 513, 767, 1020, 1021, 1022, 1023], np.uint16)
    Y = np.tile(yy, (len(cc) * len(cc), 1))
    U = np.repeat(np.repeat(cc, len(cc)), 1024).reshape(-1, 1024)
    V = np.repeat(np.tile(cc, len(cc)), 1024).reshape(-1, 1024)
    for m in (2, 1, 9):
        for al in (False, True):
            np.testing.assert_array_equal(scorer.yuv444_to_rgb8(Y, U, V, 10, m, al),
                                          oracle.yuv444_to_rgb8(Y, U, V, 10, m, al))
    Y8, U8, V8 = (Y >> 2).astype(np.uint8), (U >> 2).astype(np.uint8), (V >> 2).astype(np.uint8)
    np.testing.assert_array_equal(scorer.yuv444_to_rgb8(Y8, U8, V8, 8, 2), oracle.yuv444_to_rgb8(Y8, U8, V8, 8, 2))


# ---- K1-K3: pyramid planes ------------------------------------------------------------------------
@pytest.mark.parametrize("size", [(64, 64), (65, 63), (100, 75), (257, 129), (333, 257), (1027, 771)])
def test_xyb_pyramid_bit_identical(scorer, oracle, size):
    w, h = size
    src = synth.synth(w, h, "mixture", w)
    dist = synth.distort(src, 0.3, seed=h)
    scorer.set_blur(ssimu2.BLUR_FIR)
    scorer.set_source(src)
    scorer.score_rgb8(dist)
    n = scorer.detail().n_scales
    assert n >= 5
    for s in range(n):
        ws, wd = oracle.xyb_at_scale(src, s), oracle.xyb_at_scale(dist, s)
        for c in range(3):
            np.testing.assert_array_equal(bits(scorer.xyb(0, s, c)), bits(ws[c]), err_msg=f"src s{s} c{c}")
            np.testing.assert_array_equal(bits(scorer.xyb(1, s, c)), bits(wd[c]), err_msg=f"dist s{s} c{c}")


# ---- K4, product kernel: the rows pass the scored path leaves behind ----------------------------------
@pytest.fixture(params=[ssimu2.TILES_TMA, ssimu2.TILES_CP_ASYNC], ids=["tma", "cp_async"])
def tile_path(request, scorer):
    """Both tile-movement forms of the RECURSIVE kernels (include/oavif_ssimu2.h, OAVIF_SSIMU2_OPT_TILE_PATH)."""
    scorer.set_tile_path(request.param)
    yield request.param
    assert scorer.get_option(ssimu2.OPT_TILE_PATH) == request.param   # no silent switch to the other form
    scorer.set_tile_path(ssimu2.TILES_TMA)


@pytest.fixture(params=[ssimu2.TILES_TMA, ssimu2.TILES_CP_ASYNC, ssimu2.TILES_FUSED, ssimu2.TILES_TMA_DECOUPLED],
                ids=["tma", "cp_async", "fused", "tma_decoupled"])
def tile_path3(request, scorer):
    """... the fused kernel (no row-filtered planes exist there: everything but the rows tap applies) and the columns
    kernel without its per-batch block barrier."""
    scorer.set_tile_path(request.param)
    yield request.param
    assert scorer.get_option(ssimu2.OPT_TILE_PATH) == request.param
    scorer.set_tile_path(ssimu2.TILES_TMA)


@pytest.mark.parametrize("size", [(64, 64), (65, 63), (100, 75), (333, 257), (1027, 771)])
def test_rows_pass_of_the_scored_path_is_bit_identical(scorer, oracle, size, tile_path):
    """k_iir_rows_tma / k_iir_rows (packed-pair recursion, interleaved pair planes) against the oracle's horizontal
    pass, bit for bit, for all five quantities, both call forms (first call after set_source / cached source) and
    a batch."""
    w, h = size
    src = synth.synth(w, h, "mixture", w + 1)
    d0, d1 = synth.distort(src, 0.3, seed=h), synth.distort(src, 0.8, seed=h + 1)
    scorer.set_blur(ssimu2.BLUR_RECURSIVE)
    scorer.set_source(src)
    scorer.score_rgb8(d0)                      # first call: the source half rides along
    n = scorer.detail().n_scales

    def expect(s, c, cand_rgb):
        a, b = oracle.xyb_at_scale(src, s)[c], oracle.xyb_at_scale(cand_rgb, s)[c]
        return [oracle.blur(q, oracle.BLUR_IIR, rows_only=True) for q in (a, b, a * a, b * b, a * b)]

    def check(cand_index, cand_rgb, what):
        for s in range(n):
            for c in range(3):
                for q, want in enumerate(expect(s, c, cand_rgb)):
                    np.testing.assert_array_equal(bits(scorer.rows(cand_index, q, s, c)), bits(want),
                                                  err_msg=f"{what}: quantity {q} scale {s} channel {c}")

    check(0, d0, "first call")
    scorer.score_rgb8(d1)                      # cached source: candidate half only
    check(0, d1, "cached source")
    scorer.set_source(src)
    scorer.score_batch_rgb8([d0, d1])          # batch: candidate 0's CTAs refill the source cache
    check(0, d0, "batch[0]")
    check(1, d1, "batch[1]")


# ---- K4+K5, product kernel: what the columns pass hands to the error maps ----------------------------------
@pytest.mark.parametrize("size", [(64, 64), (65, 63), (100, 75), (333, 257), (700, 300)])
def test_cols_pass_of_the_scored_path_is_bit_identical(scorer, oracle, size, tile_path3):
    """k_iir_cols (the kernel on the scored path, not the plain debug filter): the five fully blurred values it
    feeds the SSIM / edge-diff maps — mu1, mu2, sigma11, sigma22, sigma12 — against the oracle's two-pass blur,
    bit for bit, every scale and channel, single call and second candidate of a batch."""
    w, h = size
    src = synth.synth(w, h, "mixture", w + 2)
    d0, d1 = synth.distort(src, 0.3, seed=h), synth.distort(src, 0.8, seed=h + 1)
    scorer.set_blur(ssimu2.BLUR_RECURSIVE)
    scorer.set_source(src)
    s_single = scorer.score_rgb8(d0)
    sums_single = scorer.sums(0).copy()
    n = scorer.detail().n_scales

    def check(cand_index, cand_rgb, what):
        for s in range(n):
            for c in range(3):
                a, b = oracle.xyb_at_scale(src, s)[c], oracle.xyb_at_scale(cand_rgb, s)[c]
                want = [oracle.blur(q, oracle.BLUR_IIR) for q in (a, b, a * a, b * b, a * b)]
                got = scorer.cols(cand_index, s, c)
                for q in range(5):
                    np.testing.assert_array_equal(bits(got[q]), bits(want[q]),
                                                  err_msg=f"{what}: quantity {q} scale {s} channel {c}")

    check(0, d0, "single")
    assert (scorer.sums(0) == sums_single).all()      # the tap's re-run rewrote the same pooled sums
    batch = scorer.score_batch_rgb8([d0, d1])
    assert batch[0] == s_single
    check(1, d1, "batch[1]")


def test_all_input_forms_build_the_same_pyramid(scorer, oracle):
    w, h = 203, 117
    src = synth.synth(w, h, "noise", 8)
    dist = synth.distort(src, 0.5)
    scorer.set_blur(ssimu2.BLUR_FIR)
    scorer.set_source(src)
    for depth in (8, 10):
        for rgba in (False, True):
            y, u, v = synth.rgb8_to_yuv444(dist, depth, 1)
            rgb = oracle.yuv444_to_rgb8(y, u, v, depth, 1, rgba)
            a = scorer.score_yuv444(y, u, v, depth, 1, rgba)
            planes = [scorer.xyb(1, 0, c) for c in range(3)]
            b = scorer.score_rgb8(rgb)
            assert a == b
            for c in range(3):
                np.testing.assert_array_equal(bits(planes[c]), bits(scorer.xyb(1, 0, c)))


def test_loader_native_layouts_follow_toRGB8(scorer, oracle):
    """Image.toRGB8 (io.zig:57-133) and the RGBA repack (io.zig:654-663) on the device: every channel
    count and bit depth must score exactly like the CPU-reduced RGB8."""
    rng = np.random.default_rng(12)
    w, h = 150, 100
    base = synth.synth(w, h, "mixture", 4)
    dist = synth.distort(base, 0.3)
    scorer.set_blur(ssimu2.BLUR_RECURSIVE)
    for ch in (1, 2, 3, 4):
        for dt in (np.uint8, np.uint16):
            px = np.empty((h, w, ch), dt)
            hi = base.astype(np.uint16) * 257 + rng.integers(0, 200, base.shape).astype(np.uint16) if dt == np.uint16 else base
            if ch >= 3:
                px[..., :3] = hi
            else:
                px[..., 0] = hi[..., 1]
            if ch in (2, 4):
                px[..., -1] = rng.integers(0, np.iinfo(dt).max, (h, w))
            want_rgb = oracle.to_rgb8(px, ch, dt == np.uint16)
            scorer.set_source_pixels(px)
            got = scorer.score_rgb8(dist)
            scorer.set_source(want_rgb)
            assert got == scorer.score_rgb8(dist), (ch, dt)
            np.testing.assert_array_equal(bits(scorer.xyb(0, 0, 1)), bits(oracle.xyb_at_scale(want_rgb, 0)[1]))
    scorer.set_source(base)
    rgba = np.dstack([dist, rng.integers(0, 255, (h, w)).astype(np.uint8)])
    assert scorer.score_pixels(rgba) == scorer.score_rgb8(dist)
    with pytest.raises(ssimu2.Ssimu2Error) as e:
        scorer.set_source_pixels(np.zeros((8, 8, 5), np.uint8))
    assert e.value.code == ssimu2.E_UNSUPPORTED


def test_source_samples_bit_exact(scorer, oracle):
    """the encoder-side depth conversions (io.zig:566-609) from the staged source, every channel kept."""
    rng = np.random.default_rng(11)
    for (w, h) in ((64, 64), (131, 77), (257, 9)):
        for ch in (1, 2, 3, 4):
            p8 = rng.integers(0, 256, (h, w, ch)).astype(np.uint8)
            p8.reshape(-1)[:256] = np.arange(256, dtype=np.uint8)[: min(256, p8.size)]     # every sample value once
            p16 = rng.integers(0, 65536, (h, w, ch)).astype(np.uint16)
            scorer.set_source_pixels(p8)
            got = scorer.source_samples(10)
            assert got.dtype == np.uint16 and got.shape == p8.shape
            np.testing.assert_array_equal(got, oracle.source_samples(p8, 10))
            with pytest.raises(ssimu2.Ssimu2Error) as e:
                scorer.source_samples(8)
            assert e.value.code == ssimu2.E_UNSUPPORTED
            scorer.set_source_pixels(p16)
            np.testing.assert_array_equal(scorer.source_samples(10), oracle.source_samples(p16, 10))
            np.testing.assert_array_equal(scorer.source_samples(8), oracle.source_samples(p16, 8))
            scorer.check_guards()
    # strided RGB8 through set_source_rgb8: the staged copy is tight
    base = synth.synth(150, 90, "mixture", 2)
    pad = np.zeros((90, 161, 3), np.uint8)
    pad[:, :150] = base
    scorer.set_source(pad[:, :150])
    np.testing.assert_array_equal(scorer.source_samples(10), oracle.source_samples(base, 10))
    # the cached source is untouched by the call
    dist = synth.distort(base, 0.3)
    a = scorer.score_rgb8(dist)
    scorer.source_samples(10)
    assert scorer.score_rgb8(dist) == a
    scorer.set_source(synth.synth(7, 7, "noise", 0))      # below 8x8 nothing is staged
    with pytest.raises(ssimu2.Ssimu2Error) as e:
        scorer.source_samples(10)
    assert e.value.code == ssimu2.E_STATE


# ---- K4: the filter alone ----------------------------------------------------------------------------
@pytest.mark.parametrize("mode,omode,name", MODES)
@pytest.mark.parametrize("size", [(9, 9), (40, 33), (131, 97), (640, 360)])
def test_blur_bit_identical(scorer, oracle, mode, omode, name, size):
    w, h = size
    rng = np.random.default_rng(w * 1000 + h)
    plane = rng.random((h, w), dtype=np.float32)
    scorer.set_blur(mode)
    np.testing.assert_array_equal(bits(scorer.blur(plane)), bits(oracle.blur(plane, omode)))


# ---- whole path ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode,omode,name", MODES)
@pytest.mark.parametrize("case", [(64, 64, "mixture", 0.3), (100, 75, "noise", 0.1), (257, 129, "edges", 0.5),
                                  (320, 240, "gradient", 0.05), (333, 257, "mixture", 1.0), (640, 360, "mixture", 0.2),
                                  (1027, 771, "noise", 0.02), (31, 200, "noise", 0.4), (8, 8, "noise", 0.5)])
def test_score_and_pooled_sums_vs_oracle(scorer, oracle, mode, omode, name, case):
    w, h, kind, strength = case
    src = synth.synth(w, h, kind, 21)
    dist = synth.distort(src, strength, seed=5)
    want, det = oracle.ssimu2_rgb8(src, dist, omode, detail=True)
    scorer.set_blur(mode)
    scorer.set_source(src)
    got = scorer.score_rgb8(dist)
    d = scorer.detail()
    assert d.n_scales == det.n_scales
    assert abs(got - want) <= SCORE_TOL, (got, want)
    gs, ws = scorer.sums(), oracle.detail_sums(det)
    np.testing.assert_allclose(gs, ws, rtol=SUM_RTOL, atol=1e-12)


@pytest.mark.parametrize("mode,omode,name", MODES)
def test_committed_oracle_scores(scorer, mode, omode, name, golden_dir):
    with open(os.path.join(golden_dir, "oracle_scores.json")) as f:
        cases = json.load(f)
    key = "iir" if mode == ssimu2.BLUR_RECURSIVE else "fir"
    scorer.set_blur(mode)
    for c in cases:
        src = synth.synth(c["w"], c["h"], c["kind"], c["seed"])
        dist = synth.distort(src, c["strength"], seed=c["seed"] + 100)
        scorer.set_source(src)
        got = scorer.score_rgb8(dist)
        assert abs(got - c[key]["score"]) <= SCORE_TOL, (c["w"], c["h"], got, c[key]["score"])
        assert scorer.detail().n_scales == c[key]["n_scales"]


@pytest.mark.parametrize("mode,omode,name", MODES)
def test_real_decoder_output(scorer, oracle, mode, omode, name, golden_dir):
    """Decoded planes of a real libavif/libaom round trip: scoring the planes directly equals scoring
    the RGB8 libavif made from them (io.zig:470-478), and both equal the oracle."""
    g = np.load(os.path.join(golden_dir, "avif_roundtrip.npz"))
    src = g["src"]
    scorer.set_blur(mode)
    scorer.set_source(src)
    for tag, depth, rgba in (("65", 8, False), ("40", 8, False), ("10", 10, False), ("10", 10, True)):
        y, u, v = g[f"y{tag}"], g[f"u{tag}"], g[f"v{tag}"]
        rgb = g["rgb10a"] if (tag == "10" and rgba) else g[f"rgb{tag}"]
        a = scorer.score_yuv444(y, u, v, depth, 2, rgba)
        b = scorer.score_rgb8(rgb)
        assert a == b
        assert abs(a - oracle.ssimu2_rgb8(src, rgb, omode)) <= SCORE_TOL
    assert scorer.score_yuv444(g["y65"], g["u65"], g["v65"], 8) > scorer.score_yuv444(g["y40"], g["u40"], g["v40"], 8)


# ---- edge cases ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode,omode,name", MODES)
def test_identical_pair_is_exactly_100(scorer, mode, omode, name):
    scorer.set_blur(mode)
    for (w, h) in [(64, 64), (333, 257), (1027, 771)]:
        src = synth.synth(w, h, "mixture", 3)
        scorer.set_source(src)
        assert scorer.score_rgb8(src) == 100.0
        assert not scorer.sums().any()


def test_below_8_pixels_scores_100_like_the_published_loop(scorer, oracle):
    src = synth.synth(7, 64, "noise", 6)
    dist = synth.distort(src, 0.4)
    assert oracle.ssimu2_rgb8(src, dist) == 100.0
    scorer.set_source(src)
    assert scorer.score_rgb8(dist) == 100.0
    assert scorer.detail().n_scales == 0


@pytest.mark.parametrize("mode,omode,name", MODES)
def test_strided_rows(scorer, mode, omode, name):
    w, h = 150, 90
    src = synth.synth(w, h, "mixture", 1)
    dist = synth.distort(src, 0.3)
    scorer.set_blur(mode)
    scorer.set_source(src)
    tight = scorer.score_rgb8(dist)
    pad = np.zeros((h, w + 9, 3), np.uint8)
    pad[:, :w] = dist
    assert scorer.score_rgb8(pad[:, :w]) == tight          # row stride 3*(w+9)
    spad = np.full((h, w + 5, 3), 77, np.uint8)
    spad[:, :w] = src
    scorer.set_source(spad[:, :w])
    assert scorer.score_rgb8(dist) == tight
    y, u, v = synth.rgb8_to_yuv444(dist, 10)
    t2 = scorer.score_yuv444(y, u, v, 10)
    big = np.zeros((3, h, w + 6), np.uint16)
    big[0, :, :w], big[1, :, :w], big[2, :, :w] = y, u, v
    assert scorer.score_yuv444(big[0, :, :w], big[1, :, :w], big[2, :, :w], 10) == t2


@pytest.mark.parametrize("mode,omode,name", MODES)
def test_batch_equals_single_and_is_deterministic(scorer, mode, omode, name):
    w, h = 400, 300
    src = synth.synth(w, h, "mixture", 9)
    cands = [synth.distort(src, s, seed=i) for i, s in enumerate((0.1, 0.3, 0.6, 1.0))]
    scorer.set_blur(mode)
    scorer.set_source(src)
    single = [scorer.score_rgb8(c) for c in cands]
    batch = scorer.score_batch_rgb8(cands)
    assert batch == single                                   # bit-identical doubles
    sums0 = scorer.sums(2).copy()
    again = scorer.score_batch_rgb8(cands)
    assert again == batch and (scorer.sums(2) == sums0).all()
    yuv = [synth.rgb8_to_yuv444(c, 10) for c in cands]
    by = scorer.score_batch_yuv444(yuv, 10)
    assert by == [scorer.score_yuv444(*p, 10) for p in yuv]
    assert all(a > b for a, b in zip(batch, batch[1:]))


def test_source_rows_cache_follows_the_source_and_the_blur_mode(scorer, oracle):
    """The rows pass of the source-only quantities is cached per source: it must be rebuilt for a new
    source, built late when the source was set under the other blur, and shared by a batch."""
    a = synth.synth(300, 200, "mixture", 1)
    b = synth.synth(300, 200, "noise", 2)
    da, db = synth.distort(a, 0.3), synth.distort(b, 0.3)
    scorer.set_blur(ssimu2.BLUR_RECURSIVE)
    scorer.set_source(a)
    s_a = scorer.score_rgb8(da)
    scorer.set_source(b)                       # new source right away: cache must not go stale
    s_b = scorer.score_rgb8(db)
    scorer.set_source(a)
    scorer.set_source(b)                       # two set_source calls back to back, no score between
    assert scorer.score_rgb8(db) == s_b
    scorer.set_blur(ssimu2.BLUR_FIR)
    scorer.set_source(a)                       # set under FIR ...
    scorer.set_blur(ssimu2.BLUR_RECURSIVE)
    assert scorer.score_rgb8(da) == s_a        # ... scored under RECURSIVE: cache built on demand
    assert scorer.score_batch_rgb8([da, a, da]) == [s_a, 100.0, s_a]
    assert abs(s_a - oracle.ssimu2_rgb8(a, da)) <= SCORE_TOL and abs(s_b - oracle.ssimu2_rgb8(b, db)) <= SCORE_TOL


def test_tile_paths_give_the_same_bits(scorer):
    """TMA, cp.async, fused and barrier-free-columns forms of the recursive kernels: identical pooled sums and scores,
    single and batch."""
    src = synth.synth(1000, 700, "mixture", 14)
    cands = [synth.distort(src, s, seed=i) for i, s in enumerate((0.2, 0.5, 0.9))]
    scorer.set_blur(ssimu2.BLUR_RECURSIVE)
    res = {}
    for path in (ssimu2.TILES_TMA, ssimu2.TILES_CP_ASYNC, ssimu2.TILES_FUSED, ssimu2.TILES_TMA_DECOUPLED):
        scorer.set_tile_path(path)
        scorer.set_source(src)
        single = [scorer.score_rgb8(c) for c in cands]
        scorer.set_source(src)
        batch = scorer.score_batch_rgb8(cands)
        res[path] = (single, batch, [scorer.sums(i).copy() for i in range(3)])
        assert scorer.get_option(ssimu2.OPT_TILE_PATH) == path
    scorer.set_tile_path(ssimu2.TILES_TMA)
    a = res[ssimu2.TILES_TMA]
    for other in (ssimu2.TILES_CP_ASYNC, ssimu2.TILES_FUSED, ssimu2.TILES_TMA_DECOUPLED):
        b = res[other]
        assert a[0] == b[0] == a[1] == b[1], other
        for x, y in zip(a[2], b[2]):
            np.testing.assert_array_equal(x, y)


def test_modes_differ_only_by_recursion_roundoff(scorer):
    src = synth.synth(640, 360, "mixture", 4)
    dist = synth.distort(src, 0.6)
    scorer.set_source(src)
    scorer.set_blur(ssimu2.BLUR_RECURSIVE)
    a = scorer.score_rgb8(dist)
    scorer.set_blur(ssimu2.BLUR_FIR)
    b = scorer.score_rgb8(dist)
    assert a != b and abs(a - b) < 0.5


def test_error_behaviour(oracle):
    src = synth.synth(64, 48, "noise", 0)
    with ssimu2.Scorer(64, 48, 2) as sc:
        with pytest.raises(ssimu2.Ssimu2Error) as e:
            sc.score_rgb8(src)
        assert e.value.code == ssimu2.E_ARG or e.value.code == ssimu2.E_STATE
        L = ssimu2.load()
        out = C.c_double()
        assert L.oavif_ssimu2_score_rgb8(sc._ctx, src.ctypes.data, 192, C.byref(out)) == ssimu2.E_STATE
        with pytest.raises(ssimu2.Ssimu2Error) as e:
            sc.set_source(synth.synth(512, 512, "noise", 0))
        assert e.value.code == ssimu2.E_STATE
        sc.set_source(src)
        assert L.oavif_ssimu2_score_rgb8(sc._ctx, src.ctypes.data, 100, C.byref(out)) == ssimu2.E_ARG
        assert L.oavif_ssimu2_score_rgb8(sc._ctx, None, 192, C.byref(out)) == ssimu2.E_ARG
        y = np.zeros((48, 64), np.uint8)
        with pytest.raises(ssimu2.Ssimu2Error) as e:
            sc.score_yuv444(y, y, y, 8, matrix=0)
        assert e.value.code == ssimu2.E_UNSUPPORTED
        with pytest.raises(ssimu2.Ssimu2Error) as e:
            sc.score_yuv444(y.astype(np.uint16), y.astype(np.uint16), y.astype(np.uint16), 12)
        assert e.value.code == ssimu2.E_UNSUPPORTED
        with pytest.raises(ssimu2.Ssimu2Error):
            sc.score_batch_rgb8([src, src, src])             # beyond max_batch
        assert sc.score_rgb8(src) == 100.0                   # context still usable after errors
    with pytest.raises(ssimu2.Ssimu2Error):
        ssimu2.Scorer(64, 64, device=99)


def test_stateless_call_mirrors_the_reference_signature(oracle):
    src = synth.synth(320, 200, "mixture", 12)
    dist = synth.distort(src, 0.4)
    want = oracle.ssimu2_rgb8(src, dist, oracle.BLUR_IIR)
    got = ssimu2.compute_ssimu2(src.reshape(-1), dist.reshape(-1), 320, 200, 3)
    assert abs(got - want) <= SCORE_TOL
    assert ssimu2.compute_ssimu2(src, src) == 100.0
    big = synth.synth(500, 400, "edges", 1)                  # cached context regrows
    assert ssimu2.compute_ssimu2(big, big) == 100.0
    with pytest.raises(ssimu2.Ssimu2Error):
        ssimu2.compute_ssimu2(np.zeros((4, 4, 4), np.uint8), np.zeros((4, 4, 4), np.uint8))


def test_device_resident_inputs_and_external_stream(oracle):
    import torch
    w, h = 256, 192
    src = synth.synth(w, h, "mixture", 2)
    dist = synth.distort(src, 0.3)
    y, u, v = synth.rgb8_to_yuv444(dist, 10)
    with ssimu2.Scorer(w, h, 2) as sc:
        sc.set_source(src)
        host = sc.score_yuv444(y, u, v, 10)
        st = torch.cuda.Stream()
        sc.set_stream(st.cuda_stream)
        dsrc = torch.from_numpy(src).cuda()
        dy, du, dv = (torch.from_numpy(p.view(np.int16)).cuda() for p in (y, u, v))
        drgb = torch.from_numpy(dist).cuda()
        torch.cuda.synchronize()
        sc.set_source_dev(dsrc.data_ptr(), w, h, 3 * w)
        a = sc.score_batch_dev("yuv444", [[dy.data_ptr(), du.data_ptr(), dv.data_ptr()]] * 2, [2 * w] * 3, depth=10)
        assert a == [host, host]
        b = sc.score_batch_dev("rgb8", [[drgb.data_ptr()]], [3 * w])
        assert abs(b[0] - oracle.ssimu2_rgb8(src, dist)) <= SCORE_TOL
        # the encoder-side samples of a DEVICE source come straight from the caller's buffer (still alive here)
        np.testing.assert_array_equal(sc.source_samples(10), oracle.source_samples(src, 10))
        sc.set_stream(None)


@pytest.mark.parametrize("size", [(8, 8), (9, 65), (63, 64), (64, 65), (65, 127), (129, 31), (250, 250), (333, 511)])
def test_no_kernel_writes_past_its_buffers(size):
    """compute-sanitizer is closed on the target pool: contexts sized EXACTLY for the image, every
    buffer followed by a guard band, all entry points exercised, bands verified."""
    w, h = size
    src = synth.synth(w, h, "noise", w + h)
    d1, d2 = synth.distort(src, 0.2), synth.distort(src, 0.7)
    with ssimu2.Scorer(w, h, 2) as sc:
        for mode, path in ((ssimu2.BLUR_RECURSIVE, ssimu2.TILES_TMA), (ssimu2.BLUR_RECURSIVE, ssimu2.TILES_CP_ASYNC),
                           (ssimu2.BLUR_RECURSIVE, ssimu2.TILES_FUSED), (ssimu2.BLUR_RECURSIVE, ssimu2.TILES_TMA_DECOUPLED),
                           (ssimu2.BLUR_FIR, ssimu2.TILES_TMA)):
            sc.set_blur(mode)
            sc.set_tile_path(path)
            sc.set_source(src)
            sc.score_batch_rgb8([d1, d2])
            for depth in (8, 10):
                y, u, v = synth.rgb8_to_yuv444(d1, depth)
                sc.score_yuv444(y, u, v, depth, 2, depth == 10)
            sc.set_source_pixels(np.dstack([src, src[..., :1]]).astype(np.uint16) * 257)
            sc.score_pixels(np.dstack([d2, d2[..., :1]]))
            sc.check_guards()
        y, u, v = synth.rgb8_to_yuv444(d1, 10)
        sc.yuv444_to_rgb8(y, u, v, 10)
        sc.check_guards()


# ---- BASELINE.json full sizes: properties that do not need the oracle's minutes ----------------------------
@pytest.mark.parametrize("mode,omode,name", MODES)
def test_full_size_4k_properties(oracle, mode, omode, name):
    w, h = 3840, 2160
    src = synth.synth(w, h, "mixture", 0)
    d1, d2 = synth.distort(src, 0.2, seed=1), synth.distort(src, 0.6, seed=2)
    with ssimu2.Scorer(w, h, 3, blur=mode) as sc:
        sc.set_source(src)
        assert sc.score_rgb8(src) == 100.0
        s = sc.score_batch_rgb8([d1, d2, d1])
        assert s[0] == s[2] and s[0] > s[1]
        y, u, v = synth.rgb8_to_yuv444(d1, 10)
        rgb = oracle.yuv444_to_rgb8(y, u, v, 10)
        assert sc.score_yuv444(y, u, v, 10) == sc.score_rgb8(rgb)
        # swapping which image is the source changes artifact <-> detail_lost sums only
        sc.score_rgb8(d1)
        sums_ab = sc.sums(0).copy()
        sc.set_source(d1)
        sc.score_rgb8(src)
        sums_ba = sc.sums(0)
        np.testing.assert_allclose(sums_ab[:, 0::6], sums_ba[:, 0::6], rtol=1e-12)  # SSIM term is symmetric
        np.testing.assert_allclose(sums_ab[:, 1::6], sums_ba[:, 1::6], rtol=1e-12)
        if mode == ssimu2.BLUR_RECURSIVE:                                           # one oracle run at full size
            sc.set_source(src)
            got = sc.score_rgb8(d2)
            assert abs(got - oracle.ssimu2_rgb8(src, d2, omode, fast=True)) <= SCORE_TOL


# ---- BASELINE.json configs 3, 4, 5 and config 2 in FIR mode: CUDA path vs the oracle at the full sizes --------
def _tiled(img, ny, nx):
    """A big image from a small procedural one (the generator is numpy and costs ~1 s/Mpx): tiles plus a smooth
    image-wide ramp so that no two tiles carry the same pixels."""
    big = np.tile(img, (ny, nx, 1)).astype(np.int16)
    hh, ww = big.shape[:2]
    ramp = (np.arange(ww, dtype=np.int32)[None, :] * 24 // ww + np.arange(hh, dtype=np.int32)[:, None] * 16 // hh)
    big[..., :3] += ramp[..., None].astype(np.int16) - 20
    return np.clip(big, 0, 255).astype(np.uint8)


@pytest.mark.slow
@pytest.mark.parametrize("kind", ["gradient", "edges", "noise", "mixture"])
def test_cfg5_1080p_every_kind_vs_oracle(oracle, kind):
    w, h = 1920, 1080
    src = synth.synth(w, h, kind, 5)
    dist = synth.distort(src, 0.3, seed=6)
    y, u, v = synth.rgb8_to_yuv444(dist, 10)
    rgb = oracle.yuv444_to_rgb8(y, u, v, 10)
    with ssimu2.Scorer(w, h, 1) as sc:
        for mode, omode, _ in MODES:
            sc.set_blur(mode)
            sc.set_source(src)
            got = sc.score_yuv444(y, u, v, 10)
            want, det = oracle.ssimu2_rgb8(src, rgb, omode, fast=True, detail=True)
            assert abs(got - want) <= SCORE_TOL, (kind, mode, got, want)
            np.testing.assert_allclose(sc.sums(), oracle.detail_sums(det), rtol=SUM_RTOL, atol=1e-12)


@pytest.mark.slow
def test_cfg2_4k_fir_vs_oracle(oracle):
    w, h = 3840, 2160
    src = _tiled(synth.synth(1920, 1080, "mixture", 0), 2, 2)
    dist = _tiled(synth.distort(synth.synth(1920, 1080, "mixture", 0), 0.4, seed=3), 2, 2)
    with ssimu2.Scorer(w, h, 1, blur=ssimu2.BLUR_FIR) as sc:
        sc.set_source(src)
        got = sc.score_rgb8(dist)
        want, det = oracle.ssimu2_rgb8(src, dist, oracle.BLUR_FIR, fast=True, detail=True)
        assert abs(got - want) <= SCORE_TOL, (got, want)
        np.testing.assert_allclose(sc.sums(), oracle.detail_sums(det), rtol=SUM_RTOL, atol=1e-12)


@pytest.mark.slow
def test_cfg3_24mp_batch16_vs_oracle(oracle):
    """Config 3: sixteen candidate decodes of one 6000x4000 image in one launch; candidates 0, 5 and 15 are
    checked against the oracle, the rest against their duplicates (four distinct frames cycle through the batch)."""
    w, h = 6000, 4000
    small = synth.synth(2000, 1000, "mixture", 3)
    src = _tiled(small, 4, 3)
    frames = [_tiled(synth.distort(small, s, seed=10 + i), 4, 3) for i, s in enumerate((0.1, 0.25, 0.5, 0.9))]
    cands = [frames[i % 4] for i in range(16)]
    with ssimu2.Scorer(w, h, 16) as sc:
        sc.set_source(src)
        got = sc.score_batch_rgb8(cands)
        assert all(got[i] == got[i % 4] for i in range(16))
        assert got[0] > got[1] > got[2] > got[3]
        for i in (0, 5, 15):
            want, det = oracle.ssimu2_rgb8(src, cands[i], oracle.BLUR_IIR, fast=True, detail=True)
            assert abs(got[i] - want) <= SCORE_TOL, (i, got[i], want)
            np.testing.assert_allclose(sc.sums(i), oracle.detail_sums(det), rtol=SUM_RTOL, atol=1e-12)
        sc.check_guards()


@pytest.mark.slow
def test_cfg4_8k_rgba_i410_path_vs_oracle(oracle):
    """Config 4: 7680x4320 RGBA source (alpha never scored, io.zig:106-111), candidate = 10-bit planes of an image
    that carries an alpha plane, i.e. libavif's I410 conversion (io.zig:473)."""
    w, h = 7680, 4320
    small = synth.synth_rgba(1920, 1080, "mixture", 7)
    src_rgba = _tiled(small, 4, 4)
    dist = _tiled(synth.distort(small[..., :3], 0.3, seed=8), 4, 4)
    y, u, v = synth.rgb8_to_yuv444(dist, 10)
    src_rgb = oracle.to_rgb8(src_rgba, 4, False)
    dist_rgb = oracle.yuv444_to_rgb8(y, u, v, 10, 2, True)
    with ssimu2.Scorer(w, h, 1) as sc:
        sc.set_source_pixels(src_rgba)
        got = sc.score_yuv444(y, u, v, 10, 2, True)
        want, det = oracle.ssimu2_rgb8(src_rgb, dist_rgb, oracle.BLUR_IIR, fast=True, detail=True)
        assert abs(got - want) <= SCORE_TOL, (got, want)
        np.testing.assert_allclose(sc.sums(), oracle.detail_sums(det), rtol=SUM_RTOL, atol=1e-12)
        assert sc.score_yuv444(y, u, v, 10, 2, False) != got      # the RGB and RGBA conversions are different functions
        sc.check_guards()


# ---- context capacity: an image that fits the pyramid and input buffers can still need more CTAs ---------------
@pytest.mark.parametrize("ctx_size,img_size", [((100, 2426), (4509, 53)), ((3767, 3785), (3717, 3835)),
                                               ((640, 360), (360, 640)), ((1000, 200), (500, 390))])
def test_context_sized_for_another_shape(ctx_size, img_size):
    """Per-CTA partial sums are indexed by CTA: an image of another shape must either be refused (E_STATE) or be
    scored without touching a guard band, never silently overrun (the contexts above accept / refuse both ways)."""
    w, h = img_size
    rng = np.random.default_rng(w)
    src = rng.integers(0, 255, (h, w, 3), dtype=np.uint8)
    dist = np.clip(src.astype(np.int16) + rng.integers(-9, 9, src.shape), 0, 255).astype(np.uint8)
    with ssimu2.Scorer(*ctx_size, 2) as sc:
        for mode in (ssimu2.BLUR_RECURSIVE, ssimu2.BLUR_FIR):
            sc.set_blur(mode)
            try:
                sc.set_source(src)
            except ssimu2.Ssimu2Error as e:
                assert e.code == ssimu2.E_STATE
                continue
            a = sc.score_batch_rgb8([dist, src])
            sc.check_guards()
            with ssimu2.Scorer(w, h, 2, blur=mode) as exact:
                exact.set_source(src)
                assert exact.score_batch_rgb8([dist, src]) == a


def test_weight_layout_option_matches_the_oracle_variant(scorer, oracle):
    """Images with fewer than six scales: the two readings of the final sum (include/oavif_ssimu2.h,
    OAVIF_SSIMU2_OPT_WEIGHTS) against the oracle's two variants; with six scales they coincide."""
    for (w, h) in [(100, 75), (31, 200), (640, 360)]:
        src = synth.synth(w, h, "mixture", 9)
        dist = synth.distort(src, 0.4)
        scorer.set_blur(ssimu2.BLUR_RECURSIVE)
        scorer.set_source(src)
        try:
            six = scorer.score_rgb8(dist)
            scorer.set_weights(ssimu2.WEIGHTS_CONTIGUOUS)
            oracle.set_variant(oracle.VARIANT_CONTIGUOUS_WEIGHTS)
            contiguous = scorer.score_rgb8(dist)
            want_c = oracle.ssimu2_rgb8(src, dist)
        finally:
            scorer.set_weights(ssimu2.WEIGHTS_SIX_SLOTS)
            oracle.set_variant(0)
        assert abs(six - oracle.ssimu2_rgb8(src, dist)) <= SCORE_TOL
        assert abs(contiguous - want_c) <= SCORE_TOL
        assert (six == contiguous) == (scorer.detail().n_scales == 6)


def test_transfer_table_option_matches_the_oracle_variant(scorer, oracle):
    """OAVIF_SSIMU2_OPT_TRANSFER: the sRGB table from binary32 powf against the oracle's F32_TRANSFER variant — XYB bit
    for bit, the score to the usual tolerance — and back; switching drops the cached source."""
    w, h = 333, 257
    src = synth.synth(w, h, "mixture", 21)
    dist = synth.distort(src, 0.5)
    scorer.set_blur(ssimu2.BLUR_RECURSIVE)
    scorer.set_source(src)
    f64 = scorer.score_rgb8(dist)
    try:
        scorer.set_option(ssimu2.OPT_TRANSFER, ssimu2.TRANSFER_F32)
        assert scorer.get_option(ssimu2.OPT_TRANSFER) == ssimu2.TRANSFER_F32
        with pytest.raises(ssimu2.Ssimu2Error) as e:
            scorer.score_rgb8(dist)                       # the cached pyramid came from the other table
        assert e.value.code == ssimu2.E_STATE
        oracle.set_variant(oracle.VARIANT_F32_TRANSFER)
        scorer.set_source(src)
        f32 = scorer.score_rgb8(dist)
        assert abs(f32 - oracle.ssimu2_rgb8(src, dist)) <= SCORE_TOL
        want = oracle.xyb_at_scale(src, 0)
        for c in range(3):
            np.testing.assert_array_equal(bits(scorer.xyb(0, 0, c)), bits(want[c]))
    finally:
        oracle.set_variant(0)
        scorer.set_option(ssimu2.OPT_TRANSFER, ssimu2.TRANSFER_F64)
    scorer.set_source(src)
    assert scorer.score_rgb8(dist) == f64 and f32 != f64
    assert abs(f64 - oracle.ssimu2_rgb8(src, dist)) <= SCORE_TOL


@pytest.mark.parametrize("size", [(100, 75), (333, 257), (640, 360)])
def test_vertical_order_option_matches_the_oracle_variant(scorer, oracle, size):
    """OAVIF_SSIMU2_OPT_VERTICAL_ORDER: fma(n2, sum, fma(-d1, y1, -y2)) in the columns pass against the oracle's
    VERTICAL_ORDER variant — the five blurred values the product kernel hands to the maps bit for bit, the score to the
    usual tolerance; refused on the tile paths that do not carry the instance."""
    w, h = size
    src = synth.synth(w, h, "mixture", 33)
    dist = synth.distort(src, 0.5)
    scorer.set_blur(ssimu2.BLUR_RECURSIVE)
    scorer.set_source(src)
    default = scorer.score_rgb8(dist)
    try:
        scorer.set_option(ssimu2.OPT_VERTICAL_ORDER, ssimu2.VERTICAL_FUSED_OUTER)
        oracle.set_variant(oracle.VARIANT_VERTICAL_ORDER)
        got = scorer.score_rgb8(dist)
        assert abs(got - oracle.ssimu2_rgb8(src, dist)) <= SCORE_TOL
        a, b = oracle.xyb_at_scale(src, 1)[1], oracle.xyb_at_scale(dist, 1)[1]
        want = [oracle.blur(p) for p in (a, b, a * a, b * b, a * b)]
        cols = scorer.cols(0, 1, 1)
        for q in range(5):
            np.testing.assert_array_equal(bits(cols[q]), bits(want[q]), err_msg=f"quantity {q}")
        scorer.set_tile_path(ssimu2.TILES_CP_ASYNC)
        with pytest.raises(ssimu2.Ssimu2Error) as e:
            scorer.score_rgb8(dist)
        assert e.value.code == ssimu2.E_UNSUPPORTED
    finally:
        oracle.set_variant(0)
        scorer.set_tile_path(ssimu2.TILES_TMA)
        scorer.set_option(ssimu2.OPT_VERTICAL_ORDER, ssimu2.VERTICAL_AS_HORIZONTAL)
    assert scorer.score_rgb8(dist) == default
    assert abs(default - oracle.ssimu2_rgb8(src, dist)) <= SCORE_TOL


def test_conversion_call_leaves_the_cached_source_alone(scorer):
    src = synth.synth(200, 120, "mixture", 3)
    dist = synth.distort(src, 0.3)
    y, u, v = synth.rgb8_to_yuv444(dist, 10)
    scorer.set_blur(ssimu2.BLUR_RECURSIVE)
    scorer.set_source(src)
    a = scorer.score_yuv444(y, u, v, 10)
    rgb = scorer.yuv444_to_rgb8(y, u, v, 10)          # in the middle of a search
    assert scorer.score_yuv444(y, u, v, 10) == a
    assert scorer.score_rgb8(rgb) == a


# ---- pipelined form: submit / wait ---------------------------------------------------------------------------
def test_submit_wait_equals_the_synchronous_calls():
    """Two submissions in flight, set_source of the next image while the previous one is still being scored, results
    retired in order and bit-identical to the synchronous calls."""
    w, h = 640, 360
    imgs = [synth.synth(w, h, "mixture", 60 + k) for k in range(4)]
    cands = [[synth.distort(s, 0.15 + 0.2 * i, seed=i) for i in range(2)] for s in imgs]
    yuv = [[synth.rgb8_to_yuv444(c, 10) for c in cs] for cs in cands]
    with ssimu2.Scorer(w, h, 2) as sc:
        want, want_sums = [], []
        for s, ys in zip(imgs, yuv):
            sc.set_source(s)
            want.append(sc.score_batch_yuv444(ys, 10))
            want_sums.append(sc.sums(1).copy())
        for mode in (ssimu2.BLUR_RECURSIVE,):
            got = []
            for k, (s, ys) in enumerate(zip(imgs, yuv)):
                sc.set_source(s)                      # image k+1 goes up while image k is in flight
                sc.submit_yuv444(ys, 10)
                assert sc.in_flight() == (1 if k == 0 else 2)
                if k > 0:
                    got.append(sc.wait())
                    np.testing.assert_array_equal(sc.sums(1), want_sums[k - 1])
            got.append(sc.wait())
            assert got == want and sc.in_flight() == 0
        # rgb8 form, two submissions against ONE source, then a third submit must be refused until a wait
        sc.set_source(imgs[0])
        want_rgb = sc.score_rgb8(cands[0][0])
        sc.submit_rgb8([cands[0][0]])
        sc.submit_rgb8([cands[0][1], cands[0][0]])
        with pytest.raises(ssimu2.Ssimu2Error) as e:
            sc.submit_rgb8([cands[0][0]])
        assert e.value.code == ssimu2.E_STATE
        with pytest.raises(ssimu2.Ssimu2Error):
            sc.score_rgb8(cands[0][0])                # synchronous call while submissions are in flight
        a, b = sc.wait(), sc.wait()
        assert a == [b[1]] == [want_rgb]
        with pytest.raises(ssimu2.Ssimu2Error):
            sc.wait()
        assert sc.score_rgb8(imgs[0]) == 100.0
        sc.check_guards()


def test_every_visible_device_scores_the_same():
    """The corpus driver runs contexts on several devices from one process: kernel attributes are per device."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one visible GPU")
    src = synth.synth(640, 360, "mixture", 2)
    dist = synth.distort(src, 0.4)
    got = []
    for dev in range(n):
        for mode in (ssimu2.BLUR_RECURSIVE, ssimu2.BLUR_FIR):
            with ssimu2.Scorer(640, 360, 1, device=dev, blur=mode) as sc:
                sc.set_source(src)
                got.append((mode, sc.score_rgb8(dist)))
    assert len(set(got)) == 2


def test_cuda_against_fssimu2_vectors():
    """The CUDA path against vectors of the reference's own scorer, when a machine with zig has produced them
    (scripts/pin_fssimu2.md).  Skipped — "parity unpinned" — while tests/golden/fssimu2_scores.json is absent."""
    from test_oracle import fssimu2_vectors
    cases = fssimu2_vectors()
    for c in cases:
        src = synth.synth(c["w"], c["h"], c["kind"], c["seed"])
        dst = synth.distort(src, c["strength"], seed=c["seed"] + 100)
        with ssimu2.Scorer(c["w"], c["h"], 1) as sc:
            sc.set_source(src)
            assert abs(sc.score_rgb8(dst) - c["score"]) <= 0.05, c          # north_star's bar


# ---- several callers per GPU: one context per host thread (the corpus driver's workers-per-gpu) ------------
def test_two_contexts_on_two_threads_score_like_one():
    """Contexts are single-owner, but several may run on one GPU at once (each on its own stream): concurrent
    callers must get bit-identical scores to a lone caller."""
    import threading
    w, h = 640, 360
    srcs = [synth.synth(w, h, "mixture", 40 + k) for k in range(2)]
    cands = [[synth.distort(s, 0.2 + 0.2 * i, seed=i) for i in range(3)] for s in srcs]
    with ssimu2.Scorer(w, h, 1) as sc:
        want = []
        for s, cs in zip(srcs, cands):
            sc.set_source(s)
            want.append([sc.score_rgb8(c) for c in cs])
    got = [None, None]

    def work(k):
        with ssimu2.Scorer(w, h, 1) as sck:
            out = []
            for _ in range(10):           # repeat: the two threads stay in flight together
                sck.set_source(srcs[k])
                out = [sck.score_rgb8(c) for c in cands[k]]
            got[k] = out

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert got == want
