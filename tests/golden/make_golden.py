#!/usr/bin/env python
"""Regenerates the fixtures in tests/golden/ (run in the build container; needs Pillow's libavif).

  yuv2rgb_libavif.npz   random YUV444 planes (8/10 bit) and the RGB8 that libavif 1.4.1's
                        avifImageYUVToRGB produced for every (matrix, alpha-plane) variant — the
                        function the reference calls at src/io.zig:478.  REFERENCE-PINNED.
  avif_roundtrip.npz    a 160x120 synthetic image encoded by libavif/libaom (8-bit YUV444, speed 9,
                        tune=iq; q=65 and q=40) and decoded again: decoded YUV planes + libavif's RGB8
                        for them, plus a 10-bit plane set (synthetic: the bundled libaom is built
                        without high-bit-depth ENCODE support, so a real 10-bit round trip cannot be
                        made here) with libavif's RGB8 for both the RGB and the RGBA conversion.
                        REFERENCE-PINNED for the conversion; also the realistic "distorted" input of
                        the parity tests.
  oracle_scores.json    scores and pooled sums of the CPU oracle on seeded synthetic pairs.
                        SELF-GOLDEN: guards the oracle against regressions; it pins nothing against
                        fssimu2 (parity unpinned, see oracle/ssimu2_oracle.h).
"""
import json
import os
import sys
import ctypes as C

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import avif_ctypes as A  # noqa: E402
from oavif_b200.host import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def golden_yuv():
    rng = np.random.default_rng(20261018)
    out = {}
    for depth in (8, 10):
        dt = np.uint8 if depth == 8 else np.uint16
        y, u, v = (rng.integers(0, 1 << depth, (40, 52)).astype(dt) for _ in range(3))
        # make sure the extremes are present
        y[0, :4] = [0, (1 << depth) - 1, 0, (1 << depth) - 1]
        u[0, :4] = [0, 0, (1 << depth) - 1, (1 << depth) - 1]
        v[0, :4] = [(1 << depth) - 1, 0, (1 << depth) - 1, 0]
        out[f"y{depth}"], out[f"u{depth}"], out[f"v{depth}"] = y, u, v
        for m in (1, 2, 5, 6, 9):
            for al in (0, 1):
                out[f"rgb_d{depth}_m{m}_a{al}"] = A.yuv444_to_rgb8(y, u, v, depth, m, bool(al))
    np.savez_compressed(os.path.join(HERE, "yuv2rgb_libavif.npz"), **out)
    print("yuv2rgb_libavif.npz", len(out), "arrays")


def encode_decode(rgb, q=65, depth=10, speed=9):
    """encodeAvifToBuffer + decode (io.zig:544-636, 452-466) with the reference's defaults."""
    L = A.lib()
    h, w, _ = rgb.shape
    img = L.avifImageCreate(w, h, depth, 1)
    for off in ("colorPrimaries", "transferCharacteristics", "matrixCoefficients"):
        C.c_uint16.from_address(img + A.OFF[off]).value = 2
    r = A.RGBImage()
    L.avifRGBImageSetDefaults(C.byref(r), img)
    r.format = 0
    r.depth = depth
    assert L.avifRGBImageAllocatePixels(C.byref(r)) == 0
    px = np.ctypeslib.as_array((C.c_uint8 * (r.rowBytes * h)).from_address(r.pixels)).reshape(h, r.rowBytes)
    if depth == 8:
        px[:, : 3 * w] = rgb.reshape(h, 3 * w)
    else:
        v16 = ((rgb.astype(np.uint32) * 1023 + 127) // 255).astype(np.uint16)  # io.zig:572
        px[:, : 6 * w] = v16.reshape(h, 3 * w).view(np.uint8)
    assert L.avifImageRGBToYUV(img, C.byref(r)) == 0
    L.avifRGBImageFreePixels(C.byref(r))
    enc = L.avifEncoderCreate()
    C.c_int.from_address(enc + A.ENC["maxThreads"]).value = 1
    C.c_int.from_address(enc + A.ENC["speed"]).value = speed
    C.c_int.from_address(enc + A.ENC["quality"]).value = q
    C.c_int.from_address(enc + A.ENC["qualityAlpha"]).value = 0
    C.c_int.from_address(enc + A.ENC["autoTiling"]).value = 1
    L.avifEncoderSetCodecSpecificOption(enc, b"tune", b"iq")
    assert L.avifEncoderAddImage(enc, img, 1, 2) == 0  # AVIF_ADD_IMAGE_FLAG_SINGLE
    data = A.RWData()
    assert L.avifEncoderFinish(enc, C.byref(data)) == 0
    blob = C.string_at(data.data, data.size)
    L.avifRWDataFree(C.byref(data))
    L.avifEncoderDestroy(enc)
    L.avifImageDestroy(img)

    dec = L.avifDecoderCreate()
    assert L.avifDecoderSetIOMemory(dec, blob, len(blob)) == 0
    assert L.avifDecoderParse(dec) == 0
    assert L.avifDecoderNextImage(dec) == 0
    dimg = None
    for off in A.DEC_IMAGE_CANDIDATES:  # locate decoder->image by its width/height/depth
        p = C.c_void_p.from_address(dec + off).value
        if p and p > 0x10000:
            try:
                if A._u32(p, 0) == w and A._u32(p, 4) == h and A._u32(p, 8) == depth:
                    dimg = p
                    break
            except Exception:
                pass
    assert dimg, "decoder->image not found"
    bps = 1 if depth == 8 else 2
    planes = []
    for i in range(3):
        src = A._ptr(dimg, A.OFF["yuvPlanes"] + 8 * i)
        rb = A._u32(dimg, A.OFF["yuvRowBytes"] + 4 * i)
        raw = np.ctypeslib.as_array((C.c_uint8 * (rb * h)).from_address(src)).reshape(h, rb)[:, : w * bps]
        planes.append(raw.copy().view(np.uint8 if depth == 8 else np.uint16).reshape(h, w))
    mc = C.c_uint16.from_address(dimg + A.OFF["matrixCoefficients"]).value
    rgb_out = A.image_to_rgb8(dimg)
    L.avifDecoderDestroy(dec)
    return blob, planes, mc, rgb_out


def golden_roundtrip():
    src = synth.synth(160, 120, "mixture", 7)
    blob65, (y65, u65, v65), mc65, rgb65 = encode_decode(src, 65, 8)
    blob40, (y40, u40, v40), mc40, rgb40 = encode_decode(src, 40, 8)
    y10, u10, v10 = synth.rgb8_to_yuv444(rgb40, 10, 2)
    rgb10 = A.yuv444_to_rgb8(y10, u10, v10, 10, 2, False)
    rgb10a = A.yuv444_to_rgb8(y10, u10, v10, 10, 2, True)
    np.savez_compressed(os.path.join(HERE, "avif_roundtrip.npz"), src=src,
                        y65=y65, u65=u65, v65=v65, matrix65=mc65, rgb65=rgb65, avif_bytes65=len(blob65),
                        y40=y40, u40=u40, v40=v40, matrix40=mc40, rgb40=rgb40, avif_bytes40=len(blob40),
                        y10=y10, u10=u10, v10=v10, rgb10=rgb10, rgb10a=rgb10a)
    print("avif_roundtrip.npz", len(blob65), len(blob40), "matrix", mc65, mc40)


def golden_scores():
    cases = []
    for (w, h, kind, seed, strength) in [(64, 64, "mixture", 0, 0.3), (100, 75, "noise", 1, 0.1),
                                          (257, 129, "edges", 2, 0.5), (320, 240, "gradient", 3, 0.2),
                                          (333, 257, "mixture", 4, 1.0), (15, 40, "noise", 5, 0.4),
                                          (7, 64, "noise", 6, 0.4)]:
        src = synth.synth(w, h, kind, seed)
        dist = synth.distort(src, strength, seed=seed + 100)
        rec = dict(w=w, h=h, kind=kind, seed=seed, strength=strength)
        for name, mode in (("iir", O.BLUR_IIR), ("fir", O.BLUR_FIR)):
            sc, det = O.ssimu2_rgb8(src, dist, mode, detail=True)
            rec[name] = dict(score=sc, n_scales=det.n_scales, sums=O.detail_sums(det).tolist())
        cases.append(rec)
    with open(os.path.join(HERE, "oracle_scores.json"), "w") as f:
        json.dump(cases, f, indent=1)
    print("oracle_scores.json", [round(c["iir"]["score"], 3) for c in cases])


if __name__ == "__main__":
    golden_yuv()
    golden_roundtrip()
    golden_scores()
