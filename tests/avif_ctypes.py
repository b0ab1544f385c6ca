"""Minimal ctypes view of libavif (the shared object Pillow bundles) — TEST INFRASTRUCTURE.

Only what is needed to pin oracle/yuv2rgb_oracle.c against the function the reference calls at
/root/reference/src/io.zig:478 (avifImageYUVToRGB), and to make real encode/decode round trips
for fixtures.  Struct offsets are for libavif 1.4.x on x86-64 and are
re-checked by `self_check()` before use.
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np


def find_libavif() -> str | None:
    try:
        import PIL
    except Exception:
        return None
    base = os.path.join(os.path.dirname(os.path.dirname(PIL.__file__)), "pillow.libs")
    hits = sorted(glob.glob(os.path.join(base, "libavif-*.so*")))
    return hits[0] if hits else None


class RGBImage(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("depth", C.c_uint32), ("format", C.c_uint32),
                ("chromaUpsampling", C.c_uint32), ("chromaDownsampling", C.c_uint32), ("avoidLibYUV", C.c_int),
                ("ignoreAlpha", C.c_int), ("alphaPremultiplied", C.c_int), ("isFloat", C.c_int),
                ("maxThreads", C.c_int), ("_pad", C.c_int), ("pixels", C.c_void_p), ("rowBytes", C.c_uint32),
                ("_pad2", C.c_uint32)]


class RWData(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t)]


_L = None


def lib():
    global _L
    if _L is None:
        p = find_libavif()
        if not p:
            raise RuntimeError("libavif not found")
        L = C.CDLL(p)
        L.avifVersion.restype = C.c_char_p
        L.avifImageCreate.restype = C.c_void_p
        L.avifImageCreate.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.avifImageDestroy.argtypes = [C.c_void_p]
        L.avifImageAllocatePlanes.argtypes = [C.c_void_p, C.c_int]
        L.avifImageYUVToRGB.argtypes = [C.c_void_p, C.POINTER(RGBImage)]
        L.avifImageRGBToYUV.argtypes = [C.c_void_p, C.POINTER(RGBImage)]
        L.avifRGBImageSetDefaults.argtypes = [C.POINTER(RGBImage), C.c_void_p]
        L.avifRGBImageAllocatePixels.argtypes = [C.POINTER(RGBImage)]
        L.avifRGBImageFreePixels.argtypes = [C.POINTER(RGBImage)]
        L.avifEncoderCreate.restype = C.c_void_p
        L.avifEncoderDestroy.argtypes = [C.c_void_p]
        L.avifEncoderAddImage.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        L.avifEncoderFinish.argtypes = [C.c_void_p, C.POINTER(RWData)]
        L.avifEncoderSetCodecSpecificOption.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.avifRWDataFree.argtypes = [C.POINTER(RWData)]
        L.avifDecoderCreate.restype = C.c_void_p
        L.avifDecoderDestroy.argtypes = [C.c_void_p]
        L.avifDecoderSetIOMemory.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.avifDecoderParse.argtypes = [C.c_void_p]
        L.avifDecoderNextImage.argtypes = [C.c_void_p]
        _L = L
    return _L


# avifImage field offsets (libavif 1.4.x, x86-64)
OFF = dict(width=0, height=4, depth=8, yuvFormat=12, yuvRange=16, yuvPlanes=24, yuvRowBytes=48, alphaPlane=64,
           alphaRowBytes=72, colorPrimaries=104, transferCharacteristics=106, matrixCoefficients=108)
# avifEncoder field offsets
ENC = dict(maxThreads=4, speed=8, quality=32, qualityAlpha=36, tileRowsLog2=56, tileColsLog2=60, autoTiling=64)
# avifDecoder: the `image` pointer
DEC_IMAGE_CANDIDATES = (40, 48, 56, 64, 72, 80)


def _u32(p, off):
    return C.c_uint32.from_address(p + off).value


def _ptr(p, off):
    return C.c_void_p.from_address(p + off).value


def image_from_planes(y, u, v, depth, matrix=2, alpha=None):
    """avifImage* (YUV444, full range) holding copies of the given planes."""
    L = lib()
    h, w = y.shape
    img = L.avifImageCreate(w, h, depth, 1)  # AVIF_PIXEL_FORMAT_YUV444
    assert img
    assert L.avifImageAllocatePlanes(img, 1 | (2 if alpha is not None else 0)) == 0
    assert _u32(img, OFF["width"]) == w and _u32(img, OFF["depth"]) == depth
    C.c_uint16.from_address(img + OFF["matrixCoefficients"]).value = matrix
    bps = 1 if depth == 8 else 2
    for i, p in enumerate((y, u, v)):
        dst = _ptr(img, OFF["yuvPlanes"] + 8 * i)
        rb = _u32(img, OFF["yuvRowBytes"] + 4 * i)
        p = np.ascontiguousarray(p)
        for r in range(h):
            C.memmove(dst + r * rb, p.ctypes.data + r * w * bps, w * bps)
    if alpha is not None:
        dst = _ptr(img, OFF["alphaPlane"])
        rb = _u32(img, OFF["alphaRowBytes"])
        a = np.ascontiguousarray(alpha)
        for r in range(h):
            C.memmove(dst + r * rb, a.ctypes.data + r * w * bps, w * bps)
    return img


def image_to_rgb8(img) -> np.ndarray:
    """decodeAvifCommon's conversion (io.zig:466-478) + decodeAvifToRgb's repack (io.zig:654-663)."""
    L = lib()
    rgb = RGBImage()
    L.avifRGBImageSetDefaults(C.byref(rgb), img)
    rgb.depth = 8
    has_alpha = _ptr(img, OFF["alphaPlane"]) is not None
    rgb.format = 1 if has_alpha else 0
    assert L.avifRGBImageAllocatePixels(C.byref(rgb)) == 0
    rc = L.avifImageYUVToRGB(img, C.byref(rgb))
    assert rc == 0, rc
    ch = 4 if has_alpha else 3
    w, h = rgb.width, rgb.height
    raw = np.ctypeslib.as_array((C.c_uint8 * (rgb.rowBytes * h)).from_address(rgb.pixels)).reshape(h, rgb.rowBytes)
    out = raw[:, : w * ch].reshape(h, w, ch)[:, :, :3].copy()
    L.avifRGBImageFreePixels(C.byref(rgb))
    return out


def yuv444_to_rgb8(y, u, v, depth, matrix=2, with_alpha=False) -> np.ndarray:
    L = lib()
    alpha = None
    if with_alpha:
        alpha = np.full(y.shape, (1 << depth) - 1, y.dtype)
    img = image_from_planes(y, u, v, depth, matrix, alpha)
    try:
        return image_to_rgb8(img)
    finally:
        L.avifImageDestroy(img)
