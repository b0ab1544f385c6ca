"""ctypes binding of the C ABI in include/oavif_ssimu2.h (liboavif_ssimu2.so, CUDA sm_100a).

This is the host-side mirror of the one interface the reference has for the scored path:
``fssimu2.computeSsimu2(allocator, ref, dist, w, h, channels, null) !f64`` called at
/root/reference/src/tq.zig:37 — see :func:`compute_ssimu2` (same argument meaning, raises where
the Zig call returns an error).  :class:`Scorer` is the stateful form the search loop uses:
source uploaded once per image (main.zig:86), one or several candidates scored per pass
(tq.zig:150).

There is no CPU path here.  If the shared library is missing or CUDA is unavailable every call
raises :class:`Ssimu2Error`; nothing in this module imports ``oracle``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "lib", "liboavif_ssimu2.so"))

MAX_SCALES = 6
BLUR_RECURSIVE, BLUR_FIR = 0, 1
WEIGHTS_SIX_SLOTS, WEIGHTS_CONTIGUOUS = 0, 1
TILES_TMA, TILES_CP_ASYNC, TILES_FUSED, TILES_TMA_DECOUPLED = 0, 1, 2, 3
SOURCE_ROWS_AT_SET_SOURCE, SOURCE_ROWS_WITH_FIRST_SCORE = 0, 1
OPT_BLUR, OPT_WEIGHTS, OPT_TILE_PATH, OPT_SOURCE_ROWS, OPT_TRANSFER, OPT_VERTICAL_ORDER = 1, 2, 3, 4, 5, 6
TRANSFER_F64, TRANSFER_F32 = 0, 1
VERTICAL_AS_HORIZONTAL, VERTICAL_FUSED_OUTER = 0, 1

E_ARG, E_CUDA, E_NOMEM, E_STATE, E_UNSUPPORTED = -1, -2, -3, -4, -5
_ENAMES = {E_ARG: "InvalidArgument", E_CUDA: "CudaError", E_NOMEM: "OutOfMemory",
           E_STATE: "InvalidState", E_UNSUPPORTED: "Unsupported"}

# every symbol include/oavif_ssimu2.h declares (tests check the .so exports all of them)
SYMBOLS = (
    "oavif_ssimu2_abi_version", "oavif_ssimu2_ctx_create", "oavif_ssimu2_ctx_destroy",
    "oavif_ssimu2_set_option", "oavif_ssimu2_set_stream", "oavif_ssimu2_last_error",
    "oavif_ssimu2_pinned_alloc", "oavif_ssimu2_pinned_free", "oavif_ssimu2_set_source_rgb8",
    "oavif_ssimu2_set_source_pixels", "oavif_ssimu2_score_pixels",
    "oavif_ssimu2_score_rgb8", "oavif_ssimu2_score_yuv444", "oavif_ssimu2_score_batch_rgb8",
    "oavif_ssimu2_score_batch_yuv444", "oavif_ssimu2_set_source_rgb8_dev",
    "oavif_ssimu2_score_batch_rgb8_dev", "oavif_ssimu2_score_batch_yuv444_dev",
    "oavif_ssimu2_compute_rgb8", "oavif_ssimu2_yuv444_to_rgb8", "oavif_ssimu2_get_detail",
    "oavif_ssimu2_get_timing", "oavif_ssimu2_debug_get_xyb", "oavif_ssimu2_debug_get_rows", "oavif_ssimu2_debug_blur",
    "oavif_ssimu2_debug_time_rows", "oavif_ssimu2_debug_check_guards", "oavif_ssimu2_debug_get_cols",
    "oavif_ssimu2_set_default_device", "oavif_ssimu2_release_cached",
    "oavif_ssimu2_get_option", "oavif_ssimu2_device_pci_bus_id", "oavif_ssimu2_debug_wave_trace", "oavif_ssimu2_submit_rgb8", "oavif_ssimu2_submit_yuv444",
    "oavif_ssimu2_submit_rgb8_dev", "oavif_ssimu2_submit_yuv444_dev", "oavif_ssimu2_wait", "oavif_ssimu2_in_flight",
    "oavif_ssimu2_source_samples",
)


class Ssimu2Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{_ENAMES.get(code, code)}: {msg}")
        self.code = code


class Detail(C.Structure):
    _fields_ = [("n_scales", C.c_int32), ("w", C.c_int32 * MAX_SCALES), ("h", C.c_int32 * MAX_SCALES),
                ("sums", (C.c_double * 18) * MAX_SCALES), ("score", C.c_double)]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("pyramid_ms", C.c_float), ("blur_ms", C.c_float),
                ("blur_a_ms", C.c_float), ("blur_b_ms", C.c_float), ("finalize_ms", C.c_float), ("total_ms", C.c_float), ("launches", C.c_uint32)]


_lib = None


def load() -> C.CDLL:
    """dlopen the CUDA library and declare prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Ssimu2Error(E_CUDA, f"{LIB_PATH} not built — run `python -m oavif_b200.build` "
                                  "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u8p, dp, szt, u32 = C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_size_t, C.c_uint32
    L.oavif_ssimu2_abi_version.restype = C.c_int
    L.oavif_ssimu2_ctx_create.argtypes = [C.c_int, u32, u32, u32, C.POINTER(vp)]
    L.oavif_ssimu2_ctx_destroy.argtypes = [vp]
    L.oavif_ssimu2_ctx_destroy.restype = None
    L.oavif_ssimu2_set_option.argtypes = [vp, C.c_int, C.c_int]
    L.oavif_ssimu2_get_option.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
    L.oavif_ssimu2_set_stream.argtypes = [vp, vp]
    L.oavif_ssimu2_last_error.argtypes = [vp]
    L.oavif_ssimu2_last_error.restype = C.c_char_p
    L.oavif_ssimu2_pinned_alloc.argtypes = [szt]
    L.oavif_ssimu2_pinned_alloc.restype = vp
    L.oavif_ssimu2_pinned_free.argtypes = [vp]
    L.oavif_ssimu2_pinned_free.restype = None
    L.oavif_ssimu2_set_source_rgb8.argtypes = [vp, u8p, u32, u32, szt]
    L.oavif_ssimu2_set_source_rgb8_dev.argtypes = [vp, u8p, u32, u32, szt]
    L.oavif_ssimu2_score_rgb8.argtypes = [vp, u8p, szt, dp]
    L.oavif_ssimu2_set_source_pixels.argtypes = [vp, vp, u32, u32, szt, C.c_int, C.c_int]
    L.oavif_ssimu2_score_pixels.argtypes = [vp, vp, szt, C.c_int, C.c_int, dp]
    L.oavif_ssimu2_score_yuv444.argtypes = [vp, vp, vp, vp, szt, szt, szt, C.c_int, C.c_int, C.c_int, dp]
    L.oavif_ssimu2_score_batch_rgb8.argtypes = [vp, u32, C.POINTER(vp), szt, dp]
    L.oavif_ssimu2_score_batch_rgb8_dev.argtypes = [vp, u32, C.POINTER(vp), szt, dp]
    for f in (L.oavif_ssimu2_score_batch_yuv444, L.oavif_ssimu2_score_batch_yuv444_dev):
        f.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), szt, szt, szt, C.c_int,
                      C.c_int, C.c_int, dp]
    L.oavif_ssimu2_submit_rgb8.argtypes = [vp, u32, C.POINTER(vp), szt]
    L.oavif_ssimu2_submit_yuv444.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), szt, szt, szt,
                                             C.c_int, C.c_int, C.c_int]
    L.oavif_ssimu2_submit_rgb8_dev.argtypes = L.oavif_ssimu2_submit_rgb8.argtypes
    L.oavif_ssimu2_submit_yuv444_dev.argtypes = L.oavif_ssimu2_submit_yuv444.argtypes
    L.oavif_ssimu2_wait.argtypes = [vp, dp]
    L.oavif_ssimu2_in_flight.argtypes = [vp]
    L.oavif_ssimu2_compute_rgb8.argtypes = [u8p, u8p, u32, u32, u32, dp]
    L.oavif_ssimu2_yuv444_to_rgb8.argtypes = [vp, vp, vp, vp, szt, szt, szt, u32, u32, C.c_int, C.c_int,
                                              C.c_int, u8p]
    L.oavif_ssimu2_source_samples.argtypes = [vp, C.c_int, vp, szt]
    L.oavif_ssimu2_get_detail.argtypes = [vp, u32, C.POINTER(Detail)]
    L.oavif_ssimu2_get_timing.argtypes = [vp, C.POINTER(Timing)]
    L.oavif_ssimu2_debug_get_xyb.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.POINTER(u32), C.POINTER(u32)]
    L.oavif_ssimu2_debug_get_rows.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.POINTER(u32), C.POINTER(u32)]
    L.oavif_ssimu2_debug_get_cols.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.POINTER(u32), C.POINTER(u32)]
    L.oavif_ssimu2_set_default_device.argtypes = [C.c_int]
    L.oavif_ssimu2_release_cached.restype = None
    L.oavif_ssimu2_debug_wave_trace.argtypes = [vp, C.c_int, vp, u32, C.POINTER(u32)]
    L.oavif_ssimu2_debug_blur.argtypes = [vp, vp, u32, u32, vp]
    L.oavif_ssimu2_debug_check_guards.argtypes = [vp]
    L.oavif_ssimu2_debug_time_rows.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
    _lib = L
    return L


def _check(rc: int, ctx=None):
    if rc != 0:
        msg = load().oavif_ssimu2_last_error(ctx)
        raise Ssimu2Error(rc, msg.decode() if msg else "")


def _rgb8(a) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise Ssimu2Error(E_ARG, "expected an HxWx3 uint8 array")
    if a.strides[2] != 1 or a.strides[1] != 3:
        a = np.ascontiguousarray(a)
    return a


def compute_ssimu2(ref, dist, w: int | None = None, h: int | None = None, channels: int = 3) -> float:
    """``fssimu2.computeSsimu2(ref, dist, w, h, channels)`` (tq.zig:37) on the GPU.

    ref/dist: HxWx3 uint8 arrays, or flat byte buffers of w*h*channels bytes with w, h given."""
    L = load()
    ref = np.asarray(ref)
    dist = np.asarray(dist)
    if ref.ndim == 3:
        h, w = ref.shape[:2]
        channels = ref.shape[2]
    if w is None or h is None:
        raise Ssimu2Error(E_ARG, "w and h are required for flat buffers")
    ref = np.ascontiguousarray(ref, np.uint8).reshape(-1)
    dist = np.ascontiguousarray(dist, np.uint8).reshape(-1)
    if ref.size != w * h * channels or dist.size != ref.size:
        raise Ssimu2Error(E_ARG, "buffer size does not match w*h*channels")
    out = C.c_double()
    _check(L.oavif_ssimu2_compute_rgb8(ref.ctypes.data, dist.ctypes.data, w, h, channels, C.byref(out)))
    return out.value


class Scorer:
    """One scorer context: a CUDA device, a capacity, a cached source (single-owner, not thread-safe)."""

    def __init__(self, max_w: int, max_h: int, max_batch: int = 1, device: int = 0, blur: int = BLUR_RECURSIVE):
        self._L = load()
        self._ctx = C.c_void_p()
        _check(self._L.oavif_ssimu2_ctx_create(device, max_w, max_h, max_batch, C.byref(self._ctx)))
        self.max_batch = max_batch
        self.w = self.h = 0
        if blur != BLUR_RECURSIVE:
            self.set_blur(blur)

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._L.oavif_ssimu2_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_blur(self, mode: int):
        _check(self._L.oavif_ssimu2_set_option(self._ctx, OPT_BLUR, mode), self._ctx)

    def set_tile_path(self, path: int):
        _check(self._L.oavif_ssimu2_set_option(self._ctx, OPT_TILE_PATH, path), self._ctx)

    def set_option(self, option: int, value: int):
        _check(self._L.oavif_ssimu2_set_option(self._ctx, option, value), self._ctx)

    def get_option(self, option: int) -> int:
        v = C.c_int()
        _check(self._L.oavif_ssimu2_get_option(self._ctx, option, C.byref(v)), self._ctx)
        return v.value

    def set_weights(self, layout: int):
        _check(self._L.oavif_ssimu2_set_option(self._ctx, OPT_WEIGHTS, layout), self._ctx)

    def set_stream(self, cuda_stream: int | None):
        _check(self._L.oavif_ssimu2_set_stream(self._ctx, C.c_void_p(cuda_stream or 0)), self._ctx)

    # ---- source ------------------------------------------------------------------------------
    def set_source(self, rgb):
        a = _rgb8(rgb)
        self._src_keep = a
        self._src_shape = a.shape
        self.h, self.w = a.shape[:2]
        _check(self._L.oavif_ssimu2_set_source_rgb8(self._ctx, a.ctypes.data, self.w, self.h, a.strides[0]),
               self._ctx)

    @staticmethod
    def _pixels(a):
        a = np.ascontiguousarray(a)
        if a.ndim == 2:
            a = a[..., None]
        if a.dtype not in (np.uint8, np.uint16) or a.ndim != 3:
            raise Ssimu2Error(E_ARG, "expected HxW[xC] uint8/uint16 pixels")
        return a

    def set_source_pixels(self, pixels):
        """Loader-native layouts (io.zig's Image): HxWxC, C in 1..4, uint8 or uint16 -> Image.toRGB8 on the GPU."""
        a = self._pixels(pixels)
        self._src_keep = a
        self._src_shape = a.shape
        self.h, self.w = a.shape[:2]
        _check(self._L.oavif_ssimu2_set_source_pixels(self._ctx, a.ctypes.data, self.w, self.h, a.strides[0],
                                                      a.shape[2], 8 * a.itemsize), self._ctx)

    def score_pixels(self, pixels) -> float:
        a = self._pixels(pixels)
        out = C.c_double()
        _check(self._L.oavif_ssimu2_score_pixels(self._ctx, a.ctypes.data, a.strides[0], a.shape[2], 8 * a.itemsize,
                                                 C.byref(out)), self._ctx)
        return out.value

    def set_source_dev(self, dptr: int, w: int, h: int, stride: int):
        self.w, self.h = w, h
        self._src_shape = (h, w, 3)
        _check(self._L.oavif_ssimu2_set_source_rgb8_dev(self._ctx, C.c_void_p(dptr), w, h, stride), self._ctx)

    # ---- candidates ----------------------------------------------------------------------------
    def score_rgb8(self, dist) -> float:
        a = _rgb8(dist)
        if a.shape[:2] != (self.h, self.w):
            raise Ssimu2Error(E_ARG, "candidate size differs from the source")
        out = C.c_double()
        _check(self._L.oavif_ssimu2_score_rgb8(self._ctx, a.ctypes.data, a.strides[0], C.byref(out)), self._ctx)
        return out.value

    def score_batch_rgb8(self, dists: Sequence[np.ndarray]) -> list[float]:
        arrs = [_rgb8(d) for d in dists]
        n = len(arrs)
        if n == 0:
            raise Ssimu2Error(E_ARG, "empty batch")
        strides = {a.strides[0] for a in arrs}
        if len(strides) != 1 or any(a.shape[:2] != (self.h, self.w) for a in arrs):
            raise Ssimu2Error(E_ARG, "candidates must share the source size and one stride")
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        out = (C.c_double * n)()
        _check(self._L.oavif_ssimu2_score_batch_rgb8(self._ctx, n, ptrs, strides.pop(), out), self._ctx)
        return list(out)

    @staticmethod
    def _planes(y, u, v, depth):
        dt = np.uint8 if depth == 8 else np.uint16
        ps = []
        for p in (y, u, v):
            p = np.asarray(p)
            if p.dtype != dt or p.ndim != 2:
                raise Ssimu2Error(E_ARG, f"planes must be 2-D {np.dtype(dt).name} for depth {depth}")
            if p.strides[1] != p.itemsize:
                p = np.ascontiguousarray(p)
            ps.append(p)
        return ps

    def score_yuv444(self, y, u, v, depth: int, matrix: int = 2, rgba_path: bool = False) -> float:
        if depth not in (8, 10):
            raise Ssimu2Error(E_UNSUPPORTED, f"depth {depth}")
        y, u, v = self._planes(y, u, v, depth)
        out = C.c_double()
        _check(self._L.oavif_ssimu2_score_yuv444(self._ctx, y.ctypes.data, u.ctypes.data, v.ctypes.data,
                                                 y.strides[0], u.strides[0], v.strides[0], depth, matrix,
                                                 int(rgba_path), C.byref(out)), self._ctx)
        return out.value

    def score_batch_yuv444(self, cands: Sequence[tuple], depth: int, matrix: int = 2,
                           rgba_path: bool = False) -> list[float]:
        n = len(cands)
        if n == 0:
            raise Ssimu2Error(E_ARG, "empty batch")
        if depth not in (8, 10):
            raise Ssimu2Error(E_UNSUPPORTED, f"depth {depth}")
        ps = [self._planes(*c, depth) for c in cands]
        ys = (C.c_void_p * n)(*[p[0].ctypes.data for p in ps])
        us = (C.c_void_p * n)(*[p[1].ctypes.data for p in ps])
        vs = (C.c_void_p * n)(*[p[2].ctypes.data for p in ps])
        out = (C.c_double * n)()
        _check(self._L.oavif_ssimu2_score_batch_yuv444(self._ctx, n, ys, us, vs, ps[0][0].strides[0],
                                                       ps[0][1].strides[0], ps[0][2].strides[0], depth, matrix,
                                                       int(rgba_path), out), self._ctx)
        return list(out)

    # ---- pipelined form: up to two submissions in flight, uploads run under the previous one's kernels ----------
    def submit_rgb8(self, dists: Sequence[np.ndarray]):
        arrs = [_rgb8(d) for d in dists]
        n = len(arrs)
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        _check(self._L.oavif_ssimu2_submit_rgb8(self._ctx, n, ptrs, arrs[0].strides[0]), self._ctx)
        self._pending = getattr(self, "_pending", []) + [(n, arrs)]     # keeps the arrays alive until wait()

    def submit_yuv444(self, cands: Sequence[tuple], depth: int, matrix: int = 2, rgba_path: bool = False):
        n = len(cands)
        ps = [self._planes(*c, depth) for c in cands]
        ys = (C.c_void_p * n)(*[p[0].ctypes.data for p in ps])
        us = (C.c_void_p * n)(*[p[1].ctypes.data for p in ps])
        vs = (C.c_void_p * n)(*[p[2].ctypes.data for p in ps])
        _check(self._L.oavif_ssimu2_submit_yuv444(self._ctx, n, ys, us, vs, ps[0][0].strides[0], ps[0][1].strides[0],
                                                  ps[0][2].strides[0], depth, matrix, int(rgba_path)), self._ctx)
        self._pending = getattr(self, "_pending", []) + [(n, ps)]

    def submit_dev(self, kind: str, ptrs: Sequence[Sequence[int]], strides: Sequence[int], depth: int = 8,
                   matrix: int = 2, rgba_path: bool = False):
        """Device-resident candidates, pipelined form (see score_batch_dev for the pointer layout)."""
        n = len(ptrs)
        if kind == "rgb8":
            arr = (C.c_void_p * n)(*[p[0] for p in ptrs])
            _check(self._L.oavif_ssimu2_submit_rgb8_dev(self._ctx, n, arr, strides[0]), self._ctx)
        else:
            ys = (C.c_void_p * n)(*[p[0] for p in ptrs])
            us = (C.c_void_p * n)(*[p[1] for p in ptrs])
            vs = (C.c_void_p * n)(*[p[2] for p in ptrs])
            _check(self._L.oavif_ssimu2_submit_yuv444_dev(self._ctx, n, ys, us, vs, strides[0], strides[1], strides[2],
                                                          depth, matrix, int(rgba_path)), self._ctx)
        self._pending = getattr(self, "_pending", []) + [(n, None)]

    def wait(self) -> list[float]:
        pend = getattr(self, "_pending", [])
        n = pend[0][0] if pend else self.max_batch
        out = (C.c_double * max(n, 1))()
        _check(self._L.oavif_ssimu2_wait(self._ctx, out), self._ctx)
        self._pending = pend[1:]
        return list(out)[:n]

    def in_flight(self) -> int:
        return self._L.oavif_ssimu2_in_flight(self._ctx)

    def score_batch_dev(self, kind: str, ptrs: Sequence[Sequence[int]], strides: Sequence[int], depth: int = 8,
                        matrix: int = 2, rgba_path: bool = False) -> list[float]:
        """Device-resident candidates.  kind 'rgb8': ptrs = [[d_rgb], ...]; 'yuv444': [[d_y, d_u, d_v], ...]."""
        n = len(ptrs)
        out = (C.c_double * n)()
        if kind == "rgb8":
            arr = (C.c_void_p * n)(*[p[0] for p in ptrs])
            _check(self._L.oavif_ssimu2_score_batch_rgb8_dev(self._ctx, n, arr, strides[0], out), self._ctx)
        else:
            ys = (C.c_void_p * n)(*[p[0] for p in ptrs])
            us = (C.c_void_p * n)(*[p[1] for p in ptrs])
            vs = (C.c_void_p * n)(*[p[2] for p in ptrs])
            _check(self._L.oavif_ssimu2_score_batch_yuv444_dev(self._ctx, n, ys, us, vs, strides[0], strides[1],
                                                               strides[2], depth, matrix, int(rgba_path), out),
                   self._ctx)
        return list(out)

    # ---- decode-side helper --------------------------------------------------------------------------
    def yuv444_to_rgb8(self, y, u, v, depth: int, matrix: int = 2, rgba_path: bool = False) -> np.ndarray:
        if depth not in (8, 10):
            raise Ssimu2Error(E_UNSUPPORTED, f"depth {depth}")
        y, u, v = self._planes(y, u, v, depth)
        h, w = y.shape
        out = np.empty((h, w, 3), np.uint8)
        _check(self._L.oavif_ssimu2_yuv444_to_rgb8(self._ctx, y.ctypes.data, u.ctypes.data, v.ctypes.data,
                                                   y.strides[0], u.strides[0], v.strides[0], w, h, depth, matrix,
                                                   int(rgba_path), out.ctypes.data), self._ctx)
        return out

    def source_samples(self, out_depth: int) -> np.ndarray:
        """encodeAvifToBuffer's per-pass depth conversion of the source (io.zig:562-609), once, on the GPU: HxWxC samples
        at `out_depth` (uint16 for 10, uint8 for 8) from the pixels the last set_source / set_source_pixels staged."""
        out = np.empty(self._src_shape, np.uint16 if out_depth > 8 else np.uint8)
        _check(self._L.oavif_ssimu2_source_samples(self._ctx, out_depth, out.ctypes.data, out.nbytes), self._ctx)
        return out

    # ---- introspection ----------------------------------------------------------------------------------
    def detail(self, candidate: int = 0) -> Detail:
        d = Detail()
        _check(self._L.oavif_ssimu2_get_detail(self._ctx, candidate, C.byref(d)), self._ctx)
        return d

    def sums(self, candidate: int = 0) -> np.ndarray:
        d = self.detail(candidate)
        return np.array([[d.sums[s][i] for i in range(18)] for s in range(MAX_SCALES)], np.float64)

    def timing(self) -> Timing:
        t = Timing()
        _check(self._L.oavif_ssimu2_get_timing(self._ctx, C.byref(t)), self._ctx)
        return t

    def xyb(self, which: int, scale: int, channel: int) -> np.ndarray:
        buf = np.empty(self.w * self.h, np.float32)
        w, h = C.c_uint32(), C.c_uint32()
        _check(self._L.oavif_ssimu2_debug_get_xyb(self._ctx, which, scale, channel, buf.ctypes.data, C.byref(w),
                                                  C.byref(h)), self._ctx)
        return buf[: w.value * h.value].reshape(h.value, w.value).copy()

    def rows(self, candidate: int, quantity: int, scale: int, channel: int) -> np.ndarray:
        """Row-filtered plane of the last RECURSIVE score call; quantity 0..4 = a, b, a*a, b*b, a*b."""
        buf = np.empty(self.w * self.h, np.float32)
        w, h = C.c_uint32(), C.c_uint32()
        _check(self._L.oavif_ssimu2_debug_get_rows(self._ctx, candidate, quantity, scale, channel, buf.ctypes.data,
                                                   C.byref(w), C.byref(h)), self._ctx)
        return buf[: w.value * h.value].reshape(h.value, w.value).copy()

    def cols(self, candidate: int, scale: int, channel: int) -> np.ndarray:
        """(5, h, w): mu1, mu2, sigma11, sigma22, sigma12 as the product columns kernel hands them to the maps."""
        buf = np.empty(5 * self.w * self.h, np.float32)
        w, h = C.c_uint32(), C.c_uint32()
        _check(self._L.oavif_ssimu2_debug_get_cols(self._ctx, candidate, scale, channel, buf.ctypes.data,
                                                   C.byref(w), C.byref(h)), self._ctx)
        return buf[: 5 * w.value * h.value].reshape(5, h.value, w.value).copy()

    def wave_trace(self, mode: int = 2) -> np.ndarray:
        """(n_units, 5) uint64: unit, start, prologue end, middle phase, end (ns) of every CTA of one FUSED launch."""
        buf = np.zeros((8192, 5), np.uint64)
        n = C.c_uint32()
        _check(self._L.oavif_ssimu2_debug_wave_trace(self._ctx, mode, buf.ctypes.data, 8192, C.byref(n)), self._ctx)
        return buf[: min(n.value, 8192)].copy()

    def check_guards(self):
        """Raises if any kernel wrote past one of the context's device buffers."""
        _check(self._L.oavif_ssimu2_debug_check_guards(self._ctx), self._ctx)

    def time_rows(self, variant: int = 0, iters: int = 10) -> float:
        ms = C.c_float()
        _check(self._L.oavif_ssimu2_debug_time_rows(self._ctx, variant, iters, C.byref(ms)), self._ctx)
        return ms.value

    def blur(self, plane: np.ndarray) -> np.ndarray:
        p = np.ascontiguousarray(plane, np.float32)
        out = np.empty_like(p)
        _check(self._L.oavif_ssimu2_debug_blur(self._ctx, p.ctypes.data, p.shape[1], p.shape[0], out.ctypes.data),
               self._ctx)
        return out
