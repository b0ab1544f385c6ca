"""ctypes loader for the CPU oracle (oracle/liboracle*.so).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs — never from the product package ``oavif_b200``.
Status of the scorer restatement: **parity unpinned** vs fssimu2 0.1.1 (see ssimu2_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_SCALES = 6
BLUR_IIR, BLUR_FIR, BLUR_FIR64 = 0, 1, 2
VARIANT_CONTIGUOUS_WEIGHTS, VARIANT_VERTICAL_ORDER, VARIANT_F32_MAPS, VARIANT_F32_TRANSFER = 1, 2, 4, 8


class Detail(C.Structure):
    _fields_ = [
        ("n_scales", C.c_int),
        ("w", C.c_int * MAX_SCALES),
        ("h", C.c_int * MAX_SCALES),
        ("sums", (C.c_double * 18) * MAX_SCALES),
        ("avg_ssim", (C.c_double * 6) * MAX_SCALES),
        ("avg_edgediff", (C.c_double * 12) * MAX_SCALES),
        ("score", C.c_double),
    ]


def build(force: bool = False) -> None:
    """Compile the oracle with its Makefile (gcc only; seconds)."""
    have = (os.path.exists(os.path.join(_HERE, "liboracle.so"))
            and os.path.exists(os.path.join(_HERE, "liboracle_fast.so")))
    try:    # make is incremental: a no-op when both libraries are newer than the sources
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    except (OSError, subprocess.CalledProcessError):
        if force or not have:
            raise


_libs: dict[str, C.CDLL] = {}


def _native_path() -> str | None:
    """liboracle_fast built with -march=native ON THIS MACHINE (the timed CPU baseline should use the host's own
    vector width; the shipped liboracle_fast.so is x86-64-v3 so that it runs anywhere).  Cached per CPU flag set
    under oracle/_native/ (git-ignored); None when it cannot be built."""
    import hashlib
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags"))
    except (OSError, StopIteration):
        return None
    d = os.path.join(_HERE, "_native")
    path = os.path.join(d, "liboracle_native_%s.so" % hashlib.sha1(flags.encode()).hexdigest()[:10])
    srcs = [os.path.join(_HERE, f) for f in ("ssimu2_oracle.c", "yuv2rgb_oracle.c")]
    try:
        if not os.path.exists(path) or any(os.path.getmtime(path) < os.path.getmtime(f) for f in srcs):
            os.makedirs(d, exist_ok=True)
            subprocess.check_call([os.environ.get("CC", "gcc"), "-std=c11", "-fPIC", "-fno-math-errno", "-O3", "-march=native",
                                   "-ffp-contract=off", "-shared", "-o", path + ".tmp", *srcs, "-lm"],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            os.replace(path + ".tmp", path)
        return path
    except (OSError, subprocess.CalledProcessError):
        return None


def lib(fast: bool = False) -> C.CDLL:
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    if fast and os.environ.get("OAVIF_ORACLE_NATIVE", "1") != "0":
        path = _native_path() or path
    L = C.CDLL(path)
    u8p, f32p, f64p, ip = (C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_double),
                           C.POINTER(C.c_int))
    L.oracle_srgb_lut.argtypes = [f32p]
    L.oracle_rgb8_to_linear.argtypes = [u8p, C.c_int, C.c_int, C.c_int, f32p]
    L.oracle_downsample2x.argtypes = [f32p, C.c_int, C.c_int, f32p]
    L.oracle_linear_to_xyb.argtypes = [f32p, C.c_int, f32p]
    L.oracle_rg_coeffs.argtypes = [C.c_double, f64p, f64p, ip]
    L.oracle_fir_taps.argtypes = [C.c_double, f64p, C.c_int]
    L.oracle_fir_taps.restype = C.c_int
    L.oracle_blur.argtypes = [f32p, C.c_int, C.c_int, C.c_int, f32p, f32p]
    L.oracle_final_score.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    L.oracle_final_score.restype = C.c_double
    L.oracle_weights.restype = f64p
    L.oracle_set_variant.argtypes = [C.c_int]
    L.oracle_set_libm_cbrt.argtypes = [C.c_int]
    L.oracle_ssimu2_rgb8.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     f64p, C.POINTER(Detail)]
    L.oracle_ssimu2_rgb8.restype = C.c_int
    L.oracle_xyb_at_scale.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, f32p, ip, ip]
    L.oracle_xyb_at_scale.restype = C.c_int
    L.oracle_yuv444_to_rgb8.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
    L.oracle_yuv444_to_rgb8.restype = C.c_int
    L.oracle_to_rgb8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
    L.oracle_to_rgb8.restype = C.c_int
    L.oracle_source_samples.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    L.oracle_source_samples.restype = C.c_int
    _libs[name] = L
    return L


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def set_variant(flags: int, fast: bool = False, libm_cbrt: bool = False) -> None:
    """Process-wide variant switches of one library (see ssimu2_oracle.h); 0 restores the default."""
    lib(fast).oracle_set_variant(flags)
    lib(fast).oracle_set_libm_cbrt(int(libm_cbrt))


def fast_flavour() -> str:
    """What `fast=True` loads on this machine: 'native' (-O3 -march=native, built here) or 'x86-64-v3'."""
    lib(True)
    return "native" if "_native" in (getattr(_libs.get("liboracle_fast.so"), "_name", "") or "") else "x86-64-v3"


def srgb_lut() -> np.ndarray:
    out = np.empty(256, np.float32)
    lib().oracle_srgb_lut(_f32(out))
    return out


def ssimu2_rgb8(ref: np.ndarray, dist: np.ndarray, blur: int = BLUR_IIR, fast: bool = False,
                detail: bool = False):
    """Score two HxWx3 uint8 arrays (the contract of tq.zig:37).  Returns score or (score, Detail)."""
    ref = np.ascontiguousarray(ref, np.uint8)
    dist = np.ascontiguousarray(dist, np.uint8)
    if ref.shape != dist.shape or ref.ndim != 3 or ref.shape[2] != 3:
        raise ValueError("expected two HxWx3 uint8 arrays of equal shape")
    h, w, _ = ref.shape
    score = C.c_double()
    d = Detail()
    rc = lib(fast).oracle_ssimu2_rgb8(_u8(ref), 3 * w, _u8(dist), 3 * w, w, h, blur,
                                      C.byref(score), C.byref(d))
    if rc != 0:
        raise RuntimeError(f"oracle_ssimu2_rgb8 failed: {rc}")
    return (score.value, d) if detail else score.value


def detail_sums(d: Detail) -> np.ndarray:
    """(6, 18) array of the raw pooled sums; rows >= n_scales are zero."""
    return np.array([[d.sums[s][i] for i in range(18)] for s in range(MAX_SCALES)], np.float64)


def xyb_at_scale(rgb: np.ndarray, scale: int) -> np.ndarray:
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    out = np.empty((3, h, w), np.float32)
    ws, hs = C.c_int(), C.c_int()
    rc = lib().oracle_xyb_at_scale(_u8(rgb), 3 * w, w, h, scale, _f32(out), C.byref(ws), C.byref(hs))
    if rc != 0:
        raise RuntimeError(f"oracle_xyb_at_scale failed: {rc}")
    return out.reshape(-1)[: 3 * ws.value * hs.value].reshape(3, hs.value, ws.value).copy()


def blur(plane: np.ndarray, mode: int = BLUR_IIR, rows_only: bool = False) -> np.ndarray:
    """Both passes of the blur, or (rows_only) the horizontal pass alone as it stands before the vertical one."""
    plane = np.ascontiguousarray(plane, np.float32)
    h, w = plane.shape
    tmp = np.empty_like(plane)
    out = np.empty_like(plane)
    lib().oracle_blur(_f32(plane), w, h, mode, _f32(tmp), _f32(out))
    return tmp if rows_only else out


def fir_taps(sigma: float = 1.5) -> np.ndarray:
    buf = np.zeros(32, np.float64)
    n = lib().oracle_fir_taps(sigma, buf.ctypes.data_as(C.POINTER(C.c_double)), 32)
    return buf[:n].copy()


def rg_coeffs(sigma: float = 1.5):
    n2 = np.zeros(3, np.float64)
    d1 = np.zeros(3, np.float64)
    r = C.c_int()
    lib().oracle_rg_coeffs(sigma, n2.ctypes.data_as(C.POINTER(C.c_double)),
                           d1.ctypes.data_as(C.POINTER(C.c_double)), C.byref(r))
    return n2, d1, r.value


def yuv444_to_rgb8(y: np.ndarray, u: np.ndarray, v: np.ndarray, depth: int, matrix: int = 2,
                   rgba_path: bool = False) -> np.ndarray:
    dt = np.uint8 if depth == 8 else np.uint16
    y, u, v = (np.ascontiguousarray(p, dt) for p in (y, u, v))
    h, w = y.shape
    out = np.empty((h, w, 3), np.uint8)
    bs = y.itemsize * w
    rc = lib().oracle_yuv444_to_rgb8(y.ctypes.data, u.ctypes.data, v.ctypes.data, bs, bs, bs, w, h,
                                     depth, matrix, int(rgba_path), _u8(out))
    if rc != 0:
        raise RuntimeError(f"oracle_yuv444_to_rgb8 failed: {rc}")
    return out


def to_rgb8(data: np.ndarray, channels: int, hbd: bool) -> np.ndarray:
    data = np.ascontiguousarray(data)
    h, w = data.shape[:2]
    out = np.empty((h, w, 3), np.uint8)
    rc = lib().oracle_to_rgb8(data.ctypes.data, w, h, channels, int(hbd), _u8(out))
    if rc != 0:
        raise RuntimeError(f"oracle_to_rgb8 failed: {rc}")
    return out


def source_samples(data: np.ndarray, out_depth: int) -> np.ndarray:
    """encodeAvifToBuffer's depth conversion of the source samples (io.zig:562-609); uint8 or uint16 input."""
    data = np.ascontiguousarray(data)
    out = np.empty(data.shape, np.uint16 if out_depth > 8 else np.uint8)
    rc = lib().oracle_source_samples(data.ctypes.data, data.size, int(data.dtype == np.uint16), out_depth, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_source_samples: no conversion from {data.dtype} to depth {out_depth}")
    return out


_contracted = None


def ssimu2_rgb8_contracted(ref: np.ndarray, dist: np.ndarray, blur: int = BLUR_IIR) -> float | None:
    """The same source compiled the way a C or LLVM tool chain contracts by default on an FMA machine
    (-O3 -march=native -ffp-contract=fast: every a*b + c the compiler sees becomes one fused operation) — one more
    reading of "the same algorithm" for scripts/variant_envelope.py.  None where it cannot be built or the CPU has no
    FMA.  Built under oracle/_native/ (git-ignored)."""
    global _contracted
    if _contracted is None:
        try:
            if " fma " not in next(l for l in open("/proc/cpuinfo") if l.startswith("flags")) + " ":
                return None
            d = os.path.join(_HERE, "_native")
            os.makedirs(d, exist_ok=True)
            path = os.path.join(d, "liboracle_contracted.so")
            srcs = [os.path.join(_HERE, f) for f in ("ssimu2_oracle.c", "yuv2rgb_oracle.c")]
            if not os.path.exists(path) or any(os.path.getmtime(path) < os.path.getmtime(f) for f in srcs):
                subprocess.check_call([os.environ.get("CC", "gcc"), "-std=c11", "-fPIC", "-fno-math-errno", "-O3", "-march=native",
                                       "-ffp-contract=fast", "-shared", "-o", path + ".tmp", *srcs, "-lm"],
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                os.replace(path + ".tmp", path)
            L = C.CDLL(path)
            L.oracle_ssimu2_rgb8.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.POINTER(C.c_double), C.POINTER(Detail)]
            L.oracle_ssimu2_rgb8.restype = C.c_int
            _contracted = L
        except (OSError, StopIteration, subprocess.CalledProcessError):
            return None
    ref = np.ascontiguousarray(ref, np.uint8)
    dist = np.ascontiguousarray(dist, np.uint8)
    h, w, _ = ref.shape
    score = C.c_double()
    d = Detail()
    if _contracted.oracle_ssimu2_rgb8(_u8(ref), 3 * w, _u8(dist), 3 * w, w, h, blur, C.byref(score), C.byref(d)) != 0:
        raise RuntimeError("oracle_ssimu2_rgb8 (contracted build) failed")
    return score.value
