"""ctypes view of the C++ host harness (include/oavif_host.h, oavif_b200/lib/liboavif_host.so).

The harness restates the callers of the scored path in C++ (the reference's host language, Zig,
cannot be compiled here): the search policy of /root/reference/src/tq.zig, the libavif glue of
src/io.zig and the corpus loop of scripts/measure.py.  This module only forwards to it.
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "lib", "liboavif_host.so"))


class Opts(C.Structure):
    _fields_ = [("quality_alpha", C.c_uint32), ("speed", C.c_uint32), ("max_threads", C.c_uint32),
                ("tile_rows_log2", C.c_uint32), ("tile_cols_log2", C.c_uint32), ("auto_tiling", C.c_uint32),
                ("score_tgt", C.c_double), ("tenbit", C.c_uint32), ("tune", C.c_char * 16),
                ("tolerance", C.c_double), ("max_pass", C.c_uint32), ("quality", C.c_int32),
                ("color_primaries", C.c_uint32), ("transfer_characteristics", C.c_uint32),
                ("matrix_coefficients", C.c_uint32)]


class Result(C.Structure):
    _fields_ = [("q", C.c_uint32), ("score", C.c_double), ("num_pass", C.c_uint32), ("early_exit", C.c_uint32),
                ("n_history", C.c_uint32), ("hist_q", C.c_uint32 * 16), ("hist_score", C.c_double * 16),
                ("device_passes", C.c_uint32), ("probes", C.c_uint32), ("wasted", C.c_uint32),
                ("size", C.c_uint64), ("reencoded", C.c_uint32), ("encode_ms", C.c_double),
                ("decode_ms", C.c_double), ("score_ms", C.c_double), ("total_ms", C.c_double),
                ("log", C.c_char * 512)]

    def history(self):
        return [(int(self.hist_q[i]), float(self.hist_score[i])) for i in range(self.n_history)]


PROBE_FN = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_uint32)
PROBE_BATCH_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_double))
SET_SOURCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32)
SCORE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                       C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double))

SCORE_PAIR_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint8), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                            C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int,
                            C.POINTER(C.c_double))


class CorpusStats(C.Structure):
    _fields_ = [("wall_s", C.c_double), ("scorer_device_ms", C.c_double), ("n_ok", C.c_uint32), ("n_err", C.c_uint32),
                ("workers", C.c_uint32), ("host_cpus", C.c_uint32), ("mean_encode_ms", C.c_double),
                ("mean_decode_ms", C.c_double), ("mean_score_ms", C.c_double), ("mean_passes", C.c_double),
                ("final_bytes_total", C.c_uint64), ("margin_hist", C.c_uint32 * 7)]


MARGIN_BUCKETS = ("<=1e-4", "<=1e-3", "<=0.01", "<=0.05", "<=0.1", "<=0.5", ">0.5")

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built — run `python -m oavif_b200.build`")
        L = C.CDLL(LIB_PATH)
        L.oavif_host_last_error.restype = C.c_char_p
        L.oavif_host_predict_q.restype = C.c_uint32
        L.oavif_host_predict_q.argtypes = [C.c_double]
        L.oavif_host_interpolate_q.restype = C.c_uint32
        L.oavif_host_interpolate_q.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_double),
                                               C.c_uint32, C.c_double]
        L.oavif_host_tq_search.argtypes = [C.c_double, C.c_double, C.c_uint32, PROBE_FN, C.c_void_p, C.POINTER(Result)]
        L.oavif_host_tq_search_batched.argtypes = [C.c_double, C.c_double, C.c_uint32, C.c_uint32, PROBE_BATCH_FN,
                                                   C.c_void_p, C.POINTER(Result)]
        L.oavif_host_search_image.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                              C.POINTER(Opts), C.c_uint32, C.c_int, C.c_int, SET_SOURCE_FN, SCORE_FN,
                                              C.c_void_p, C.POINTER(Result), C.c_void_p, C.c_size_t]
        L.oavif_host_corpus_synth.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                              C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                              C.POINTER(Opts), C.c_char_p, C.c_char_p, C.c_size_t,
                                              C.POINTER(CorpusStats)]
        L.oavif_host_tq_margins.argtypes = [C.c_double, C.c_double, C.c_uint32, C.POINTER(C.c_uint32),
                                            C.POINTER(C.c_double), C.c_uint32, C.c_double, C.POINTER(C.c_double),
                                            C.POINTER(C.c_double)]
        L.oavif_host_encode.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                        C.POINTER(Opts), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.oavif_host_decode_rgb8.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


def find_libavif() -> str | None:
    """The only libavif in this image is the one Pillow bundles."""
    env = os.environ.get("OAVIF_LIBAVIF")
    if env and os.path.exists(env):
        return env
    try:
        import PIL
    except Exception:
        return None
    hits = sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(PIL.__file__)), "pillow.libs", "libavif-*.so*")))
    return hits[0] if hits else None


def default_opts(**kw) -> Opts:
    o = Opts()
    load().oavif_host_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v.encode() if isinstance(v, str) else v)
    return o


def _err():
    return (load().oavif_host_last_error() or b"").decode()


def predict_q(tgt: float) -> int:
    return load().oavif_host_predict_q(tgt)


def interpolate_q(lo: int, hi: int, history, target: float) -> int:
    n = len(history)
    qs = (C.c_uint32 * max(n, 1))(*[h[0] for h in history])
    sc = (C.c_double * max(n, 1))(*[h[1] for h in history])
    return load().oavif_host_interpolate_q(lo, hi, qs, sc, n, target)


def tq_search(score_of_q, tgt=80.0, tol=2.0, max_pass=6) -> Result:
    r = Result()
    cb = PROBE_FN(lambda _u, q: float(score_of_q(int(q))))
    assert load().oavif_host_tq_search(tgt, tol, max_pass, cb, None, C.byref(r)) == 0
    return r


def tq_search_batched(score_of_q, width: int, tgt=80.0, tol=2.0, max_pass=6) -> Result:
    r = Result()

    def batch(_u, n, qs, out):
        for i in range(n):
            out[i] = float(score_of_q(int(qs[i])))

    cb = PROBE_BATCH_FN(batch)
    assert load().oavif_host_tq_search_batched(tgt, tol, max_pass, width, cb, None, C.byref(r)) == 0
    return r


def search_image(pixels: np.ndarray, opts: Opts | None = None, batch_width: int = 1, device: int = 0,
                 blur_mode: int = 0, scorer=None, libavif: str | None = None, want_bytes: bool = True):
    """main.zig:86-116 for one HxWx{3,4} uint8 image.  scorer=None: CUDA scorer on `device`;
    otherwise an object with set_source(rgb HxWx3) and score(y, u, v, depth, matrix, rgba) (tests)."""
    L = load()
    libavif = libavif or find_libavif()
    if not libavif:
        raise RuntimeError("libavif not found")
    px = np.ascontiguousarray(pixels, np.uint8)
    h, w, ch = px.shape
    opts = opts or default_opts()
    r = Result()
    buf = np.empty(w * h * 4 + 65536, np.uint8) if want_bytes else None
    keep = {}

    def ss(_u, p, ww, hh):
        try:
            scorer.set_source(np.ctypeslib.as_array(p, shape=(hh, ww, 3)).copy())
            return 0
        except Exception as e:  # pragma: no cover
            keep["err"] = e
            return -1

    def sc(_u, y, u, v, ys, us, vs, ww, hh, depth, matrix, rgba, out):
        try:
            dt, bps = (np.uint8, 1) if depth == 8 else (np.uint16, 2)

            def plane(ptr, stride):
                raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(hh, stride))
                return raw[:, : ww * bps].copy().view(dt).reshape(hh, ww)

            out[0] = float(scorer.score(plane(y, ys), plane(u, us), plane(v, vs), depth, matrix, bool(rgba)))
            return 0
        except Exception as e:  # pragma: no cover
            keep["err"] = e
            return -1

    ss_cb, sc_cb = SET_SOURCE_FN(ss), SCORE_FN(sc)
    rc = L.oavif_host_search_image(libavif.encode(), px.ctypes.data, w, h, ch, C.byref(opts), batch_width,
                                   -1 if scorer is not None else device, blur_mode, ss_cb, sc_cb, None, C.byref(r),
                                   buf.ctypes.data if want_bytes else None, buf.size if want_bytes else 0)
    if rc != 0:
        raise RuntimeError(f"oavif_host_search_image failed ({rc}): {_err()} {keep.get('err', '')}")
    return r, (bytes(buf[: r.size]) if want_bytes else None)


def tq_margins(history, tgt=80.0, tol=2.0, max_pass=6, limit=4.0):
    """[(flip_up, flip_down)] per pass: the smallest change of that pass's score that alters the search (tq.hpp)."""
    n = len(history)
    qs = (C.c_uint32 * n)(*[h[0] for h in history])
    sc = (C.c_double * n)(*[h[1] for h in history])
    up, dn = (C.c_double * n)(), (C.c_double * n)()
    assert load().oavif_host_tq_margins(tgt, tol, max_pass, qs, sc, n, limit, up, dn) == 0
    return list(zip(up, dn))


def corpus_synth(count: int, w: int, h: int, n_gpus: int = 1, first_gpu: int = 0, workers_per_gpu: int = 1,
                 batch_width: int = 1, blur_mode: int = 0, opts: Opts | None = None, csv_path: str | None = None,
                 libavif: str | None = None, pinned_staging: bool = True, score_pair=None):
    """The corpus sweep.  score_pair=None: the CUDA scorer.  Otherwise the CPU-scored arm: a callable
    (src_rgb HxWx3, y, u, v, depth, matrix, rgba) -> score that worker threads call concurrently (test / bench
    tooling passes the CPU oracle; ctypes releases the GIL inside its C calls)."""
    libavif = libavif or find_libavif()
    opts = opts or default_opts()
    summary = C.create_string_buffer(8192)
    st = CorpusStats()
    keep = {}
    cb = None
    if score_pair is not None:
        def pair(_u, src, y, u, v, ys, us, vs, ww, hh, depth, matrix, rgba, out):
            try:
                dt, bps = (np.uint8, 1) if depth == 8 else (np.uint16, 2)

                def plane(ptr, stride):
                    raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(hh, stride))
                    return raw[:, : ww * bps].view(dt).reshape(hh, ww) if stride == ww * bps else \
                        raw[:, : ww * bps].copy().view(dt).reshape(hh, ww)

                rgb = np.ctypeslib.as_array(src, shape=(hh, ww, 3))
                out[0] = float(score_pair(rgb, plane(y, ys), plane(u, us), plane(v, vs), depth, matrix, bool(rgba)))
                return 0
            except Exception as e:  # pragma: no cover
                keep["err"] = e
                return -1

        cb = SCORE_PAIR_FN(pair)
    rc = load().oavif_host_corpus_synth(libavif.encode(), count, w, h, first_gpu, n_gpus, workers_per_gpu, batch_width,
                                        blur_mode, int(pinned_staging), C.cast(cb, C.c_void_p) if cb else None, None,
                                        C.byref(opts), csv_path.encode() if csv_path else None, summary, 8192,
                                        C.byref(st))
    if rc != 0:
        raise RuntimeError(f"corpus failed: {_err()} {keep.get('err', '')}")
    return dict(wall_s=st.wall_s, ok=st.n_ok, errors=st.n_err, summary=summary.value.decode(), last_error=_err(),
                workers=st.workers, host_cpus=st.host_cpus, scorer_device_ms=st.scorer_device_ms,
                mean_encode_ms=st.mean_encode_ms, mean_decode_ms=st.mean_decode_ms, mean_score_ms=st.mean_score_ms,
                mean_passes=st.mean_passes, final_bytes_total=int(st.final_bytes_total),
                margin_hist=dict(zip(MARGIN_BUCKETS, [int(x) for x in st.margin_hist])))


def encode(pixels: np.ndarray, q: int, opts: Opts | None = None, libavif: str | None = None) -> bytes:
    px = np.ascontiguousarray(pixels, np.uint8)
    h, w, ch = px.shape
    buf = np.empty(w * h * 4 + 65536, np.uint8)
    size = C.c_size_t()
    opts = opts or default_opts()
    rc = load().oavif_host_encode((libavif or find_libavif()).encode(), px.ctypes.data, w, h, ch, q, C.byref(opts),
                                  buf.ctypes.data, buf.size, C.byref(size))
    if rc != 0:
        raise RuntimeError(f"encode failed: {_err()}")
    return bytes(buf[: size.value])


def decode_rgb8(avif: bytes, w: int, h: int, libavif: str | None = None) -> np.ndarray:
    out = np.empty((h, w, 3), np.uint8)
    rc = load().oavif_host_decode_rgb8((libavif or find_libavif()).encode(), avif, len(avif), out.ctypes.data, out.size)
    if rc != 0:
        raise RuntimeError(f"decode failed: {_err()}")
    return out
