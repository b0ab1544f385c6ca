"""Procedural test images: gradients, edges, textured noise, mixtures.

No files, no network: every pixel is a pure function of (seed, kind, x, y) through a
counter-based splitmix64 hash, so any host language can regenerate the same bytes.
Used by bench.py, the tests and the corpus driver (config 5: kind = seed mod 4).
"""
from __future__ import annotations

import numpy as np

KINDS = ("gradient", "edges", "noise", "mixture")

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser on uint64 counters."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _uniform(seed: int, stream: int, h: int, w: int) -> np.ndarray:
    """float64 uniforms in [0,1) per pixel, independent per (seed, stream)."""
    idx = np.arange(h * w, dtype=np.uint64).reshape(h, w)
    key = np.uint64(((seed & 0xFFFFFFFF) << 32) | ((stream & 0xFFFF) << 16))
    with np.errstate(over="ignore"):
        bits = splitmix64(idx * np.uint64(0x100000001B3) + key)
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def _params(seed: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        bits = splitmix64(np.arange(n, dtype=np.uint64) + np.uint64(0xABCD0000) * np.uint64(seed + 1))
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def _box3(a: np.ndarray) -> np.ndarray:
    p = np.pad(a, 1, mode="edge")
    return (p[:-2, :-2] + p[:-2, 1:-1] + p[:-2, 2:] + p[1:-1, :-2] + p[1:-1, 1:-1] + p[1:-1, 2:]
            + p[2:, :-2] + p[2:, 1:-1] + p[2:, 2:]) / 9.0


def _gradient(w: int, h: int, seed: int) -> np.ndarray:
    p = _params(seed, 12).reshape(3, 4) * 255.0
    u = np.linspace(0.0, 1.0, w)[None, :]
    v = np.linspace(0.0, 1.0, h)[:, None]
    out = np.empty((h, w, 3), np.float64)
    for c in range(3):
        c00, c01, c10, c11 = p[c]
        out[..., c] = (c00 * (1 - u) * (1 - v) + c01 * u * (1 - v) + c10 * (1 - u) * v + c11 * u * v)
    return out


def _edges(w: int, h: int, seed: int) -> np.ndarray:
    p = _params(seed + 7919, 16)
    per = [int(8 * 2 ** int(p[i] * 4.999)) for i in range(3)]  # 8..128
    x = np.arange(w)[None, :]
    y = np.arange(h)[:, None]
    out = np.empty((h, w, 3), np.float64)
    checker = ((x // per[0]) + (y // per[0])) & 1
    bars = (x // per[1]) & 1
    cx, cy = w * (0.3 + 0.4 * p[3]), h * (0.3 + 0.4 * p[4])
    rings = (np.sqrt((x - cx) ** 2 + (y - cy) ** 2) // per[2]).astype(np.int64) & 1
    lo = 16 + 64 * p[5:8]
    hi = 255 - 64 * p[8:11]
    out[..., 0] = np.where(checker, hi[0], lo[0])
    out[..., 1] = np.where(bars, hi[1], lo[1]) * np.ones((h, 1))
    out[..., 2] = np.where(rings, hi[2], lo[2])
    return out


def _noise(w: int, h: int, seed: int) -> np.ndarray:
    base = _gradient(w, h, seed + 104729) * 0.6 + 50.0
    out = np.empty((h, w, 3), np.float64)
    for c in range(3):
        g = sum(_uniform(seed, 4 * c + k, h, w) for k in range(4)) - 2.0  # ~N(0, 1/3)
        g = _box3(g) * (12.0 * 3.0 * np.sqrt(3.0))                       # band-limit, sigma ~ 12 LSB
        out[..., c] = base[..., c] + g
    return out


def synth(w: int, h: int, kind: str | int = "mixture", seed: int = 0) -> np.ndarray:
    """HxWx3 uint8 RGB image."""
    if isinstance(kind, int):
        kind = KINDS[kind % 4]
    if kind == "gradient":
        img = _gradient(w, h, seed)
    elif kind == "edges":
        img = _edges(w, h, seed)
    elif kind == "noise":
        img = _noise(w, h, seed)
    elif kind == "mixture":
        g, e, n = _gradient(w, h, seed), _edges(w, h, seed), _noise(w, h, seed)
        x = np.arange(w)[None, :, None] / max(w - 1, 1)
        y = np.arange(h)[:, None, None] / max(h - 1, 1)
        img = np.where(x + y < 0.7, g, np.where(x - y > 0.1, e, n))
        img = 0.85 * img + 0.15 * n
    else:
        raise ValueError(f"unknown kind {kind!r}")
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synth_rgba(w: int, h: int, kind: str | int = "mixture", seed: int = 0) -> np.ndarray:
    """HxWx4 uint8: synth() plus a smooth radial alpha ramp (config 4)."""
    rgb = synth(w, h, kind, seed)
    x = (np.arange(w)[None, :] - w / 2) / (w / 2)
    y = (np.arange(h)[:, None] - h / 2) / (h / 2)
    a = np.clip(255.0 * (1.2 - np.sqrt(x * x + y * y)), 0, 255)
    return np.dstack([rgb, np.rint(a).astype(np.uint8)])


def distort(rgb: np.ndarray, strength: float, seed: int = 1) -> np.ndarray:
    """Codec-like degradation without a codec: blur + coarse quantisation + noise.

    strength 0 returns the input; ~1 is a visibly bad encode.  Only used where a real
    libavif round trip is not wanted (fast unit tests, GPU-box smoke)."""
    if strength <= 0:
        return rgb.copy()
    h, w, _ = rgb.shape
    f = rgb.astype(np.float64)
    out = np.empty_like(f)
    step = 1.0 + 14.0 * strength
    for c in range(3):
        p = f[..., c]
        b = _box3(p)
        p = (1 - min(1.0, strength)) * p + min(1.0, strength) * b
        p = np.rint(p / step) * step
        p = p + (_uniform(seed, 32 + c, h, w) - 0.5) * 6.0 * strength
        out[..., c] = p
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def rgb8_to_yuv444(rgb: np.ndarray, depth: int = 10, matrix: int = 2) -> tuple[np.ndarray, ...]:
    """Full-range RGB8 -> YUV444 planes at 8/10 bits (float BT.601/709/2020 forward matrix).

    A stand-in for a decoder's output planes when a synthetic distorted frame is needed in
    YUV form; the forward transform is NOT part of the scored path."""
    kr, kb = {1: (0.2126, 0.0722), 9: (0.2627, 0.0593)}.get(matrix, (0.299, 0.114))
    kg = 1.0 - kr - kb
    f = rgb.astype(np.float64) / 255.0
    Y = kr * f[..., 0] + kg * f[..., 1] + kb * f[..., 2]
    U = (f[..., 2] - Y) / (2 * (1 - kb))
    V = (f[..., 0] - Y) / (2 * (1 - kr))
    mx = (1 << depth) - 1
    half = 1 << (depth - 1)
    dt = np.uint8 if depth == 8 else np.uint16
    q = lambda p, off: np.clip(np.rint(p * mx + off), 0, mx).astype(dt)
    return q(Y, 0), q(U, half), q(V, half)
