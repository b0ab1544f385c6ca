#!/usr/bin/env python
"""A/B timing of the rows pass on a 4K pair: TMA instances (ring depth / staging buffers) against the cp.async kernel.

    python scripts/rows_ab.py        (on a B200; prints mean ms over 50 launches, three repeats)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oavif_b200.host import ssimu2, synth  # noqa: E402

W, H = 3840, 2160
base = synth.synth(1920, 1080, "mixture", 0)
src = np.tile(base, (2, 2, 1))
dist = np.tile(synth.distort(base, 0.3), (2, 2, 1))
SHAPES = {0: "stages 4, bufs 3", 1: "stages 4, bufs 2", 2: "stages 6, bufs 3", 4: "4/3, NO recursion"}
with ssimu2.Scorer(W, H, 1) as sc:
    sc.set_source(src)
    s0 = sc.score_rgb8(dist)
    for shape, name in SHAPES.items():
        for what, bits in (("both in one", 0), ("cand only  ", 4), ("source only", 128)):
            ms = [sc.time_rows(bits | (shape << 4), 50) for _ in range(3)]
            print(f"tma {name:18s} {what}", " ".join(f"{m:.4f}" for m in ms))
    for cand_only in (0, 4):
        ms = [sc.time_rows(8 | cand_only, 50) for _ in range(3)]
        print(f"cp.async {'':14s}{'cand only' if cand_only else 'both     '}", " ".join(f"{m:.4f}" for m in ms))
    # columns pass: every variant is bracketed by the shipped form (A B A), so that drift over the run (the later
    # measurements of a long sequence come out slower on these boxes) shows up in the A columns instead of in B
    print("columns pass, ms: shipped | variant | shipped")
    for name, bits in (("TMA, no block barrier", 512 | 4096), ("TMA, 2 CTAs per SM", 512 | 8192),
                       ("TMA, 256 B L2 promotion", 512 | 32768), ("TMA, strip-major reads*", 512 | 16384),
                       ("cp.async loader", 512 | 8)):
        a0 = min(sc.time_rows(512, 50) for _ in range(2))
        b = min(sc.time_rows(bits, 50) for _ in range(2))
        a1 = min(sc.time_rows(512, 50) for _ in range(2))
        print(f"  {name:28s} {a0:.4f} | {b:.4f} | {a1:.4f}   variant / shipped = {2 * b / (a0 + a1):.3f}")
    sc.set_tile_path(ssimu2.TILES_CP_ASYNC)
    sc.set_source(src)
    s1 = sc.score_rgb8(dist)
    print("scores", s0, s1, s0 == s1)
