#!/usr/bin/env python
"""Where the first milliseconds of a context go: creation, first set_source + score, later ones (1024x1024, 8-bit YUV)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oavif_b200.host import ssimu2, synth
src = synth.synth(1024, 1024, "mixture", 0); d = synth.distort(src, 0.3); y, u, v = synth.rgb8_to_yuv444(d, 8)
for rep in range(3):
    t0 = time.perf_counter(); sc = ssimu2.Scorer(1024, 1024, 1); t1 = time.perf_counter()
    ts = []
    for i in range(4):
        a = time.perf_counter(); sc.set_source(src); b = time.perf_counter(); sc.score_yuv444(y, u, v, 8); c = time.perf_counter()
        ts.append((round((b - a) * 1e3, 3), round((c - b) * 1e3, 3)))
    sc.close()
    print(f"context {rep}: create {1e3*(t1-t0):.1f} ms; (set_source ms, score ms) x4: {ts}")
