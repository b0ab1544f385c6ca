#!/usr/bin/env python
"""A few 4K evaluations (RGB8 source, 10-bit YUV candidate) for a launch-list pass restricted to k_pyramid:
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_pyramid python scripts/pyr_time.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oavif_b200.host import ssimu2, synth
w, h = 3840, 2160
src = synth.synth(w, h, "mixture", 0); d = synth.distort(src, 0.3); y, u, v = synth.rgb8_to_yuv444(d, 10)
with ssimu2.Scorer(w, h, 1) as sc:
    for i in range(4):
        sc.set_source(src)
        print(sc.score_yuv444(y, u, v, 10))
