#!/usr/bin/env python
"""A small end-to-end case for compute-sanitizer (memcheck): odd sizes, both blurs, batch, YUV paths."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oavif_b200.host import ssimu2, synth
for (w, h) in [(100, 75), (333, 257), (65, 130), (640, 360)]:
    src = synth.synth(w, h, "mixture", 1)
    d1, d2 = synth.distort(src, 0.2), synth.distort(src, 0.6)
    with ssimu2.Scorer(w, h, 3) as sc:
        for mode in (ssimu2.BLUR_RECURSIVE, ssimu2.BLUR_FIR):
            sc.set_blur(mode)
            sc.set_source(src)
            a = sc.score_batch_rgb8([d1, d2, src])
            y, u, v = synth.rgb8_to_yuv444(d1, 10)
            b = sc.score_yuv444(y, u, v, 10)
            y8, u8, v8 = synth.rgb8_to_yuv444(d2, 8)
            c = sc.score_yuv444(y8, u8, v8, 8, 1, True)
            print(w, h, mode, [round(x, 4) for x in a], round(b, 4), round(c, 4))
        sc.yuv444_to_rgb8(y, u, v, 10)
print("sanitize case done")
