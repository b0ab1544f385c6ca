#!/usr/bin/env python
"""Time the recursive rows pass alone: both halves vs. the candidate half only (GPU only)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oavif_b200.host import ssimu2
import torch
(src, yuv), = bench.make_pairs(0, 1)
with ssimu2.Scorer(bench.W, bench.H, 1) as sc:
    sc.set_source(src)
    sc.score_yuv444(*yuv, 10)
    names = {0: "source + candidate halves", 4: "candidate half only"}
    for v, nm in names.items():
        sc.time_rows(v, 3)
        print(f"variant {v} ({nm}): {sc.time_rows(v, 20):.4f} ms")
