#!/usr/bin/env python
"""Host-side cost of the two calls of a step: wall time per call on an image so small that the GPU work is
a few tens of microseconds (GPU only)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oavif_b200.host import ssimu2, synth

for (w, h) in ((64, 64), (256, 256)):
    src = synth.synth(w, h, "mixture", 0)
    dist = synth.distort(src, 0.3)
    ds, dd = torch.from_numpy(src).cuda(), torch.from_numpy(dist).cuda()
    with ssimu2.Scorer(w, h, 1) as sc:
        sc.set_source_dev(ds.data_ptr(), w, h, 3 * w)
        sc.score_batch_dev("rgb8", [[dd.data_ptr()]], [3 * w])
        n = 2000
        t0 = time.perf_counter()
        for _ in range(n):
            sc.set_source_dev(ds.data_ptr(), w, h, 3 * w)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(n):
            sc.score_batch_dev("rgb8", [[dd.data_ptr()]], [3 * w])
        t2 = time.perf_counter()
        for _ in range(n):
            sc.set_source_dev(ds.data_ptr(), w, h, 3 * w)
            sc.score_batch_dev("rgb8", [[dd.data_ptr()]], [3 * w])
        t3 = time.perf_counter()
        tm = sc.timing()
        print(f"{w}x{h}: set_source_dev {1e6 * (t1 - t0) / n:.1f} us/call (async), score (cached source) {1e6 * (t2 - t1) / n:.1f} us/call, "
              f"pair {1e6 * (t3 - t2) / n:.1f} us/step; device time of the last score call {1e3 * tm.total_ms:.1f} us")
