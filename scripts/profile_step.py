#!/usr/bin/env python
"""The bench workload (config 2: one full 4K evaluation per step) without the reporting, for ncu.

    python scripts/profile_step.py [--blur recursive|fir] [--steps 3]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oavif_b200.host import ssimu2  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blur", default="recursive")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import torch
    pairs = bench.make_pairs(0, 2)
    dev = [(torch.from_numpy(s).cuda(), tuple(torch.from_numpy(p.view(np.int16)).cuda() for p in yuv)) for s, yuv in pairs]
    torch.cuda.synchronize()
    with ssimu2.Scorer(bench.W, bench.H, 1, blur=ssimu2.BLUR_FIR if a.blur == "fir" else ssimu2.BLUR_RECURSIVE) as sc:
        for i in range(a.steps):
            s, (y, u, v) = dev[i % 2]
            sc.set_source_dev(s.data_ptr(), bench.W, bench.H, 3 * bench.W)
            sc.score_batch_dev("yuv444", [[y.data_ptr(), u.data_ptr(), v.data_ptr()]], [2 * bench.W] * 3, depth=10)
            t = sc.timing()
            print(f"step {i}: pyr {t.pyramid_ms:.3f} a {t.blur_a_ms:.3f} b {t.blur_b_ms:.3f} fin {t.finalize_ms:.3f} total {t.total_ms:.3f}")
        # one more candidate against the same source: the cached-source path (k_iir_rows<1>)
        s, (y, u, v) = dev[a.steps % 2]
        sc.score_batch_dev("yuv444", [[y.data_ptr(), u.data_ptr(), v.data_ptr()]], [2 * bench.W] * 3, depth=10)
        t = sc.timing()
        print(f"cached: pyr {t.pyramid_ms:.3f} a {t.blur_a_ms:.3f} b {t.blur_b_ms:.3f} fin {t.finalize_ms:.3f} total {t.total_ms:.3f}")


if __name__ == "__main__":
    main()
