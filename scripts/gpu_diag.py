#!/usr/bin/env python
"""Stage-by-stage GPU-vs-oracle report (never stops at the first mismatch).

Run on a GPU box:  python scripts/gpu_diag.py > gpurun_out/diag.txt
Uses the CPU oracle as the checker only.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oavif_b200.host import ssimu2, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def cmp_planes(name, got, want):
    if got.shape != want.shape:
        print(f"  {name}: SHAPE {got.shape} vs {want.shape}")
        return
    neq = int((got.view(np.uint32) != want.view(np.uint32)).sum())
    mad = float(np.abs(got.astype(np.float64) - want).max())
    print(f"  {name}: {'BIT-EXACT' if neq == 0 else f'{neq} of {got.size} differ'}  max|d|={mad:.3e}")
    if neq:
        idx = np.argwhere(got.view(np.uint32) != want.view(np.uint32))[:4]
        for i in idx:
            print(f"     at {tuple(i)}: gpu {got[tuple(i)]!r} oracle {want[tuple(i)]!r}")


def main():
    sizes = [(64, 64), (100, 75), (333, 257), (640, 360), (1027, 771)]
    if "--big" in sys.argv:
        sizes += [(1920, 1080), (3840, 2160)]
    for (w, h) in sizes:
        print(f"=== {w}x{h}")
        src = synth.synth(w, h, "mixture", 1)
        dist = synth.distort(src, 0.25, seed=3)
        y, u, v = synth.rgb8_to_yuv444(dist, 10)
        y8, u8, v8 = synth.rgb8_to_yuv444(dist, 8)
        with ssimu2.Scorer(w, h, 3) as sc:
            # K0 alone
            for depth, planes in ((10, (y, u, v)), (8, (y8, u8, v8))):
                for rgba in (False, True):
                    for m in (2, 1, 9):
                        got = sc.yuv444_to_rgb8(*planes, depth, m, rgba)
                        want = O.yuv444_to_rgb8(*planes, depth, m, rgba)
                        bad = int((got != want).sum())
                        if bad:
                            print(f"  yuv2rgb depth={depth} rgba={rgba} m={m}: {bad} mismatches")
            print("  yuv2rgb: checked 12 variants")
            dist_rgb = O.yuv444_to_rgb8(y, u, v, 10)
            for mode, omode, nm in ((ssimu2.BLUR_FIR, O.BLUR_FIR, "FIR"), (ssimu2.BLUR_RECURSIVE, O.BLUR_IIR, "IIR")):
                sc.set_blur(mode)
                sc.set_source(src)
                t0 = time.time()
                got = sc.score_yuv444(y, u, v, 10)
                t1 = time.time()
                want, det = O.ssimu2_rgb8(src, dist_rgb, omode, detail=True)
                print(f"  [{nm}] score gpu {got:.9f} oracle {want:.9f} diff {got - want:+.3e}  ({(t1 - t0) * 1e3:.2f} ms wall)")
                tm = sc.timing()
                print(f"     timing h2d {tm.h2d_ms:.3f} pyr {tm.pyramid_ms:.3f} blur {tm.blur_ms:.3f} fin {tm.finalize_ms:.3f} total {tm.total_ms:.3f} launches {tm.launches}")
                gs, ws = sc.sums(0), O.detail_sums(det)
                d = sc.detail(0)
                print(f"     n_scales gpu {d.n_scales} oracle {det.n_scales}")
                rel = np.abs(gs - ws) / np.maximum(np.abs(ws), 1e-30)
                rel[ws == 0] = np.abs(gs[ws == 0])
                for s in range(det.n_scales):
                    print(f"     scale {s} max rel sum err {rel[s].max():.3e}")
                if mode == ssimu2.BLUR_FIR:
                    for s in range(det.n_scales):
                        for which, img, nm2 in ((0, src, "src"), (1, dist_rgb, "dist")):
                            wx = O.xyb_at_scale(img, s)
                            for c in range(3):
                                cmp_planes(f"xyb {nm2} s{s} c{c}", sc.xyb(which, s, c), wx[c])
                # filter alone
                rng = np.random.default_rng(5)
                plane = rng.random((h, w), dtype=np.float32)
                cmp_planes(f"blur[{nm}] random plane", sc.blur(plane), O.blur(plane, omode))
                # identical pair and batch
                b = sc.score_batch_rgb8([dist_rgb, src, dist_rgb])
                print(f"     batch: {b}  (want [{want:.6f}, 100, {want:.6f}])")
    print("diag done")


if __name__ == "__main__":
    main()
