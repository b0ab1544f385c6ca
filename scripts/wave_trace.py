#!/usr/bin/env python
"""Timeline of the fused kernel's wavefront on a 4K pair (scale 0, channel 0): when each strip starts working, how long
its walk takes, how far it lags its left neighbour."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oavif_b200.host import ssimu2, synth
W, H = 3840, 2160
base = synth.synth(1920, 1080, "mixture", 0)
src = np.tile(base, (2, 2, 1)); dist = np.tile(synth.distort(base, 0.3), (2, 2, 1))
with ssimu2.Scorer(W, H, 1) as sc:
    sc.set_tile_path(ssimu2.TILES_FUSED)
    sc.set_source(src); sc.score_rgb8(dist)
    for mode in (2, 1):
        sc.wave_trace(mode)
        tr = sc.wave_trace(mode).astype(np.int64)
        t0 = tr[:, 1].min()
        unit = tr[:, 0]
        s, c, t = unit & 15, (unit >> 4) & 15, unit >> 8
        sel = (s == 0) & (c == 0)
        rows = tr[sel]; ts = t[sel]
        order = np.argsort(ts)
        rows, ts = rows[order], ts[order]
        start, pro, mid, end = [(rows[:, i] - t0) / 1e3 for i in (1, 2, 3, 4)]
        print(f"mode {mode}: kernel span {(tr[:,4].max()-t0)/1e3:.1f} us, {len(tr)} CTAs; scale 0 / channel 0: {len(ts)} strips")
        print("  strip   start  prologue_end   mid      end    walk(pro->end)  lag of prologue_end vs left")
        for i in list(range(0, 8)) + list(range(8, len(ts), 16)) + [len(ts) - 1]:
            lag = pro[i] - pro[i - 1] if i else 0.0
            print(f"  {ts[i]:5d} {start[i]:8.1f} {pro[i]:10.1f} {mid[i]:10.1f} {end[i]:8.1f} {end[i]-pro[i]:10.1f} {lag:10.2f}")
        print(f"  mean lag per hop {np.diff(pro).mean():.2f} us; mean walk {np.mean(end-pro):.1f} us; phases per walk {(H+4+15)//16+1}")
