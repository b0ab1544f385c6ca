#!/usr/bin/env python
"""First contact of the fused kernel with a GPU: tiny sizes first, each against the two-pass kernels, then timing."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oavif_b200.host import ssimu2, synth

def check(w, h):
    src = synth.synth(w, h, "mixture", 3)
    d = [synth.distort(src, 0.3, seed=1), synth.distort(src, 0.7, seed=2)]
    with ssimu2.Scorer(w, h, 2) as sc:
        sc.set_source(src); ref = [sc.score_rgb8(x) for x in d]; rs = sc.sums(0).copy()
        sc.set_tile_path(ssimu2.TILES_FUSED)
        sc.set_source(src); got = [sc.score_rgb8(x) for x in d]; gs = sc.sums(0).copy()
        sc.set_source(src); bat = sc.score_batch_rgb8(d)
        print(f"{w}x{h}: two-pass {ref}  fused {got}  batch {bat}  sums equal {np.array_equal(rs, gs)}", flush=True)
        return ref == got == bat

ok = True
for size in [(32, 32), (64, 64), (65, 63), (100, 75), (333, 257), (1027, 771), (1920, 1080), (3840, 2160)]:
    ok = check(*size) and ok
print("ALL EQUAL" if ok else "MISMATCH", flush=True)
W, H = 3840, 2160
base = synth.synth(1920, 1080, "mixture", 0)
src = np.tile(base, (2, 2, 1)); dist = np.tile(synth.distort(base, 0.3), (2, 2, 1))
with ssimu2.Scorer(W, H, 1) as sc:
    sc.set_tile_path(ssimu2.TILES_FUSED)
    sc.set_source(src); sc.score_rgb8(dist)
    for name, bits in (("fused, all five quantities", 1024), ("fused, cached source blur", 2048)):
        print(name, " ".join(f"{sc.time_rows(bits, 50):.4f}" for _ in range(3)), flush=True)
    sc.set_tile_path(ssimu2.TILES_TMA); sc.set_source(src); sc.score_rgb8(dist)
    for name, bits in (("two-pass rows, both halves on two streams", 256), ("two-pass rows, candidate half", 4), ("two-pass columns", 512)):
        print(name, " ".join(f"{sc.time_rows(bits, 50):.4f}" for _ in range(3)), flush=True)
