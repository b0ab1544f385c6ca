#!/usr/bin/env python
"""Write the raw RGB8 pairs fssimu2 is to score (see ../pin_fssimu2.md).  Pure function of the table below."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oavif_b200.host import synth  # noqa: E402

# (w, h, kind, seed, strength): sizes with six scales and with fewer (the weight-layout question), every kind
CASES = [(64, 64, "mixture", 0, 0.3), (100, 75, "noise", 1, 0.1), (257, 129, "edges", 2, 0.5), (320, 240, "gradient", 3, 0.05),
         (333, 257, "mixture", 4, 1.0), (640, 360, "mixture", 5, 0.2), (1027, 771, "noise", 6, 0.02), (31, 200, "noise", 7, 0.4),
         (1920, 1080, "gradient", 8, 0.3), (1920, 1080, "edges", 9, 0.3), (1920, 1080, "noise", 10, 0.3),
         (1920, 1080, "mixture", 11, 0.3), (1024, 1024, "mixture", 0, 0.6), (3840, 2160, "mixture", 0, 0.25)]


def main(out):
    os.makedirs(out, exist_ok=True)
    index = []
    for i, (w, h, kind, seed, strength) in enumerate(CASES):
        src = synth.synth(w, h, kind, seed)
        dst = synth.distort(src, strength, seed=seed + 100)
        name = f"{i:03d}"
        src.tofile(os.path.join(out, name + "_src.rgb"))
        dst.tofile(os.path.join(out, name + "_dst.rgb"))
        index.append({"name": name, "w": w, "h": h, "kind": kind, "seed": seed, "strength": strength})
    json.dump(index, open(os.path.join(out, "index.json"), "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "pairs")
