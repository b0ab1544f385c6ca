#!/usr/bin/env python
"""Given tests/golden/fssimu2_scores.json (made by this directory's recipe on a machine with zig), say WHICH reading
of the published algorithm fssimu2 0.1.1 follows: every combination of the oracle's variant switches is scored on
the same pairs and ranked by max |oracle - fssimu2|.

    python scripts/pin_fssimu2/which_variant.py [tests/golden/fssimu2_scores.json]

Readings (oracle/ssimu2_oracle.h, scripts/variant_envelope.py): vertical-pass operation order, sRGB table from
binary32 powf, binary32 maps and pooling, libm cbrtf, the weight layout below six scales, FIR instead of the recursion.
The combination that comes out at ~1e-6 is the one the oracle's default (and then the CUDA kernels, which are held to
the oracle bit for bit) has to adopt.  CPU only; test infrastructure.
"""
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oavif_b200.host import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

FLAGS = (("vertical_order", O.VARIANT_VERTICAL_ORDER), ("f32_transfer", O.VARIANT_F32_TRANSFER),
         ("f32_maps", O.VARIANT_F32_MAPS), ("contiguous_weights", O.VARIANT_CONTIGUOUS_WEIGHTS))


def rank(records, pairs=None):
    """[(max |delta|, mean |delta|, description)] sorted best first.  `pairs`: optional precomputed (src, dst) list."""
    if pairs is None:
        pairs = []
        for c in records:
            src = synth.synth(c["w"], c["h"], c["kind"], c["seed"])
            pairs.append((src, synth.distort(src, c["strength"], seed=c["seed"] + 100)))
    out = []
    for blur, bname in ((O.BLUR_IIR, "recursive"), (O.BLUR_FIR, "fir")):
        for libm in (False, True):
            for r in range(len(FLAGS) + 1):
                for combo in itertools.combinations(FLAGS, r):
                    flags = 0
                    for _, f in combo:
                        flags |= f
                    if blur == O.BLUR_FIR and flags & O.VARIANT_VERTICAL_ORDER:
                        continue            # the switch only exists in the recursion
                    O.set_variant(flags, fast=True, libm_cbrt=libm)
                    d = [abs(O.ssimu2_rgb8(s, t, blur, fast=True) - c["score"]) for (s, t), c in zip(pairs, records)]
                    name = " + ".join([bname] + (["libm_cbrt"] if libm else []) + [n for n, _ in combo])
                    out.append((max(d), sum(d) / len(d), name))
    O.set_variant(0, fast=True)
    return sorted(out)


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "fssimu2_scores.json")
    if not os.path.exists(path):
        sys.exit(f"{path} not found: produce it with the recipe in scripts/pin_fssimu2.md (needs zig 0.15.1 and the 0.1.1 tarball)")
    records = json.load(open(path))
    for mx, mean, name in rank(records)[:12]:
        print(f"max |d| {mx:10.6f}   mean |d| {mean:10.6f}   {name}")


if __name__ == "__main__":
    main()
