#!/usr/bin/env python
"""How far "the" SSIMULACRA2 score moves between readings of the published algorithm that this repo cannot
settle without the reference's scorer (fssimu2 0.1.1 is un-vendored, /root/reference/build.zig.zon:7-10).

    python scripts/variant_envelope.py [--images 12] [--out profiles/r2_variant_envelope.json]

Pairs: procedural 1920x1080 corpus images (seed = index, kind = seed mod 4 — config 5) against their REAL AV1
round trips (libaom through the bundled libavif, 8-bit YUV444, speed 9, q in {40, 65, 85}) decoded to RGB8 by
libavif itself.  Every pair is scored by the CPU oracle (TEST INFRASTRUCTURE, oracle/) under:

  default            the restatement the CUDA path is held to bit for bit
  vertical_order     the vertical recursion step as fma(n2, sum, fma(-d1, y1, -y2)) (lib/jxl gauss_blur.cc
                     VerticalBlock as recalled) instead of the horizontal pass's sequence
  libm_cbrt          libm cbrtf instead of the fixed binary32 sequence
  f32_transfer       the sRGB transfer function evaluated in binary32 (powf) instead of binary64 rounded once
  f32_maps           the SSIM / edge-diff maps and their fourth powers in binary32 with one binary32 accumulator per
                     image row (a vectorised f32 scorer) instead of binary64 per pixel
  fir                the exactly equivalent 9-tap FIR instead of the recursion (what a non-recursive blur gives)
  contracted         the same source compiled with -ffp-contract=fast -march=native: every a*b + c the compiler sees
                     is fused (what a tool chain that contracts by default makes of the same code)
  contiguous_weights the running weight index over the scales present (only differs below six scales: measured on
                     128x96 crops)

The maximum |delta score| per variant is the honest error bar on "matches fssimu2 to 0.05": any of these could be
what fssimu2 does.  This runs on the CPU only.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oavif_b200.host import harness as H, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=12)
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("--out", default="profiles/r2_variant_envelope.json")
    a = ap.parse_args()
    w, h = (int(x) for x in a.size.split("x"))
    O.build()
    opts = H.default_opts(tenbit=0, speed=9)
    variants = {"vertical_order": dict(flags=O.VARIANT_VERTICAL_ORDER), "libm_cbrt": dict(flags=0, libm=True),
                "f32_maps": dict(flags=O.VARIANT_F32_MAPS), "f32_transfer": dict(flags=O.VARIANT_F32_TRANSFER),
                "fir": dict(flags=0, blur=O.BLUR_FIR)}
    deltas = {k: [] for k in variants}
    deltas["contracted"] = []
    deltas["contiguous_weights_small_images"] = []
    rows = []
    for seed in range(a.images):
        src = synth.synth(w, h, seed % 4, seed)
        for q in (40, 65, 85):
            dec = H.decode_rgb8(H.encode(src, q, opts), w, h)
            O.set_variant(0, fast=True)
            base = O.ssimu2_rgb8(src, dec, O.BLUR_IIR, fast=True)
            row = {"seed": seed, "kind": seed % 4, "q": q, "default": base}
            for name, v in variants.items():
                O.set_variant(v.get("flags", 0), fast=True, libm_cbrt=v.get("libm", False))
                s = O.ssimu2_rgb8(src, dec, v.get("blur", O.BLUR_IIR), fast=True)
                O.set_variant(0, fast=True)
                row[name] = s
                deltas[name].append(s - base)
            sc = O.ssimu2_rgb8_contracted(src, dec)
            if sc is not None:
                row["contracted"] = sc
                deltas["contracted"].append(sc - base)
            # fewer than six scales: a crop (the weight layout is the only thing that changes)
            cs, cd = np.ascontiguousarray(src[:96, :128]), np.ascontiguousarray(dec[:96, :128])
            b2 = O.ssimu2_rgb8(cs, cd, O.BLUR_IIR, fast=True)
            O.set_variant(O.VARIANT_CONTIGUOUS_WEIGHTS, fast=True)
            c2 = O.ssimu2_rgb8(cs, cd, O.BLUR_IIR, fast=True)
            O.set_variant(0, fast=True)
            row["crop128x96_default"], row["crop128x96_contiguous"] = b2, c2
            deltas["contiguous_weights_small_images"].append(c2 - b2)
            rows.append(row)
            print(f"seed {seed} q{q}: {base:.4f} " + " ".join(f"{k} {row[k] - base:+.4f}" for k in variants), file=sys.stderr)

    def stats(d):
        d = np.abs(np.array(d))
        return {"max_abs": float(d.max()), "mean_abs": float(d.mean()), "p95_abs": float(np.percentile(d, 95)),
                "share_above_0.05": float((d > 0.05).mean())}

    out = {"pairs": len(rows), "size": [w, h], "what": __doc__.split("\n\n")[2].strip(),
           "envelope": {k: stats(v) for k, v in deltas.items() if v}, "rows": rows}
    text = json.dumps(out, indent=1)
    with open(os.path.join(ROOT, a.out), "w") as f:
        f.write(text + "\n")
    print(json.dumps(out["envelope"], indent=1))


if __name__ == "__main__":
    main()
