#!/usr/bin/env python
"""Does batched (speculative) probing earn its place?  Sequential search against --batch 2 / 3 / 4 on single images
with idle host cores (the only situation where extra encodes are free): wall time per image, passes the policy
consumed, probes made, probes wasted.  Same q / same bytes is asserted for every image (tq.hpp replays the
sequential decisions).

    python scripts/batched_probe.py [--images 12] [--out profiles/r2_batched_probe.json]     (on a GPU box)
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oavif_b200.host import harness as H, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=12)
ap.add_argument("--out", default="")
a = ap.parse_args()
o = H.default_opts(tenbit=0, score_tgt=80.0, speed=9, max_pass=6)
cases = [("cfg1 1024x1024 mixture", synth.synth(1024, 1024, "mixture", 0))]
cases += [(f"cfg5 1920x1080 seed {s} kind {s % 4}", synth.synth(1920, 1080, s % 4, s)) for s in range(a.images)]
H.search_image(cases[0][1], o, device=0, want_bytes=False)   # warm-up
rows = []
agg = {w: {"ms": 0.0, "probes": 0, "wasted": 0, "passes": 0} for w in (1, 2, 3, 4)}
for name, img in cases:
    ref, ref_bytes = H.search_image(img, o, batch_width=1, device=0)
    row = {"image": name, "q": ref.q, "passes": ref.num_pass, "sequential_ms": round(ref.total_ms, 1)}
    agg[1]["ms"] += ref.total_ms
    agg[1]["passes"] += ref.num_pass
    agg[1]["probes"] += ref.num_pass
    for w in (2, 3, 4):
        r, b = H.search_image(img, o, batch_width=w, device=0)
        assert (r.q, r.num_pass, r.history()) == (ref.q, ref.num_pass, ref.history()) and b == ref_bytes, name
        row[f"batch{w}"] = {"ms": round(r.total_ms, 1), "device_passes": r.device_passes, "probes": r.probes, "wasted": r.wasted}
        agg[w]["ms"] += r.total_ms
        agg[w]["probes"] += r.probes
        agg[w]["wasted"] += r.wasted
        agg[w]["passes"] += r.num_pass
    rows.append(row)
    print(row, file=sys.stderr)
n = len(cases)
out = {"host_cores": os.cpu_count(), "images": n,
       "mean_ms_per_image": {f"batch{w}" if w > 1 else "sequential": round(v["ms"] / n, 1) for w, v in agg.items()},
       "encodes_per_image": {f"batch{w}" if w > 1 else "sequential": round(v["probes"] / n, 2) for w, v in agg.items()},
       "wasted_share": {f"batch{w}": round(v["wasted"] / max(1, v["probes"]), 3) for w, v in agg.items() if w > 1},
       "rows": rows}
text = json.dumps(out, indent=1)
print(text)
if a.out:
    open(os.path.join(ROOT, a.out), "w").write(text + "\n")
