#!/usr/bin/env python
"""Per-instruction hot spots of one kernel from an .ncu-rep captured with --import-source on.

    python scripts/ncu_hot.py gpurun_out/prof.ncu-rep k_iir_rows [N]
"""
import csv
import io
import subprocess
import sys
from collections import Counter


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[iS].isdigit():
            if data:
                break  # first launch only
            continue
        data.append(r)
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[iS]) for r in data)
    totex = sum(int(r[iEx]) for r in data)
    print(f"{kernel}: {len(data)} SASS instructions, {totex} warp-instructions executed, {tot} samples")
    agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
    print("stall samples:", ", ".join(f"{k[6:]} {v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
    c = Counter()
    for r in data:
        toks = r[iSrc].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        c[op.split(".")[0]] += int(r[iEx])
    print("executed mix:", ", ".join(f"{k} {v * 100 / max(totex, 1):.1f}%" for k, v in c.most_common(16)))
    for r in sorted(data, key=lambda r: -int(r[iS]))[:topn]:
        reasons = {hdr[i][6:]: int(r[i] or 0) for i in stall_cols if int(r[i] or 0) > 0}
        rs = ", ".join(f"{k} {v}" for k, v in sorted(reasons.items(), key=lambda x: -x[1])[:3])
        print(f"{r[iS]:>7} {r[iEx]:>9}  {r[iSrc][:70]:70s} {rs}")


if __name__ == "__main__":
    main()
