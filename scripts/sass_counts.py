#!/usr/bin/env python
"""Count the Blackwell tile-movement, barrier and packed-f32 instructions per kernel in the built library.

    python scripts/sass_counts.py > profiles/r2_sass_tma.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oavif_b200", "lib", "liboavif_ssimu2.so")
WANT = ("UTMALDG", "UTMASTG", "UBLKCP", "UTMACCTL", "UTMACMDFLUSH", "SYNCS", "NANOSLEEP", "LDGSTS", "BAR", "FFMA2", "FMUL2", "FADD2")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = {}
    counts = collections.defaultdict(collections.Counter)
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            if cur not in names:
                names[cur] = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            for w in WANT:
                if op == w or op.startswith(w + "."):
                    key = op if w in ("SYNCS", "NANOSLEEP") else w
                    counts[cur][key] += 1
    print("# cuobjdump -sass oavif_b200/lib/liboavif_ssimu2.so: Blackwell tile-movement, barrier and packed-f32 instructions "
          "per kernel (scripts/sass_counts.py)")
    for fn in sorted(names, key=lambda f: names[f]):
        for k in sorted(counts[fn]):
            print(f" {names[fn]} {k} : {counts[fn][k]}")


if __name__ == "__main__":
    main()
