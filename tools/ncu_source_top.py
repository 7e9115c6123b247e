#!/usr/bin/env python
"""Top stall-sampled SASS instructions per kernel from an `ncu --page source --csv` export (optionally .gz)."""
import csv
import gzip
import sys

path, nk, ntop = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 99, int(sys.argv[3]) if len(sys.argv) > 3 else 14
op = gzip.open if path.endswith(".gz") else open
txt = op(path, "rt", errors="ignore").read()
blocks = txt.split('"Kernel Name",')[1:]
for bi, b in enumerate(blocks[:nk]):
    lines = b.split("\n")
    rd = list(csv.reader(lines[1:]))
    hdr = rd[0]
    rows = [r for r in rd[1:] if len(r) == len(hdr)]
    ci = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ci["# Samples"]]) for r in rows)
    print("=====", bi, lines[0][:100], "| SASS rows", len(rows), "| samples", tot)
    for r in sorted(rows, key=lambda r: -int(r[ci["# Samples"]]))[:ntop]:
        st = {h[6:]: int(r[ci[h]]) for h in hdr
              if h.startswith("stall_") and "Not Issued" not in h and r[ci[h]].isdigit() and int(r[ci[h]]) > 0}
        print(r[ci["# Samples"]].rjust(7), r[ci["Source"]].strip()[:64].ljust(64), st)
