#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS mnemonics in csrc/libsgg_b200.so (cuobjdump -sass): UTCHMMA / UTCQMMA
(tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor copies), UBLKCP (bulk copies), UTCBAR
(tcgen05.commit), SYNCS (mbarrier), UCGABAR (cluster barrier).  Writes a markdown table to stdout.
    python tools/sass_summary.py > profiles/sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sgg_b200", "csrc", "libsgg_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "UCGABAR", "MUFU", "RED", "ATOM"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return [o.replace("sgg::", "").split("(")[0].replace("void ", "") for o in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op.startswith(mn):
                    counts[cur][mn] += 1
    names = demangle(order)
    size = os.path.getsize(LIB)
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(f"# SASS summary of `sgg_b200/csrc/libsgg_b200.so` ({size} bytes, built from commit {head} with "
          "`nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo`)\n")
    print("`UTCHMMA` = tcgen05.mma (kind::f16), `LDTM` = tcgen05.ld, `UTMALDG` = cp.async.bulk.tensor (TMA load), `UBLKCP` = cp.async.bulk, "
          "`UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, `UCGABAR` = barrier.cluster.  Counts are static instruction counts.\n")
    cols = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "UCGABAR", "MUFU", "RED", "_total"]
    print("| kernel | " + " | ".join(c.strip("_") for c in cols) + " |\n|---|" + "---|" * len(cols))
    tot = collections.Counter()
    for mangled, name in sorted(zip(order, names), key=lambda x: x[1]):
        c = counts[mangled]
        tot.update(c)
        print(f"| `{name}` | " + " | ".join(str(c[k]) for k in cols) + " |")
    print("| **all kernels** | " + " | ".join(str(tot[k]) for k in cols) + " |")


if __name__ == "__main__":
    main()
