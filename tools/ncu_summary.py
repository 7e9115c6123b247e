#!/usr/bin/env python
"""Summarises ncu output for profiles/: a launch list (gpu__time_duration per launch) into per-kernel shares, and a
`--page raw --csv` export of a --set full capture into the handful of metrics the roofline discussion needs.

    python tools/ncu_summary.py launches gpurun_out/launches_r1.csv [n_iterations] > profiles/r1_launches.md
    python tools/ncu_summary.py raw gpurun_out/prof_r1_hbm.raw.csv > profiles/r1_hbm_kernels.md
"""
import collections
import csv
import re
import sys

KEEP = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_shared_mem", "occ lim smem"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("sgg::", "")
    return name


def launches(path, n_iter=2.0):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        key = (short(r[4]), r[8], r[7])
        a = agg.setdefault(key, [0, 0])
        a[0] += 1
        a[1] += int(float(r[-1]))
    tot = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {tot / 1e3:.1f} us of kernel time in total ({n_iter:g} training iterations; "
          f"ncu times are cold-cache and serialised: compare shares)\n")
    print("| kernel | grid | block | launches | total us | us / launch | share |")
    print("|---|---|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k[0]} | {k[1]} | {k[2]} | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / v[0] / 1e3:.1f} | {100 * v[1] / tot:.1f}% |")
    fam = collections.defaultdict(int)
    for k, v in agg.items():
        fam[k[0].split("<")[0]] += v[1]
    print("\n| kernel family | total us | share |\n|---|---|---|")
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]):
        print(f"| {k} | {v / 1e3:.1f} | {100 * v / tot:.1f}% |")


def raw(path):
    rd = csv.reader(open(path, errors="ignore"))
    header = None
    rows = []
    for r in rd:
        if header is None:
            if "Kernel Name" in r:
                header = r
            continue
        rows.append(r)
    units, rows = rows[0], rows[1:]
    col = {h: i for i, h in enumerate(header)}
    names = [(m, lab) for m, lab in KEEP if m in col]
    print("| # | kernel | " + " | ".join(f"{lab} [{units[col[m]]}]" if units[col[m]] else lab for m, lab in names) + " |")
    print("|---|---|" + "---|" * len(names))
    for r in rows:
        print(f"| {r[col['ID']]} | {short(r[col['Kernel Name']])} | " + " | ".join(r[col[m]] for m, _ in names) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 2.0)
    else:
        raw(sys.argv[2])
