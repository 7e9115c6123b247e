#!/usr/bin/env python
"""profiles/traffic.json from `ncu --page raw --csv` exports of a `--set full` capture: per kernel the median of
dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes).  bench.py copies these into `roofline.traffic` and names
this file as the source (they are NOT measured inside the bench run).
    python tools/make_traffic.py <commit> out.json a.raw.csv [b.raw.csv ...]"""
import csv
import json
import statistics
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
# bench.py roofline key -> (substring of the demangled kernel name, grid filter or None)
KEYS = {
    "attn_fwd_mma_kernel": ("attn_fwd_mma_kernel<0>", None),
    "attn_rev_mma_kernel": ("attn_rev_mma_kernel", None),
    "gemm_kernel_K1": ("gemm_kernel<256, 0, 1, 2>", None),
    "gemm_kernel_dWa": ("gemm_kernel<256, 1, 1, 2>", None),
    "adam_proj_kernel": ("adam_proj_kernel", None),
    "adam_kernel": ("adam_kernel", "(4802, 1, 1)"),
}


def main():
    commit, out, paths = sys.argv[1], sys.argv[2], sys.argv[3:]
    per = {k: [] for k in KEYS}
    for path in paths:
        rows = list(csv.reader(open(path, errors="ignore")))
        hdr, units = rows[0], rows[1]
        ci = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            if len(r) != len(hdr):
                continue
            name = r[ci["Kernel Name"]].replace("(int)", "").replace("(bool)", "").replace("sgg::", "")
            name = name.replace("false", "0").replace("true", "1")
            for key, (sub, grid) in KEYS.items():
                if sub in name and (grid is None or r[ci["Grid Size"]].strip() == grid):
                    rd = float(r[ci["dram__bytes_read.sum"]]) * UNIT[units[ci["dram__bytes_read.sum"]]]
                    wr = float(r[ci["dram__bytes_write.sum"]]) * UNIT[units[ci["dram__bytes_write.sum"]]]
                    per[key].append(rd + wr)
    res = {k: int(statistics.median(v)) for k, v in per.items() if v}
    res["_launches"] = {k: len(v) for k, v in per.items() if v}
    res["_source"] = (f"ncu --set full --clock-control none at commit {commit}: median dram__bytes_read.sum + dram__bytes_write.sum per launch "
                      f"over the launches of 1 training iteration at config 2's shape ({', '.join(p.split('/')[-1] for p in paths)}); "
                      "NOT measured inside the bench run")
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
