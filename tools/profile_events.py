#!/usr/bin/env python
"""Warm, in-sequence per-kernel durations of one training iteration measured with CUDA events around every launch
(SGG_TIMING=1; eager launches, PDL off so that intervals do not overlap).  Complements the ncu launch list, whose
times are cold-cache.  Usage (on a GPU box):  python tools/profile_events.py [--batch 256 --timesteps 3 --vocab 2000]"""
import argparse
import collections
import ctypes as C
import os
import sys

os.environ["SGG_TIMING"] = "1"
os.environ.setdefault("SGG_PDL", "0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--timesteps", type=int, default=3)
    ap.add_argument("--vocab", type=int, default=2000)
    ap.add_argument("--critic-iters", type=int, default=5)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from sgg_b200._lib import lib
    from sgg_b200.trainer import HotPathTrainer
    L = lib()
    L.sgg_timing_report.restype = C.c_int64
    tr = HotPathTrainer(a.batch, a.timesteps, a.vocab, critic_iters=a.critic_iters, use_graph=False)
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn(a.batch, 196, 512, generator=g).bfloat16().cuda(), torch.randn(a.batch, 196, 512, generator=g).bfloat16().cuda(),
                torch.randint(0, a.vocab, (a.batch, a.timesteps), generator=g).cuda()) for _ in range(2)]
    buf = C.create_string_buffer(1 << 20)
    for i in range(2):                       # warm-up
        tr.set_batch(*batches[i % 2]); tr.iteration()
    L.sgg_timing_report(buf, C.c_int64(len(buf)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.iters):
        tr.set_batch(*batches[i % 2]); tr.iteration()
    e1.record()
    L.sgg_timing_report(buf, C.c_int64(len(buf)))
    wall = e0.elapsed_time(e1) * 1e3 / a.iters
    rows = []
    for line in buf.value.decode().strip().splitlines():
        name, grid, block, n, us = line.rsplit(";", 4)
        name = name.replace("void ", "").replace("sgg::", "").split("(")[0]
        rows.append((name, grid, block, int(n) / a.iters, float(us) / a.iters))
    tot = sum(r[4] for r in rows)
    print(f"B={a.batch} T={a.timesteps} V={a.vocab} n_critic={a.critic_iters}: {sum(r[3] for r in rows):.0f} launches / iteration, "
          f"sum of event-timed kernel intervals {tot:.0f} us / iteration (eager wall {wall:.0f} us incl. event overhead)\n")
    print("| kernel | grid | block | launches/iter | us/iter | us/launch | share |\n|---|---|---|---|---|---|---|")
    for r in sorted(rows, key=lambda r: -r[4]):
        print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]:.0f} | {r[4]:.1f} | {r[4] / r[3]:.1f} | {100 * r[4] / tot:.1f}% |")
    fam = collections.defaultdict(float)
    for r in rows:
        fam[r[0].split("<")[0]] += r[4]
    print("\n| kernel family | us/iter | share |\n|---|---|---|")
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]):
        print(f"| {k} | {v:.1f} | {100 * v / tot:.1f}% |")


if __name__ == "__main__":
    main()
