#!/usr/bin/env python
"""Runs a few eager (no CUDA graph) training iterations of config 2 so that ncu sees every kernel as its own launch.
Usage (GPU box):  ncu ... python tools/profile_iter.py [--iters 2 --warm 1 --batch 256 --timesteps 3 --vocab 2000]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--timesteps", type=int, default=3)
    ap.add_argument("--vocab", type=int, default=2000)
    ap.add_argument("--critic-iters", type=int, default=5)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--warm", type=int, default=1)
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from sgg_b200._lib import lib
    from sgg_b200.trainer import HotPathTrainer
    L = lib()
    import ctypes as C
    L.sgg_launch_count.restype = C.c_int64
    tr = HotPathTrainer(a.batch, a.timesteps, a.vocab, critic_iters=a.critic_iters, use_graph=False)
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn(a.batch, 196, 512, generator=g).bfloat16().cuda(), torch.randn(a.batch, 196, 512, generator=g).bfloat16().cuda(),
                torch.randint(0, a.vocab, (a.batch, a.timesteps), generator=g).cuda()) for _ in range(2)]
    n0 = L.sgg_launch_count()
    for i in range(a.warm + a.iters):
        tr.set_batch(*batches[i % 2])
        tr.iteration()
        torch.cuda.synchronize()
        print(f"iteration {i}: launches so far {L.sgg_launch_count() - n0}", flush=True)
    print("losses", tr.losses())


if __name__ == "__main__":
    main()
