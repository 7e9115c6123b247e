#!/usr/bin/env python
"""Micro-timings of sgg_gemm on the shapes of the training step (back-to-back launches, CUDA events).
   python tools/gemm_micro.py            # on a GPU box"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    import __graft_entry__ as ge
    ge.build()
    from sgg_b200 import ops
    dev = "cuda"
    torch.manual_seed(0)

    def run(tag, M, N, K, a_mn, b_mn, nprod, splits, block_n, atomic=False, addm=False, hl=False, reps=40, bparts=False):
        KP = (K + 63) // 64 * 64
        # operands as hi/lo pairs laid out like the plan does: A [M, 2*KP] (or [2*KP... for mn-major [K, 2*MP])
        if not a_mn:
            A = torch.randn(M, 2 * KP, device=dev).bfloat16()
            a_seg = [(0, 0), (KP, 0)]
        else:
            MP = (M + 63) // 64 * 64
            A = torch.randn(K, 2 * MP, device=dev).bfloat16()
            a_seg = [(0, 0), (0, MP)]
        if not b_mn:
            NPad = (N + 63) // 64 * 64
            Bm = torch.randn(2 * NPad, KP, device=dev).bfloat16()
            b_seg = [(0, 0), (0, NPad)]
        else:
            NP8 = (N + 7) // 8 * 8
            Bm = torch.randn(2 * KP, NP8, device=dev).bfloat16()
            b_seg = [(0, 0), (KP, 0)]
        segs = [(a_seg[0][0], a_seg[0][1], b_seg[0][0], b_seg[0][1], K)]
        if bparts:      # single A part, hi/lo B (the W_a contractions)
            segs.append((a_seg[0][0], a_seg[0][1], b_seg[1][0], b_seg[1][1], K))
        elif nprod >= 2:
            segs.append((a_seg[1][0], a_seg[1][1], b_seg[0][0], b_seg[0][1], K))
        if nprod >= 3:
            segs.append((a_seg[0][0], a_seg[0][1], b_seg[1][0], b_seg[1][1], K))
        out = torch.zeros(M, (N + 63) // 64 * 64, device=dev)
        add = torch.randn(256, out.shape[1], device=dev) if addm else None
        ohl = torch.zeros(M, 2 * out.shape[1], device=dev, dtype=torch.bfloat16) if hl else None

        def call():
            ops.gemm(A, Bm, M, N, a_mn=a_mn, b_mn=b_mn, segs=segs, out=None if hl else out[:, :N], atomic=atomic,
                     out_hl=ohl, lo_off=out.shape[1], addm=add, add_mod=256 if addm else 0, splits=splits, block_n=block_n)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        # host launch cost (python + tensor-map encoding) would dominate small kernels: replay a captured graph instead
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                call()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        fl = 2.0 * M * N * K * nprod
        print(f"{tag:34s} M={M:6d} N={N:5d} K={K:6d} prod={nprod} splits={splits} bn={block_n:3d} : {us:7.2f} us  {fl / us / 1e6:8.1f} TFLOP/s", flush=True)

    for pdl in ("1", "0"):
        os.environ["SGG_PDL"] = pdl   # read once per process: only the first value is effective
        break
    print("PDL", os.environ.get("SGG_PDL"))
    # sweep (splits, block_n) for the shapes of the time loop; atomic=2 = output cleared by the producer kernel (no zero-fill launch)
    sweeps = [
        ("scores M=768", 768, 196, 512, False, True, dict(addm=True, atomic=2)),
        ("scores M=256", 256, 196, 512, False, True, dict(addm=True, atomic=2)),
        ("gates M=768", 768, 2048, 1344, False, True, dict(atomic=2)),
        ("gates M=256", 256, 2048, 1344, False, True, dict(atomic=2)),
        ("gates G M=1280", 1280, 2048, 1536, False, True, dict(atomic=2)),
        ("x_bar M=1024", 1024, 1324, 2048, False, False, dict(atomic=2)),
        ("x_bar M=256", 256, 1324, 2048, False, False, dict(atomic=2)),
        ("c_bar M=1024", 1024, 512, 256, False, False, dict(atomic=1)),
        ("c_bar M=256", 256, 512, 256, False, False, dict(atomic=1)),
        ("dK", 1324, 2048, 3072, True, True, dict(atomic=1)),
        ("dW_h", 512, 196, 3072, True, True, dict(atomic=1)),
    ]
    if os.environ.get("GEMM_MICRO_ONLY") == "dWa":
        for sp, bn in ((0, 0), (1, 256), (1, 128), (1, 64)):
            run("dW_a", 100352, 196, 256, True, True, 2, sp, bn, bparts=True, reps=10)
        return
    if os.environ.get("GEMM_MICRO_ONLY"):
        sweeps = [x for x in sweeps if x[0].startswith(os.environ["GEMM_MICRO_ONLY"])]
    for tag, M, N, K, amn, bmn, kw in sweeps:
        run(tag + " auto", M, N, K, amn, bmn, 3, 0, 0, **kw)
        for bn in (64, 128, 256):
            if bn > 64 and N <= bn // 2:
                continue
            for sp in (1, 2, 3, 4, 6, 8):
                tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
                if tiles * sp > 3 * 148 or sp > (K + 63) // 64:
                    continue
                run(tag, M, N, K, amn, bmn, 3, sp, bn, **kw)
    return


if __name__ == "__main__":
    main()
