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

    def run(tag, M, N, K, a_mn, b_mn, nprod, splits, block_n, atomic=False, addm=False, hl=False, reps=40):
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
        if nprod >= 2:
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
    # fixed-cost probe: tiny K, growing K
    for K in (64, 128, 256, 512, 1024, 2048):
        run("probe 1 tile-row", 128, 64, K, False, True, 1, 1, 64)
    for K in (64, 512, 2048):
        run("probe 3 products", 128, 64, K, False, True, 3, 1, 64)
    # scores: e = P + c W_h  (M=3B, N=196, K=512)
    for sp, bn in ((0, 0), (1, 64), (1, 128), (1, 256), (2, 64), (4, 64)):
        run("scores (auto=0/0)", 768, 196, 512, False, True, 3, sp, bn, addm=True)
    # gates: q = x K (M=3B, N=2048, K=1344)
    for sp, bn in ((0, 0), (1, 256), (1, 128), (2, 256), (3, 256), (2, 128)):
        run("gates M=768", 768, 2048, 1344, False, True, 3, sp, bn)
    for sp, bn in ((0, 0), (1, 256), (1, 128), (4, 256), (7, 256), (2, 128), (4, 128)):
        run("gates M=256 (tangent)", 256, 2048, 1344, False, True, 3, sp, bn)
    # x_bar = q_bar K^T (M=4B, N=1324, K=2048), weight K-major
    for sp, bn in ((0, 0), (1, 256), (1, 128), (2, 256), (2, 128), (3, 128)):
        run("x_bar M=1024", 1024, 1324, 2048, False, False, 3, sp, bn)
    for sp, bn in ((0, 0), (1, 128), (2, 128), (4, 128), (4, 256), (8, 256)):
        run("x_bar M=256", 256, 1324, 2048, False, False, 3, sp, bn)
    # dK = X^T QB (M=1324, N=2048, K=3072) mn/mn
    for sp, bn in ((0, 0), (1, 256), (2, 256), (3, 256), (5, 256), (3, 128)):
        run("dK", 1324, 2048, 3072, True, True, 3, sp, bn, atomic=True)
    # logits h W_dec for 5 streams (M=3840, N=2000, K=512) hi/lo out
    for sp, bn in ((0, 0), (1, 256), (1, 128)):
        run("logits", 3840, 2000, 512, False, True, 3, sp, bn, hl=True)


if __name__ == "__main__":
    main()
