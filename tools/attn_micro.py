#!/usr/bin/env python
"""Isolated timing of the attention step kernels (sgg_attn_forward / sgg_attn_reverse) over a range of batch sizes:
back-to-back launches over annotation tensors that rotate through > L2, CUDA events, best of 3 trains.  Prints us per
launch and achieved GB/s on the algorithmic bytes.  SGG_ATTN_PERSIST=0/1 selects the kernel family."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    import __graft_entry__ as ge
    ge.build()
    from sgg_b200._lib import check, lib, stream_ptr
    L = lib()
    R, nv = 196, int(os.environ.get("NV", "3"))
    print(f"SGG_ATTN_PERSIST={os.environ.get('SGG_ATTN_PERSIST', '1')} nv={nv}")
    for B in [int(x) for x in os.environ.get("BATCHES", "74,148,256,296,444,592,1184,2368").split(",")]:
        nbuf = max(2, (400 << 20) // (B * R * 1024) + 1)
        anns = [torch.randn(B, R, 512, device="cuda").bfloat16() for _ in range(nbuf)]
        E = torch.randn(nv * B, 256, device="cuda")
        alpha = torch.empty_like(E)
        X = torch.empty(nv * B, 2 * 1344, dtype=torch.bfloat16, device="cuda")
        zb = torch.randn(nv * B, 1344, device="cuda")
        EBh = torch.empty(nv * B, 512, dtype=torch.bfloat16, device="cuda")
        PB = torch.zeros(B, 256, device="cuda")

        def fwd(i):
            a = anns[i % nbuf]
            check(L.sgg_attn_forward(C.c_void_p(a.data_ptr()), C.c_int32(B), C.c_int32(R), C.c_int32(nv), C.c_void_p(E.data_ptr()),
                                     C.c_void_p(alpha.data_ptr()), C.c_int64(256), C.c_void_p(X.data_ptr()), C.c_int64(2 * 1344),
                                     C.c_int64(1344), stream_ptr()), "fwd")

        def rev(i):
            a = anns[i % nbuf]
            check(L.sgg_attn_reverse(C.c_void_p(a.data_ptr()), C.c_int32(B), C.c_int32(R), C.c_int32(nv), C.c_void_p(zb.data_ptr()),
                                     C.c_int64(1344), C.c_void_p(alpha.data_ptr()), C.c_int64(256), C.c_void_p(EBh.data_ptr()),
                                     C.c_int64(512), C.c_int64(256), C.c_void_p(PB.data_ptr()), C.c_int64(256), stream_ptr()), "rev")
        out = []
        for name, fn in (("fwd", fwd), ("rev", rev)):
            n = 24
            for i in range(4):
                fn(i)
            torch.cuda.synchronize()
            best = 1e9
            for r in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(n):
                    fn(4 + r * n + i)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e3 / n)
            nbytes = B * R * 512 * 2
            out.append(f"{name} {best:7.2f} us {nbytes / best / 1e3:6.0f} GB/s")
        print(f"B={B:5d} ({nbuf} tensors): " + " | ".join(out), flush=True)
        del anns


if __name__ == "__main__":
    main()
