#!/usr/bin/env python
"""How faithful is the tcgen05 fp32 accumulator over a long contraction?  C = A B^T with bf16-exact operands (so the products
are exact and fp64 is the truth), M 128, N 256, K from 2 k to 256 k, one accumulator per output (splits = 1) against split-K
(fp32 partial sums combined by IEEE red.add in L2).  Reports the relative L2 error of C and its mean signed error in units
of the rms magnitude of C (a bias shows a truncating accumulator).  Writes gpurun_out/<TAG>_accum_micro.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    import __graft_entry__ as ge
    ge.build()
    from sgg_b200 import ops
    out = {}
    g = torch.Generator().manual_seed(0)
    M, N = 128, 256
    for K in (2048, 8192, 32768, 131072, 262144):
        for dist in ("signed", "positive"):
            A = torch.randn(M, K, generator=g)
            Bm = torch.randn(N, K, generator=g)
            if dist == "positive":
                A, Bm = A.abs(), Bm.abs()
            A, Bm = A.bfloat16().cuda(), Bm.bfloat16().cuda()
            ref = A.double() @ Bm.double().T
            row = {}
            for splits in (1, 4, 16, 64):
                if K // 64 < 2 * splits:
                    continue
                C = torch.zeros(M, N, device="cuda")
                ops.gemm(A, Bm, M, N, segs=[(0, 0, 0, 0, K)], out=C, splits=splits, block_n=256)
                torch.cuda.synchronize()
                err = C.double() - ref
                row[f"splits_{splits}"] = {"rel_l2": float(err.norm() / ref.norm()),
                                           "mean_signed_over_rms": float(err.mean() / ref.pow(2).mean().sqrt())}
            out[f"K{K}_{dist}"] = row
    path = os.path.join(ROOT, "gpurun_out", os.environ.get("TAG", "r2") + "_accum_micro.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        print(k, {a: (f"{b['rel_l2']:.2e}", f"{b['mean_signed_over_rms']:+.2e}") for a, b in v.items()})


if __name__ == "__main__":
    main()
