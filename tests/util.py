"""Shared helpers of the parity tests."""
from __future__ import annotations

import torch

from oracle import sgg_oracle as O


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    """Per-tensor relative L2 error ||a-b|| / ||b|| (the metric of SURVEY section 7 hard part 6)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def make_problem(B, T, V, R=196, E=300, seed=0, dtype=torch.float64, bf16_exact=False, ann_bf16=True):
    """Oracle parameters + synthetic batch.  Weights are general fp32 values (the CUDA path carries every
    GEMM operand as a bf16 hi/lo pair, so it sees the same weights to 2^-17); the annotations are
    bf16-representable because bf16 IS the input format of the CUDA path.  bf16_exact additionally
    rounds the matrix weights to bf16.  ann_bf16=False keeps general fp32 annotations (the reference's
    self.downsampled is fp32, gen:68): used to measure what the bf16 input format costs."""
    gp = O.init_generator_params(V, seed=seed + 3, R=R, C=512, H=512, dtype=torch.float32)
    dp = O.init_discriminator_params(V, seed=seed + 4, R=R, C=512, H=512, E=E, dtype=torch.float32)
    g = torch.Generator().manual_seed(seed + 7)
    for p in (gp, dp):
        for k in p:
            if "gamma" in k:
                p[k] = 1 + 0.2 * torch.randn(p[k].shape, generator=g)
            elif "beta" in k or "bias" in k:
                p[k] = 0.2 * torch.randn(p[k].shape, generator=g)
            elif bf16_exact:
                p[k] = p[k].bfloat16().float()
            p[k] = p[k].to(dtype)
    ann_g, ann_d, labels, real = O.synthetic_batch(B, V, T, R, 512, seed=seed + 5, dtype=dtype, bf16_exact=ann_bf16)
    noise = torch.randn(B, 512, generator=g).to(dtype)
    alpha = torch.rand(B, generator=g).to(dtype)
    return {"gp": gp, "dp": dp, "ann_g": ann_g, "ann_d": ann_d, "labels": labels, "real": real,
            "noise": noise, "alpha": alpha}


def make_engine(prob, B, T, V, R=196, E=300, lam=10.0):
    from sgg_b200.engine import Engine
    eng = Engine(B, T, V, R, E, lam=lam)
    eng.g.load_state_dict({k: v.float() for k, v in prob["gp"].items()})
    eng.d.load_state_dict({k: v.float() for k, v in prob["dp"].items()})
    eng.set_batch(prob["ann_g"].to(torch.bfloat16).cuda().contiguous(),
                  prob["ann_d"].to(torch.bfloat16).cuda().contiguous(), prob["labels"].cuda().contiguous())
    eng.noise.copy_(prob["noise"].float())
    eng.gp_alpha.copy_(prob["alpha"].float())
    return eng
