#!/usr/bin/env python
"""One forward + reverse of the fused H*W*C layer norm + ELU kernels (csrc/frontend.cu) at the front-end's largest
activation (B 64, 221 x 221 x 32, 400 MB): the command the ncu capture profiles/r2_ln_elu_kernels.* wraps
(NCU_SCRIPT=tools/ln_elu_profile.py bash tools/gpu.sh full:ln_elu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    import __graft_entry__ as ge
    ge.build()
    from sgg_b200.frontend import layer_norm_elu
    x = torch.randn(64, 32, 221, 221, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    g = torch.ones(32, device="cuda", requires_grad=True)
    b = torch.zeros(32, device="cuda", requires_grad=True)
    dy = torch.randn_like(x)
    for _ in range(2):
        y = layer_norm_elu(x, g, b, True)
        torch.autograd.grad(y, (x, g, b), grad_outputs=dy)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
