"""tcgen05 GEMM (sgg_gemm) against torch fp32 matmul on the same bf16-exact operands.
Tolerance: operands are bf16-exact so products are exact in fp32; only the accumulation order
differs -> relative L2 <= 2e-6 (fp32 accumulate), stated per test."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _mk(rows, cols, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(rows, cols, generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16)


@pytest.mark.parametrize("bn", [64, 128, 256])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 200, 512), (768, 2048, 1344), (77, 196, 192)])
def test_gemm_layouts(bn, a_mn, b_mn, M, N, K):
    from sgg_b200 import ops
    A = _mk(M, K, 1)
    B = _mk(N, K, 2)
    ref = A.float() @ B.float().T
    Ain = A.T.contiguous() if a_mn else A
    Bin = B.T.contiguous() if b_mn else B
    # MN-major tensors need a 16-byte row pitch
    if a_mn and M % 8:
        pad = torch.zeros(K, (M + 7) // 8 * 8, dtype=torch.bfloat16, device="cuda"); pad[:, :M] = Ain; Ain = pad[:, :M]
    if b_mn and N % 8:
        pad = torch.zeros(K, (N + 7) // 8 * 8, dtype=torch.bfloat16, device="cuda"); pad[:, :N] = Bin; Bin = pad[:, :N]
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(Ain, Bin, M, N, a_mn=a_mn, b_mn=b_mn, segs=[(0, 0, 0, 0, K)], out=out, block_n=bn)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 2e-6


def test_gemm_hilo_segments_and_epilogue():
    """x (fp32, split hi|lo) times a bf16 TF-layout kernel [K,N], + bias + broadcast add, hi/lo output."""
    from sgg_b200 import ops
    M, K, N, KP = 384, 1324, 2048, 1344
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(M, K, generator=g, device="cuda")
    W = _mk(K, N, 4)
    bias = torch.randn(N, generator=g, device="cuda")
    addm = torch.randn(128, N, generator=g, device="cuda")
    xs = ops.split_hl(x, KP)
    out = torch.empty(M, N, device="cuda")
    NP = 2048
    out_hl = torch.zeros(M, 2 * NP, dtype=torch.bfloat16, device="cuda")
    ops.gemm(xs, W, M, N, b_mn=True, segs=[(0, 0, 0, 0, KP), (KP, 0, 0, 0, KP)], out=out, out_hl=out_hl,
             lo_off=NP, bias=bias, addm=addm, add_mod=128, alpha=0.5)
    torch.cuda.synchronize()
    ref = 0.5 * (x.double() @ W.double()) + bias.double() + addm.double().repeat(3, 1)
    assert _rel(out, ref) < 1e-5          # hi/lo split leaves 2^-18 relative operand error
    rec = out_hl[:, :N].float() + out_hl[:, NP:NP + N].float()
    assert _rel(rec, ref) < 1e-5


def test_gemm_weight_grad_three_products():
    """dW = x^T dy with both operands hi/lo split, contraction over rows (MN-major both)."""
    from sgg_b200 import ops
    rows, K, N, KP, NP = 1024, 1324, 2048, 1344, 2048
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(rows, K, generator=g, device="cuda")
    dy = torch.randn(rows, N, generator=g, device="cuda")
    xs, dys = ops.split_hl(x, KP), ops.split_hl(dy, NP)
    out = torch.zeros(K, N, device="cuda")
    segs = [(0, 0, 0, 0, rows), (0, 0, 0, NP, rows), (0, KP, 0, 0, rows)]
    ops.gemm(xs, dys, K, N, a_mn=True, b_mn=True, segs=segs, out=out, atomic=True, splits=4)
    torch.cuda.synchronize()
    ref = x.double().T @ dy.double()
    assert _rel(out, ref) < 1e-5


def test_gemm_attention_projection_shape():
    """K1 shape: P = flat(a) W_a with K = 196*512, TF-layout W_a [K,196], split-K with atomics."""
    from sgg_b200 import ops
    B, K, N = 256, 196 * 512, 196
    a = _mk(B, K, 6)
    Wa = (_mk(K, 200, 7) * 0.01).to(torch.bfloat16)[:, :N]        # row pitch 200 (16B multiple)
    out = torch.zeros(B, N, device="cuda")
    ops.gemm(a, Wa, B, N, b_mn=True, segs=[(0, 0, 0, 0, K)], out=out, atomic=True, splits=74, block_n=256)
    torch.cuda.synchronize()
    ref = a.float() @ Wa.float()
    assert _rel(out, ref) < 5e-6
