"""sgg_gemm_plan (host-side query of the tile / split-K choice, no GPU needed): every GEMM class of the training step and of
sampling, at the BASELINE shapes, keeps each tensor-core accumulator below GEMM_MAX_CHAIN_KB = 128 k-blocks whenever the
library picks the split -- the accumulator truncates, so this is an accuracy invariant, not a tuning detail."""
import ctypes as C

import pytest

from sgg_b200._lib import GemmDesc, lib

R, CC, H, E, RP, KXD, KXG, EP = 196, 512, 512, 300, 256, 1344, 1536, 320
DUMMY = 1 << 20           # never dereferenced by the query


def plan(M, N, K, nseg, a_mn=0, b_mn=0, atomic=0, hl_out=False, splits=0, block_n=0):
    d = GemmDesc()
    d.A, d.B = DUMMY, DUMMY
    d.a_rows, d.a_cols, d.a_ld, d.a_mn_major = (K, M, M, 1) if a_mn else (M, K, K, 0)
    d.b_rows, d.b_cols, d.b_ld, d.b_mn_major = (K, N, N, 1) if b_mn else (N, K, K, 0)
    d.M, d.N, d.nseg = M, N, nseg
    for s in range(nseg):
        d.seg_klen[s] = K
    if nseg >= 2:
        d.seg_b_k[1] = 1 << 18            # segment 1: another B part (hi/lo weight or hi/lo adjoint)
    if nseg >= 3:
        d.seg_a_k[2] = 1 << 18            # segment 2: another A part
    if hl_out:
        d.Chl, d.ld_hl, d.lo_off = DUMMY, 2 * N, N
    else:
        d.C, d.ldc = DUMMY, N
    d.atomic, d.alpha, d.splits, d.block_n = atomic, 1.0, splits, block_n
    bn, sp, kb = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    rc = lib().sgg_gemm_plan(C.byref(d), C.byref(bn), C.byref(sp), C.byref(kb))
    assert rc == 0, lib().sgg_last_error().decode()
    return bn.value, sp.value, kb.value


def step_gemms(B, T, V):
    VP = (V + 63) // 64 * 64
    rowsT = T * 4 * B
    return {
        "K1  P = flat(a) W_a": dict(M=B, N=R, K=R * CC, nseg=2, b_mn=1),
        "scores  e = c W_h (3 streams)": dict(M=3 * B, N=R, K=H, nseg=3, b_mn=1, atomic=2),
        "gates, D forward (3 streams)": dict(M=3 * B, N=4 * H, K=KXD, nseg=3, b_mn=1, atomic=2),
        "gates, G forward (5 draws)": dict(M=5 * B, N=4 * H, K=KXG, nseg=3, b_mn=1, atomic=2),
        "x_bar = q_bar K^T (4 blocks)": dict(M=4 * B, N=CC + E + H, K=4 * H, nseg=3, atomic=2),
        "c_bar += e_bar W_h^T": dict(M=4 * B, N=H, K=RP, nseg=3, atomic=1),
        "dK = X^T QB": dict(M=CC + E + H, N=4 * H, K=rowsT, nseg=3, a_mn=1, b_mn=1, atomic=1),
        "dW_h = C^T EB": dict(M=H, N=R, K=rowsT, nseg=3, a_mn=1, b_mn=1, atomic=1),
        "dW_a = flat(a)^T P_bar": dict(M=R * CC, N=R, K=B, nseg=2, a_mn=1, b_mn=1, splits=1),
        "logits (hi/lo output)": dict(M=T * 5 * B, N=V, K=H, nseg=3, b_mn=1, hl_out=True),
        "u = x W_emb": dict(M=T * B, N=E, K=VP, nseg=3, b_mn=1, atomic=2),
        "g = u_bar W_emb^T": dict(M=T * B, N=V, K=EP, nseg=3),
        "dW_emb = x^T u_bar": dict(M=V, N=E, K=T * B, nseg=3, a_mn=1, b_mn=1, atomic=1),
    }


@pytest.mark.parametrize("B,T,V", [(32, 3, 2000), (256, 3, 2000), (256, 30, 5000), (4096, 3, 2000)])
def test_no_accumulator_runs_over_more_than_128_k_blocks(B, T, V):
    for name, kw in step_gemms(B, T, V).items():
        bn, sp, kb = plan(**kw)
        assert bn in (64, 128, 256) and sp >= 1, name
        if kw.get("splits", 0) == 0 and not kw.get("hl_out"):      # the library chose the split
            assert kb <= 128, (name, bn, sp, kb)


def test_known_plans():
    assert plan(M=256, N=R, K=R * CC, nseg=2, b_mn=1) == (256, 148, 11)        # K1 at the bench shape: one CTA per SM
    assert plan(M=8192, N=R, K=R * CC, nseg=2, b_mn=1)[1:] == (13, 121)        # sampling batch: 4 splits before the cap
    bn, sp, kb = plan(M=32, N=R, K=R * CC, nseg=2, b_mn=1)                     # configs[0]: 8 splits before the cap
    assert sp >= 13 and kb <= 128
    assert plan(M=3 * 256, N=4 * H, K=KXD, nseg=3, b_mn=1, atomic=2) == (256, 3, 7)   # the (6, 8, 3) grid of profiles/r2_launches.md
    assert plan(M=1024, N=2048, K=512, nseg=3, b_mn=1, hl_out=True)[1] == 1    # hi/lo outputs cannot be split
    # an explicit split is the caller's business
    assert plan(M=128, N=256, K=1 << 17, nseg=1, splits=1)[1:] == (1, 2048)


def test_unfusable_segment_lists_have_no_single_plan():
    d = GemmDesc()
    d.A = d.B = d.C = DUMMY
    d.a_rows, d.a_cols, d.a_ld, d.b_rows, d.b_cols, d.b_ld = 128, 512, 512, 128, 512, 512
    d.M, d.N, d.nseg, d.ldc, d.alpha = 128, 128, 2, 128, 1.0
    d.seg_klen[0], d.seg_klen[1] = 256, 128                                    # different lengths: two launches
    bn, sp, kb = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    assert lib().sgg_gemm_plan(C.byref(d), C.byref(bn), C.byref(sp), C.byref(kb)) != 0
    assert b"no single plan" in lib().sgg_last_error()
