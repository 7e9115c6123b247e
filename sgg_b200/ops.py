"""Op-level Python wrappers over the C ABI (used by the engine and by the parity tests)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import GemmDesc, check, lib, stream_ptr


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def gemm(A: torch.Tensor, B: torch.Tensor, M: int, N: int, *, a_mn: bool = False, b_mn: bool = False,
         segs: Sequence[Tuple[int, int, int, int, int]] = (), out: Optional[torch.Tensor] = None,
         atomic: bool = False, out_hl: Optional[torch.Tensor] = None, lo_off: int = 0,
         bias: Optional[torch.Tensor] = None, addm: Optional[torch.Tensor] = None, add_mod: int = 0,
         alpha: float = 1.0, block_n: int = 0, splits: int = 1, argmax_keys: Optional[torch.Tensor] = None,
         argmax_stride: int = 1, gumbel: bool = False, gumbel_seed: int = 0, gumbel_offset: int = 0, stream=None) -> None:
    """D = alpha * sum_seg A_seg B_seg^T (+bias)(+addm[m % add_mod]).  A, B: 2-D bf16, last dim
    contiguous.  segs: (a_k, a_mn, b_k, b_mn, klen) per segment."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.dim() == 2 and B.dim() == 2
    assert A.stride(1) == 1 and B.stride(1) == 1
    d = GemmDesc()
    d.A, d.a_rows, d.a_cols, d.a_ld, d.a_mn_major = A.data_ptr(), A.shape[0], A.shape[1], A.stride(0), int(a_mn)
    d.B, d.b_rows, d.b_cols, d.b_ld, d.b_mn_major = B.data_ptr(), B.shape[0], B.shape[1], B.stride(0), int(b_mn)
    d.M, d.N = M, N
    d.nseg = len(segs)
    for i, (ak, am, bk, bm, kl) in enumerate(segs):
        d.seg_a_k[i], d.seg_a_mn[i], d.seg_b_k[i], d.seg_b_mn[i], d.seg_klen[i] = ak, am, bk, bm, kl
    if out is not None:
        assert out.dtype == torch.float32 and out.stride(1) == 1
        d.C, d.ldc = out.data_ptr(), out.stride(0)
    d.atomic = int(atomic)
    if out_hl is not None:
        assert out_hl.dtype == torch.bfloat16 and out_hl.stride(1) == 1
        d.Chl, d.ld_hl, d.lo_off = out_hl.data_ptr(), out_hl.stride(0), lo_off
    if bias is not None:
        d.bias = bias.data_ptr()
    if addm is not None:
        d.addm, d.ld_addm, d.add_mod = addm.data_ptr(), addm.stride(0), add_mod
    d.alpha = alpha
    d.block_n = block_n
    d.splits = splits
    if argmax_keys is not None:   # sampling epilogue: uint64 keys viewed as int64, zero-filled by the caller
        assert argmax_keys.dtype == torch.int64 and argmax_keys.is_contiguous()
        d.argmax_keys, d.argmax_stride = argmax_keys.data_ptr(), argmax_stride
        d.gumbel, d.gumbel_seed, d.gumbel_offset = int(gumbel), gumbel_seed, gumbel_offset
    check(lib().sgg_gemm(C.byref(d), stream_ptr(stream)), "sgg_gemm")


def split_hl(x: torch.Tensor, kpad: int) -> torch.Tensor:
    """Host-side helper (tests): fp32 [M,K] -> bf16 [M, 2*kpad] = [hi | 0 | lo | 0]."""
    M, K = x.shape
    out = torch.zeros(M, 2 * kpad, dtype=torch.bfloat16, device=x.device)
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    out[:, :K] = hi
    out[:, kpad:kpad + K] = lo
    return out


def decode_argmax_keys(keys: torch.Tensor) -> torch.Tensor:
    """Column index held in the low half of the sampling epilogue's (value, ~column) keys."""
    return (0xFFFFFFFF - (keys & 0xFFFFFFFF)).to(torch.int64)
