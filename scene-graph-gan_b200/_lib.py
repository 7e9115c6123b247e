"""ctypes binding of csrc/libsgg_b200.so (the C ABI declared in include/sgg_b200.h).

The product has no CPU fallback: if the shared library is missing or a call fails, this
module raises.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libsgg_b200.so")
MAX_SEG = 4


class SggError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("a_rows", C.c_int64), ("a_cols", C.c_int64), ("a_ld", C.c_int64), ("a_mn_major", C.c_int32),
        ("B", C.c_void_p), ("b_rows", C.c_int64), ("b_cols", C.c_int64), ("b_ld", C.c_int64), ("b_mn_major", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32),
        ("nseg", C.c_int32),
        ("seg_a_k", C.c_int32 * MAX_SEG), ("seg_a_mn", C.c_int32 * MAX_SEG),
        ("seg_b_k", C.c_int32 * MAX_SEG), ("seg_b_mn", C.c_int32 * MAX_SEG),
        ("seg_klen", C.c_int32 * MAX_SEG),
        ("C", C.c_void_p), ("ldc", C.c_int64), ("atomic", C.c_int32),
        ("Chl", C.c_void_p), ("ld_hl", C.c_int64), ("lo_off", C.c_int64),
        ("bias", C.c_void_p),
        ("addm", C.c_void_p), ("ld_addm", C.c_int64), ("add_mod", C.c_int32),
        ("alpha", C.c_float),
        ("block_n", C.c_int32),
        ("splits", C.c_int32),
    ]


_lib = None


def lib() -> C.CDLL:
    """Loads libsgg_b200.so; raises SggError (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SggError(
                f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
                "(nvcc, sm_100a). The sgg_b200 hot path has no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.sgg_last_error.restype = C.c_char_p
        _lib.sgg_version.restype = C.c_int
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise SggError(f"{what} failed ({rc}): {lib().sgg_last_error().decode()}")


def stream_ptr(stream=None) -> C.c_void_p:
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)
