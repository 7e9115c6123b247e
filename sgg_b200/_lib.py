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
        ("out_d0", C.c_int32), ("out_d1", C.c_int32), ("out_s0", C.c_int64), ("out_s1", C.c_int64),
        ("argmax_keys", C.c_void_p), ("argmax_stride", C.c_int64),
        ("gumbel", C.c_int32), ("gumbel_seed", C.c_uint64), ("gumbel_offset", C.c_uint64),
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


def set_option(name: str, value: int) -> None:
    check(lib().sgg_set_option(name.encode(), C.c_int32(int(value))), "sgg_set_option")


def kernel_counts(reset: bool = False) -> dict:
    """{kernel name: launches} of every kernel this library has launched or captured so far (sgg_kernel_counts)."""
    L = lib()
    L.sgg_kernel_counts.restype = C.c_int64
    buf = C.create_string_buffer(1 << 16)
    L.sgg_kernel_counts(buf, C.c_int64(len(buf)), C.c_int32(int(reset)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n = line.rsplit(";", 1)
        name = name.replace("void ", "").replace("sgg::", "").split("(")[0]
        out[name] = out.get(name, 0) + int(n)
    return out


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise SggError(f"{what} failed ({rc}): {lib().sgg_last_error().decode()}")


def stream_ptr(stream=None) -> C.c_void_p:
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "T", "V", "R", "C", "H", "E", "S")]


class ParamEntry(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("offset", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32),
                ("shadow_offset", C.c_int64), ("shadow_pitch", C.c_int32), ("shadow_rows", C.c_int32)]


class StepArgs(C.Structure):
    _fields_ = [
        ("dims", Dims), ("world", C.c_int32), ("lam", C.c_float),
        ("g_theta", C.c_void_p), ("g_shadow", C.c_void_p), ("g_grad", C.c_void_p),
        ("d_theta", C.c_void_p), ("d_shadow", C.c_void_p), ("d_grad", C.c_void_p),
        ("ann_g", C.c_void_p), ("ann_d", C.c_void_p), ("labels", C.c_void_p),
        ("noise", C.c_void_p), ("gp_alpha", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("scalars", C.c_void_p), ("logits_out", C.c_void_p), ("flags", C.c_int32),
        ("ann_g_grad", C.c_void_p), ("ann_d_grad", C.c_void_p),
    ]


class WaShard(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("slab_g", C.c_void_p), ("slab_d", C.c_void_p),
                ("scratch", C.c_void_p), ("scratch_bytes", C.c_int64)]


class IterArgs(C.Structure):
    _fields_ = [
        ("step", StepArgs), ("critic_iters", C.c_int32),
        ("g_m", C.c_void_p), ("g_v", C.c_void_p), ("d_m", C.c_void_p), ("d_v", C.c_void_p),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("seed", C.c_uint64), ("counters", C.c_void_p),
        ("noise_all", C.c_void_p), ("gp_alpha_all", C.c_void_p), ("scalars_all", C.c_void_p),
        ("comm", C.c_void_p), ("shard", WaShard),
    ]


class SampleArgs(C.Structure):
    _fields_ = [
        ("dims", Dims),
        ("g_theta", C.c_void_p), ("g_shadow", C.c_void_p), ("ann_g", C.c_void_p), ("noise", C.c_void_p),
        ("mode", C.c_int32), ("chunk", C.c_int32), ("seed", C.c_uint64), ("offset", C.c_uint64),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("tokens_out", C.c_void_p), ("logits_out", C.c_void_p),
    ]


FLAG_REFRESH_GEN_PROJ = 1
SAMPLE_GREEDY, SAMPLE_GUMBEL = 0, 1

# every symbol include/sgg_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "sgg_last_error", "sgg_version", "sgg_launch_count", "sgg_set_option", "sgg_kernel_counts", "sgg_timing_report", "sgg_gemm", "sgg_attn_forward", "sgg_attn_reverse", "sgg_param_table", "sgg_refresh_shadow", "sgg_adam_step", "sgg_adam_project",
    "sgg_rng_fill_normal", "sgg_rng_fill_uniform", "sgg_workspace_bytes", "sgg_gen_forward",
    "sgg_disc_forward", "sgg_disc_step", "sgg_gen_step", "sgg_ws_lookup", "sgg_train_iteration",
    "sgg_comm_unique_id", "sgg_comm_init", "sgg_comm_destroy", "sgg_comm_allreduce_sum",
    "sgg_sample_workspace_bytes", "sgg_gen_sample", "sgg_wa_shard_scratch_bytes", "sgg_wa_shard_slab_elems",
    "sgg_ln_elu_scratch_floats", "sgg_ln_elu_forward", "sgg_ln_elu_backward", "sgg_gemm_plan",
]
