"""C-ABI boundary checks that need no GPU: the library loads, exports every symbol the header
declares, and the host-only entry points (parameter table, workspace size, argument validation)
behave as documented in include/sgg_b200.h."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from sgg_b200 import _lib
    return _lib.lib()


def _header_symbols():
    with open(os.path.join(ROOT, "include", "sgg_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sgg_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from sgg_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/sgg_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms, "sgg_b200._lib.EXPORTS out of sync with the header"


def test_version_and_error_string(lib):
    assert lib.sgg_version() >= 100
    assert isinstance(lib.sgg_last_error(), bytes)


def test_param_table_reproduces_tf_variable_names(lib):
    """train.py:262-263 splits variables by name prefix; gen:15,79,88 / disc:15,81,90 / train:70 name them."""
    from sgg_b200._lib import Dims, ParamEntry
    d = Dims(B=32, T=3, V=2000, R=196, C=512, H=512, E=300)
    for net, prefix, n_expected in ((0, "Generator/Generator", 15), (1, "Discriminator/Discriminator", 16)):
        n, nf, ns = C.c_int(0), C.c_int64(0), C.c_int64(0)
        assert lib.sgg_param_table(net, C.byref(d), None, 0, C.byref(n), C.byref(nf), C.byref(ns)) == 0
        assert n.value == n_expected
        ents = (ParamEntry * n.value)()
        assert lib.sgg_param_table(net, C.byref(d), ents, n.value, C.byref(n), C.byref(nf), C.byref(ns)) == 0
        by = {e.name.decode(): e for e in ents}
        att = by[f"{prefix}/attention_perceptron/kernel"]
        assert (att.rows, att.cols) == (196 * 512 + 512, 196)            # gen:14-15: concat[flat(a), c] -> 196
        U = 512 if net == 0 else 300
        k = by[f"{prefix}/layer_norm_basic_lstm_cell/kernel"]
        assert (k.rows, k.cols) == (512 + U + 512, 2048)                 # [z_hat | noise-or-embedding | h] x [i|j|f|o]
        dec = by[f"{prefix}/decoder/kernel"]
        assert (dec.rows, dec.cols) == (512, 2000 if net == 0 else 1)
        for ln in ("input", "transform", "forget", "output", "state"):
            assert f"{prefix}/layer_norm_basic_lstm_cell/{ln}/gamma" in by
            assert f"{prefix}/layer_norm_basic_lstm_cell/{ln}/beta" in by
        if net == 1:
            w = by["Discriminator/W"]
            assert (w.rows, w.cols) == (2000, 300)                       # train:70
        # no overlaps, everything inside the bucket
        spans = sorted((e.offset, e.offset + e.rows * e.cols) for e in ents)
        for (a0, a1), (b0, _) in zip(spans, spans[1:]):
            assert a1 <= b0
        assert spans[-1][1] <= nf.value
        n_params = sum(e.rows * e.cols for e in ents)
        # 23.95 M (G) / 23.09 M (D) hot-path parameters at V=2000
        assert abs(n_params - (23.95e6 if net == 0 else 23.09e6)) < 0.02e6


def test_workspace_bytes_and_dim_validation(lib):
    from sgg_b200._lib import Dims
    lib.sgg_workspace_bytes.restype = C.c_int64
    good = Dims(B=256, T=3, V=2000, R=196, C=512, H=512, E=300)
    n = lib.sgg_workspace_bytes(C.byref(good))
    assert 0 < n < 4 << 30
    bad = Dims(B=256, T=3, V=2000, R=196, C=256, H=512, E=300)
    assert lib.sgg_workspace_bytes(C.byref(bad)) < 0
    assert b"512" in lib.sgg_last_error()
    bad2 = Dims(B=0, T=3, V=2000, R=196, C=512, H=512, E=300)
    assert lib.sgg_workspace_bytes(C.byref(bad2)) < 0


def test_null_arguments_are_rejected_not_crashed(lib):
    assert lib.sgg_gemm(None, None) < 0
    assert lib.sgg_gen_forward(None, None) < 0
    assert lib.sgg_disc_step(None, None) < 0
    assert lib.sgg_gen_step(None, None) < 0
    assert lib.sgg_adam_step(0, None, None, None, None, None, None, C.c_int64(1), C.c_float(1e-4), C.c_float(0.5),
                             C.c_float(0.9), C.c_float(1e-8), C.c_float(1.0), None) < 0
    assert lib.sgg_last_error() != b""


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "sgg_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(import|from)\s+oracle|import_module\(.oracle|oracle/_ref", src, flags=re.M), \
                    f"{fn} reaches into oracle/"


def test_ctypes_mirrors_match_the_c_struct_layouts(tmp_path):
    """The Python side mirrors the public structs of include/sgg_b200.h with ctypes; a drifted field order or type
    would silently corrupt arguments.  Compile a probe with gcc that prints sizeof / offsetof and compare."""
    import ctypes as C
    import shutil
    import subprocess
    from sgg_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = {
        "sgg_dims_t": (_lib.Dims, ["B", "S"]),
        "sgg_gemm_desc_t": (_lib.GemmDesc, ["A", "seg_klen", "C", "Chl", "bias", "alpha", "splits", "out_s1", "argmax_keys",
                                            "gumbel_offset"]),
        "sgg_param_entry_t": (_lib.ParamEntry, ["name", "offset", "shadow_rows"]),
        "sgg_step_args_t": (_lib.StepArgs, ["dims", "lam", "g_theta", "labels", "workspace_bytes", "flags"]),
        "sgg_wa_shard_t": (_lib.WaShard, ["enabled", "slab_g", "scratch_bytes"]),
        "sgg_iter_args_t": (_lib.IterArgs, ["step", "critic_iters", "lr", "seed", "scalars_all", "comm", "shard"]),
        "sgg_sample_args_t": (_lib.SampleArgs, ["dims", "noise", "mode", "seed", "workspace_bytes", "logits_out"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "sgg_b200.h")}"',
             "int main(void) {"]
    for cname, (_, fields) in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = dict(l.rsplit(" ", 1) for l in out.strip().splitlines())
    for cname, (ctype, fields) in pairs.items():
        assert int(got[cname]) == C.sizeof(ctype), (cname, got[cname], C.sizeof(ctype))
        for f in fields:
            assert int(got[f"{cname}.{f}"]) == getattr(ctype, f).offset, (cname, f)
