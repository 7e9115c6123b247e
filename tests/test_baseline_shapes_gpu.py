"""Oracle parity AT the BASELINE.json configuration shapes, through the kernel variants the bench runs.

configs[0] (B 32, T 3, V 2000, R 196, n_critic 5): D step, G step and one whole sgg_train_iteration.
configs[1] (B 256, same shape): D step and G step against the fp64 oracle -- this is the shape that selects the wide-B
two-accumulator K1 kernel gemm_kernel<256, false, true, 2> and the 256-row adam_proj_kernel; both are asserted by name.
configs[2] shape (T 30, V 5000) at R 196 with a small batch.  Tolerance: per-tensor relative L2 <= 1e-3 (north_star).
Also: what the bf16 annotation input format costs against fp32 annotations (the reference's self.downsampled, gen:68).
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-3
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
K1_WIDE = "gemm_kernel<256, false, true, 2>"


def _scalar_close(got, ref, tol=TOL):
    return abs(float(got) - float(ref)) <= tol * (abs(float(ref)) + 1e-2)


def _ran(before, after, name):
    return after.get(name, 0) - before.get(name, 0)


def _d_and_g_step(B, T, V, R, seed=0, ann_bf16=True, lam=10.0):
    """Runs the D step and the G step on the CUDA path and in the oracle; returns per-tensor errors."""
    from oracle import sgg_oracle as O
    from sgg_b200._lib import kernel_counts
    from tests.util import make_engine, make_problem, rel
    prob = make_problem(B, T, V, R=R, seed=seed, dtype=torch.float64, ann_bf16=ann_bf16)
    eng = make_engine(prob, B, T, V, R=R, lam=lam)
    k0 = kernel_counts()
    ref = O.disc_step_grads(prob["gp"], prob["dp"], prob["ann_g"], prob["ann_d"], prob["real"], prob["noise"], prob["alpha"], lam, T,
                            ann_grad=True)
    ann_d_grad = eng.disc_step(ann_grad=True)
    torch.cuda.synchronize()
    sc = eng.scalars.cpu()
    err = {"w_disc": (float(sc[1]), float(ref["w_disc"])), "gp": (float(sc[2]), float(ref["gp"]))}
    terr = {"slopes": rel(eng.ws_view("slopes", (B,), torch.float32), ref["slopes"]),
            "d disc_cost / d ann_d": rel(ann_d_grad, ref["ann_grad"])}      # what disc:29-68 would back-propagate
    gv = eng.d.grad_views()
    for k, v in ref["grads"].items():
        if k == "Discriminator/Discriminator/decoder/bias":      # analytically zero
            assert abs(float(gv[k])) < 1e-5
            continue
        terr[k] = rel(gv[k], v)
    refg = O.gen_step_grads(prob["gp"], prob["dp"], prob["ann_g"], prob["ann_d"], prob["noise"], T, ann_grad=True)
    ann_g_grad = eng.gen_step(ann_grad=True)
    torch.cuda.synchronize()
    err["gen_cost"] = (float(eng.scalars[3]), float(refg["gen_cost"]))
    terr["d gen_cost / d ann_g"] = rel(ann_g_grad, refg["ann_grad"])
    gv = eng.g.grad_views()
    for k, v in refg["grads"].items():
        terr[k] = rel(gv[k], v)
    return err, terr, k0, kernel_counts()


def test_config1_d_step_and_g_step():
    """BASELINE configs[0] shape: batch 32, 196x512 annotations, 1 triple, vocab 2000."""
    err, terr, _, _ = _d_and_g_step(32, 3, 2000, 196)
    for k, (got, ref) in err.items():
        assert _scalar_close(got, ref), (k, got, ref)
    for k, e in terr.items():
        assert e < TOL, (k, e)


def test_config2_d_step_and_g_step_through_the_bench_kernels():
    """BASELINE configs[1] shape: batch 256.  The attention projection must have gone through the wide-B kernel."""
    err, terr, k0, k1 = _d_and_g_step(256, 3, 2000, 196, seed=1)
    assert _ran(k0, k1, K1_WIDE) >= 3, k1        # K1 of G (D step), D, G (G step) and D again
    for k, (got, ref) in err.items():
        assert _scalar_close(got, ref), (k, got, ref)
    for k, e in terr.items():
        assert e < TOL, (k, e)


def test_config3_shape_small_batch():
    """BASELINE configs[2] shape (10 triples = 30 timesteps, vocab 5000, 196 regions) at batch 4."""
    err, terr, _, _ = _d_and_g_step(4, 30, 5000, 196, seed=2)
    for k, (got, ref) in err.items():
        assert _scalar_close(got, ref), (k, got, ref)
    for k, e in terr.items():
        assert e < TOL, (k, e)


def test_config1_train_iteration_n_critic_5():
    """One sgg_train_iteration at configs[0] (5 critic steps + 1 generator step, Adam included, device RNG) vs the
    oracle's train_iteration fed with the same noise / alpha draws; the fused Adam + projection kernel must have run."""
    from oracle import sgg_oracle as O
    from sgg_b200._lib import kernel_counts
    from sgg_b200.engine import Engine
    from tests.util import make_problem
    B, T, V, R, nc, lam = 32, 3, 2000, 196, 5, 10.0
    prob = make_problem(B, T, V, R=R, seed=3, dtype=torch.float64)
    eng = Engine(B, T, V, R, lam=lam, critic_iters=nc, seed=11)
    eng.g.load_state_dict({k: v.float() for k, v in prob["gp"].items()})
    eng.d.load_state_dict({k: v.float() for k, v in prob["dp"].items()})
    eng.set_batch(prob["ann_g"].bfloat16().cuda().contiguous(), prob["ann_d"].bfloat16().cuda().contiguous(),
                  prob["labels"].cuda().contiguous())
    k0 = kernel_counts()
    eng.train_iteration()
    torch.cuda.synchronize()
    k1 = kernel_counts()
    assert _ran(k0, k1, "adam_proj_kernel") == nc, k1      # one per critic step (single GPU, B <= 256)
    noise, alpha = eng.noise_all.cpu(), eng.gp_alpha_all.cpu()
    gp = {k: v.clone().float() for k, v in prob["gp"].items()}
    dp = {k: v.clone().float() for k, v in prob["dp"].items()}
    ag, ad = O.TFAdam(gp), O.TFAdam(dp)
    log = O.train_iteration(gp, dp, ag, ad, prob["ann_g"].float(), prob["ann_d"].float(), prob["real"].float(),
                            [noise[i] for i in range(nc + 1)], [alpha[i] for i in range(nc)], lam, nc, T)
    sc = eng.scalars_all.cpu()
    for i in range(nc):
        cost = sc[i, 1].item() + lam * sc[i, 2].item()
        assert abs(cost - log["disc_cost"][i]) < 2e-3 * max(1.0, abs(log["disc_cost"][i])), (i, cost, log["disc_cost"][i])
    assert abs(sc[nc, 3].item() - log["gen_cost"]) < 2e-3
    for bucket, ref, ref0 in ((eng.g, gp, prob["gp"]), (eng.d, dp, prob["dp"])):
        for k, v in bucket.views().items():
            if k == "Discriminator/Discriminator/decoder/bias":
                continue   # analytically zero gradient: Adam's m/sqrt(v) turns rounding noise into O(lr) steps
            th, r, r0 = v.cpu().double(), ref[k].double(), ref0[k].double()
            assert ((th - r).norm() / (r.norm() + 1e-30)).item() < 1e-3, k
            upd = (r - r0).norm().item()
            if upd > 0 and not k.endswith("decoder/bias"):
                assert ((th - r).norm().item() / upd) < 0.1, k


def test_fused_adam_projection_at_config2_shape():
    """adam_proj_kernel at M = 256 rows AND R = 196 (the shape the bench runs) vs Adam-then-projection in fp64."""
    import ctypes as C
    import math
    from sgg_b200._lib import check, lib, stream_ptr
    from sgg_b200.params import DISC, ParamBucket, make_dims
    B, R = 256, 196
    dims = make_dims(B, 3, 50, R)
    g = torch.Generator().manual_seed(5)
    bk = ParamBucket(DISC, dims)
    bk.init_reference(3)
    name, off, rows, cols, soff, pitch = next(e for e in bk.entries if e[0].endswith("attention_perceptron/kernel"))
    n_wa = R * 512 * R
    bk.grad[off:off + n_wa] = (torch.randn(n_wa, generator=g) * 1e-3).cuda()
    bk.m[off:off + n_wa] = (torch.randn(n_wa, generator=g) * 1e-3).cuda()
    bk.v[off:off + n_wa] = (torch.rand(n_wa, generator=g) * 1e-6).cuda()
    ann = torch.randn(B, R * 512, generator=g).bfloat16().cuda()
    th0, m0, v0, gr = (x[off:off + n_wa].double().cpu() for x in (bk.theta, bk.m, bk.v, bk.grad))
    step, lr, b1, b2, eps = 4, 1e-4, 0.5, 0.9, 1e-8
    P = torch.full((B, 256), 7.0, device="cuda")
    check(lib().sgg_adam_project(C.c_int(DISC), C.byref(dims), C.c_void_p(bk.theta.data_ptr()), C.c_void_p(bk.grad.data_ptr()),
                                 C.c_void_p(bk.m.data_ptr()), C.c_void_p(bk.v.data_ptr()), C.c_void_p(bk.shadow.data_ptr()),
                                 C.c_int64(step), C.c_float(lr), C.c_float(b1), C.c_float(b2), C.c_float(eps),
                                 C.c_void_p(ann.data_ptr()), C.c_void_p(P.data_ptr()), C.c_int32(1), stream_ptr()), "adam_project")
    torch.cuda.synchronize()
    m1 = b1 * m0 + (1 - b1) * gr
    v1 = b2 * v0 + (1 - b2) * gr * gr
    th1 = th0 - lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step) * m1 / (v1.sqrt() + eps)
    assert torch.allclose(bk.theta[off:off + n_wa].double().cpu(), th1, rtol=1e-6, atol=1e-9)
    ref_P = ann.double().cpu() @ th1.view(R * 512, R)
    assert ((P[:, :R].double().cpu() - ref_P).norm() / ref_P.norm()).item() < 1e-4


def test_cost_of_the_bf16_annotation_format():
    """The reference's annotations (self.downsampled, gen:68) are fp32; the CUDA path takes them as bf16 (INTEGRATION.md
    section 4).  Here the oracle sees general fp32 N(0,1) annotations and the CUDA path their bf16 cast, at configs[0]
    shape.  Every annotation element then carries a relative rounding error of up to 2^-9, which a correct kernel
    cannot undo.  Measured on B200 (round 2, gpurun_out/fp32_annotation_cost.json, copied to profiles/): losses within
    1e-3, slopes 7e-4, generator gradients 4e-4 .. 3.8e-3, discriminator gradients 3e-3 .. 1.9e-2 (this problem has an
    active penalty of ~80, whose second-order terms amplify the input perturbation).  So the 1e-3 parity tolerance of
    north_star holds for bf16-representable annotations (every other parity test) and NOT for arbitrary fp32
    annotations: against those the input format costs up to 2e-2 per gradient tensor.  Asserted here: 5e-2."""
    err, terr, _, _ = _d_and_g_step(32, 3, 2000, 196, seed=4, ann_bf16=False)
    worst = max(terr.items(), key=lambda kv: kv[1])
    out = {"scalars": {k: {"got": g, "ref": r} for k, (g, r) in err.items()}, "per_tensor_rel_l2": terr,
           "worst": {"tensor": worst[0], "rel_l2": worst[1]}}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fp32_annotation_cost.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, (got, ref) in err.items():
        assert _scalar_close(got, ref, tol=1e-2), (k, got, ref)
    for k, e in terr.items():
        assert e < 5e-2, (k, e)


def test_label_ids_are_validated():
    """Labels index rows of Discriminator/W: ids outside [0, V) are rejected on the host side of the boundary."""
    from sgg_b200.engine import Engine
    from sgg_b200.trainer import HotPathTrainer
    B, T, V, R = 4, 3, 32, 8
    eng = Engine(B, T, V, R)
    ann = torch.zeros(B, R, 512, dtype=torch.bfloat16, device="cuda")
    bad = torch.zeros(B, T, dtype=torch.int64, device="cuda")
    bad[1, 2] = V
    with pytest.raises(ValueError):
        eng.set_batch(ann, ann, bad)
    bad[1, 2] = -1
    with pytest.raises(ValueError):
        eng.set_batch(ann, ann, bad)
    tr = HotPathTrainer(B, T, V, critic_iters=1, regions=R)
    with pytest.raises(ValueError):
        tr.upload(ann.cpu().pin_memory(), ann.cpu().pin_memory(), bad.cpu().pin_memory())
    # an unchecked bad id cannot corrupt memory: the kernels clamp it (result = the clamped id's result)
    eng.g.init_reference(1); eng.d.init_reference(2)
    a = torch.randn(B, R, 512).bfloat16().cuda()
    good = torch.randint(0, V, (B, T)).cuda()
    good[0, 0] = V - 1
    eng.set_batch(a, a, good)
    eng.disc_step(); torch.cuda.synchronize()
    want = eng.d.grad.clone()
    over = good.clone(); over[0, 0] = V + 5
    eng.set_batch(a, a, over, validate=False)
    eng.disc_step(); torch.cuda.synchronize()
    assert ((eng.d.grad - want).norm() / want.norm()).item() < 1e-4


def test_projection_cache_follows_the_weights():
    """The natural eager loop (gen_step; g.adam_step; gen_forward) must not reuse the projection P of the old weights:
    the engine compares the bucket's version counter instead of relying on the caller to flag the update."""
    from tests.util import make_engine, make_problem
    B, T, V, R = 6, 3, 40, 24
    prob = make_problem(B, T, V, R=R, seed=6, dtype=torch.float64)
    eng = make_engine(prob, B, T, V, R=R)
    eng.gen_step(); torch.cuda.synchronize()
    eng.g.grad.normal_(std=1.0)                    # a large step, so that a stale P would be visible
    eng.g.adam_step(lr=1e-2)
    got = eng.gen_forward().clone()
    fresh = make_engine(prob, B, T, V, R=R)
    fresh.g.load_state_dict(eng.g.state_dict())
    want = fresh.gen_forward()
    torch.cuda.synchronize()
    assert ((got - want).norm() / want.norm()).item() < 1e-5
    # ... and load_state_dict on a live engine
    eng.g.load_state_dict({k: v.float() for k, v in prob["gp"].items()})
    got = eng.gen_forward().clone()
    orig = make_engine(prob, B, T, V, R=R)
    want = orig.gen_forward()
    torch.cuda.synchronize()
    assert ((got - want).norm() / want.norm()).item() < 1e-5


@pytest.mark.parametrize("mode", [16, 8])
def test_fused_gate_epilogue_kernel(mode):
    """csrc/gates.cu (north_star item 2: gate nonlinearities + LayerNorm fused into the tcgen05 epilogue; gen:79,87):
    selected at run time, it must run (asserted by kernel name), match the oracle at config 2's shape within 1e-3, and
    agree with the default gate GEMM + cell kernel pair.  mode = cluster shape (16 CTAs x 32 units / 8 CTAs x 64 units)."""
    from sgg_b200._lib import kernel_counts, set_option
    name = "gates_fused_kernel<32>" if mode == 16 else "gates_fused_kernel<64>"
    try:
        set_option("fused_gates", mode)
        k0 = kernel_counts()
        err, terr, _, k1 = _d_and_g_step(256, 3, 2000, 196, seed=7)
        assert _ran(k0, k1, name) >= 9, k1            # 3 timesteps x (G forward, D forward, G step's two forwards)
        for k, (got, ref) in err.items():
            assert _scalar_close(got, ref), (k, got, ref)
        for k, e in terr.items():
            assert e < TOL, (k, e)
        # ragged m-tile (rows not a multiple of 128) and T = 5
        err, terr, _, _ = _d_and_g_step(5, 5, 130, 196, seed=8)
        for k, e in terr.items():
            assert e < TOL, (k, e)
    finally:
        set_option("fused_gates", 0)


def test_persistent_attention_kernels():
    """The persistent attention kernels (selected by default above two tiles per SM, i.e. B > 296 on B200) forced on at
    small shapes: the op-level forward / reverse entry points against torch references (several samples per CTA is
    exercised by B = 333 > SM count), and a D + G step through them against the oracle."""
    import ctypes as C
    from sgg_b200._lib import check, kernel_counts, lib, set_option, stream_ptr
    try:
        set_option("attn_persistent", 1)
        k0 = kernel_counts()
        for B, R, nv in ((333, 196, 3), (5, 100, 2), (150, 29, 4)):
            g = torch.Generator().manual_seed(B)
            a = torch.randn(B, R, 512, generator=g).bfloat16()
            e = torch.randn(nv * B, 256, generator=g) * 3
            a_d, e_d = a.cuda(), e.cuda()
            alpha = torch.empty_like(e_d)
            z = torch.zeros(nv * B, 1024, dtype=torch.bfloat16, device="cuda")
            check(lib().sgg_attn_forward(C.c_void_p(a_d.data_ptr()), C.c_int32(B), C.c_int32(R), C.c_int32(nv), C.c_void_p(e_d.data_ptr()),
                                         C.c_void_p(alpha.data_ptr()), C.c_int64(256), C.c_void_p(z.data_ptr()), C.c_int64(1024),
                                         C.c_int64(512), stream_ptr()), "attn")
            torch.cuda.synchronize()
            ref_al = torch.softmax(e[:, :R].double(), -1)
            ref_z = torch.einsum("sbr,brc->sbc", ref_al.reshape(nv, B, R), a.double()).reshape(nv * B, 512)
            got_z = z[:, :512].float().double() + z[:, 512:].float().double()
            assert (alpha[:, :R].cpu().double() - ref_al).abs().max().item() < 1e-6
            assert ((got_z.cpu() - ref_z).norm() / ref_z.norm()).item() < 1e-5
            # reverse
            ed = (torch.randn(nv * B, R, generator=g, dtype=torch.float64) * 2).requires_grad_(True)
            zb = torch.randn(nv * B, 512, generator=g, dtype=torch.float64)
            al = torch.softmax(ed, -1)
            (torch.einsum("sbr,brc->sbc", al.reshape(nv, B, R), a.double()).reshape(nv * B, 512) * zb).sum().backward()
            alpha_d = torch.zeros(nv * B, 256, device="cuda"); alpha_d[:, :R] = al.detach().float().cuda()
            zb_d = zb.float().cuda().contiguous()
            eb = torch.zeros(nv * B, 512, dtype=torch.bfloat16, device="cuda")
            pb = torch.zeros(B, 256, device="cuda")
            check(lib().sgg_attn_reverse(C.c_void_p(a_d.data_ptr()), C.c_int32(B), C.c_int32(R), C.c_int32(nv), C.c_void_p(zb_d.data_ptr()),
                                         C.c_int64(512), C.c_void_p(alpha_d.data_ptr()), C.c_int64(256), C.c_void_p(eb.data_ptr()),
                                         C.c_int64(512), C.c_int64(256), C.c_void_p(pb.data_ptr()), C.c_int64(256), stream_ptr()), "attn_rev")
            torch.cuda.synchronize()
            got = eb[:, :R].float().double().cpu() + eb[:, 256:256 + R].float().double().cpu()
            assert ((got - ed.grad).norm() / ed.grad.norm()).item() < 1e-4
            ref_pb = ed.grad.reshape(nv, B, R).sum(0)
            assert ((pb[:, :R].double().cpu() - ref_pb).norm() / ref_pb.norm()).item() < 1e-4
        err, terr, _, k1 = _d_and_g_step(6, 3, 96, 196, seed=9)
        assert _ran(k0, k1, "attn_fwd_p_kernel<0>") >= 3 and _ran(k0, k1, "attn_fwd_p_kernel<1>") >= 1 and _ran(k0, k1, "attn_rev_p_kernel") >= 3, k1
        for k, (got, ref) in err.items():
            assert _scalar_close(got, ref), (k, got, ref)
        for k, e2 in terr.items():
            assert e2 < TOL, (k, e2)
    finally:
        set_option("attn_persistent", -1)
