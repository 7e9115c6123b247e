#!/usr/bin/env python
"""What the annotation adjoint costs and what the caller-side conv front-end costs, on one GPU (CUDA events, not under a
profiler):
  (a) sgg_disc_step / sgg_gen_step at config 2's shape (B 256) without and with ann_d_grad / ann_g_grad, eager launches,
      rotating annotation tensors (> L2), and the kernels the adjoint adds (by name, from sgg_kernel_counts);
  (b) the fused H*W*C layer norm + ELU kernels (csrc/frontend.cu) against torch's group_norm + elu at three of the stack's
      activation shapes, with the fraction of the HBM peak on the algorithmic 3 / 5 passes;
  (c) SceneGraphGAN.train_from_images (library convolutions around the step-level C ABI) at B = $FE_BATCH (default 64)
      with this repo's norm kernels, with torch's, and with bf16-autocast convolutions.
Writes gpurun_out/<TAG>_frontend_micro.json."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def timed(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    import __graft_entry__ as ge
    ge.build()
    from sgg_b200._lib import kernel_counts
    from sgg_b200.engine import Engine
    out = {}
    B, T, V, R = 256, 3, 2000, 196
    eng = Engine(B, T, V, R)
    eng.g.init_reference(1); eng.d.init_reference(2)
    anns = [torch.randn(B, R, 512, device="cuda").bfloat16() for _ in range(4)]       # 4 x 2 x 51 MB > L2
    labels = torch.randint(0, V, (B, T), device="cuda")
    state = {"i": 0}

    def step(which, grad):
        i = state["i"]; state["i"] += 1
        eng.set_batch(anns[i % 4], anns[(i + 1) % 4], labels)
        eng.sample_noise(); eng.sample_gp_alpha()
        (eng.disc_step if which == "d" else eng.gen_step)(ann_grad=grad)

    for which in ("d", "g"):
        plain = timed(lambda: step(which, False), 20)
        with_grad = timed(lambda: step(which, True), 20)
        out[f"{which}_step_ms"] = {"plain": plain, "with_annotation_adjoint": with_grad, "extra": with_grad - plain}
        k0 = kernel_counts(); step(which, False); k1 = kernel_counts(); step(which, True); k2 = kernel_counts()
        per = lambda a, b: {k: b[k] - a.get(k, 0) for k in b if b[k] != a.get(k, 0)}
        p0, p1 = per(k0, k1), per(k1, k2)
        out[f"{which}_step_kernels_added_by_the_adjoint"] = {k: p1[k] - p0.get(k, 0) for k in p1 if p1[k] != p0.get(k, 0)}
    out["adjoint_bytes"] = {"output_fp32": B * R * 512 * 4, "W_a_shadow_hi_lo": 2 * R * 512 * 256 * 2}
    del eng, anns
    torch.cuda.empty_cache()
    # ---- (b) the H*W*C layer norm + ELU of gen:30: this repo's kernels against torch's group_norm + elu, forward + reverse
    from sgg_b200.frontend import layer_norm_elu
    peak = 6489.0
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f).get("hbm_gbs", peak))
    except Exception:
        pass
    out["layer_norm_elu"] = {"peak_GBps": peak}
    for shape in ((64, 32, 221, 221), (64, 128, 111, 111), (64, 256, 56, 56)):
        x = torch.randn(*shape, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
        gam = torch.ones(shape[1], device="cuda", requires_grad=True)
        bet = torch.zeros(shape[1], device="cuda", requires_grad=True)
        dy = torch.randn_like(x)
        nbytes = x.numel() * 4
        row = {"tensor_MB": nbytes / 1e6}
        for name, fused in (("kernels", True), ("torch", False)):
            fwd = timed(lambda: layer_norm_elu(x, gam, bet, fused), 10)
            y = layer_norm_elu(x, gam, bet, fused)
            bwd = timed(lambda: torch.autograd.grad(y, (x, gam, bet), grad_outputs=dy, retain_graph=True), 10)
            row[name] = {"forward_ms": fwd, "reverse_ms": bwd}
            if fused:   # algorithmic traffic: 3 (forward) / 5 (reverse) passes over the tensor
                row[name]["forward_frac_of_hbm"] = 3 * nbytes / (fwd * 1e-3) / 1e9 / peak
                row[name]["reverse_frac_of_hbm"] = 5 * nbytes / (bwd * 1e-3) / 1e9 / peak
            del y
        out["layer_norm_elu"]["x".join(map(str, shape))] = row
        del x, dy
        torch.cuda.empty_cache()
    # ---- (c) the loop on pixels
    from sgg_b200.train import SceneGraphGAN
    Bf, nc = int(os.environ.get("FE_BATCH", "64")), 5
    for name, dt, fused in (("fp32_convs_fused_norms", torch.float32, None), ("fp32_convs_torch_norms", torch.float32, False),
                            ("bf16_autocast_convs", torch.bfloat16, None)):
        with tempfile.TemporaryDirectory() as tmp:
            gan = SceneGraphGAN(tmp, tmp, None, None, None, None, None, critic_iters=nc, batch_size=Bf, lambda_=10, resume=False,
                                vocab_size=V)
            gan._front().compute_dtype = dt
            gan._front().fg.fused_norm = gan._front().fd.fused_norm = fused
            torch.cuda.reset_peak_memory_stats()
            images = torch.randn(Bf, 221, 221, 3, device="cuda")
            lb = torch.randint(0, V, (Bf, 3), device="cuda")
            ms = timed(lambda: gan.train_from_images([(images, lb)]), 3, warm=1)
            out[f"train_from_images_{name}"] = {"batch": Bf, "critic_iters": nc, "ms_per_iteration": ms, "images_per_s": Bf / ms * 1e3,
                                                "peak_memory_GB": torch.cuda.max_memory_allocated() / 2 ** 30}
            gan.trainer.close()
            del gan
            torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", os.environ.get("TAG", "r2") + "_frontend_micro.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
