"""The seam between the caller-side conv front-end (library convolutions, sgg_b200/frontend.py) and the CUDA hot path:
conv-variable gradients obtained by back-propagating the annotation adjoint that sgg_disc_step / sgg_gen_step return
must equal the oracle differentiated end to end (front-end restated in fp64 + oracle/sgg_oracle.py), and the reference's
loop on pixels must run through SceneGraphGAN.train_from_images."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-3     # conv gradients are linear images of the annotation adjoint (itself within 1e-3); measured 2e-5


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def test_conv_gradients_through_the_seam_match_the_oracle_end_to_end():
    from oracle import sgg_oracle as O
    from sgg_b200.frontend import ConvFrontEnd, FrontEndTrainer
    from sgg_b200.trainer import HotPathTrainer
    B, T, V, R, lam = 4, 2, 50, 4, 10.0
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        tr = HotPathTrainer(B, T, V, critic_iters=1, lam=lam, regions=R, use_graph=False)
        ft = FrontEndTrainer(tr, seed=3)
        eng = tr.eng
        gp = O.init_generator_params(V, seed=5, R=R, C=512, H=512)
        dp = O.init_discriminator_params(V, seed=6, R=R, C=512, H=512, E=300)
        dp["Discriminator/W"] = dp["Discriminator/W"] * 40          # slopes above 1: the penalty is active
        eng.g.load_state_dict(gp); eng.d.load_state_dict(dp)
        g = torch.Generator().manual_seed(9)
        images = torch.randn(B, 17, 17, 3, generator=g)              # 17 -> 9 -> 5 -> 3 -> 2: R = 4
        labels = torch.randint(0, V, (B, T), generator=g)
        noise, alpha = torch.randn(B, 512, generator=g), torch.rand(B, generator=g)
        real = torch.nn.functional.one_hot(labels, V).double()
        gp64, dp64 = {k: v.double() for k, v in gp.items()}, {k: v.double() for k, v in dp.items()}
        dev = eng.device
        # ---------------- GPU: front-end (cuDNN, fp32) -> hot path -> annotation adjoint -> front-end backward
        ann_g = ft.fg(images.to(dev)).reshape(B, R, 512)
        ann_d = ft.fd(images.to(dev)).reshape(B, R, 512)
        ag16, ad16 = ann_g.detach().to(torch.bfloat16).contiguous(), ann_d.detach().to(torch.bfloat16).contiguous()
        eng.set_batch(ag16, ad16, labels.to(dev))
        eng.noise.copy_(noise); eng.gp_alpha.copy_(alpha)
        got_d = torch.autograd.grad(ann_d, list(ft.fd.live_parameters()), grad_outputs=eng.disc_step(ann_grad=True))
        gp_value = float(eng.scalars[2])
        got_g = torch.autograd.grad(ann_g, list(ft.fg.live_parameters()), grad_outputs=eng.gen_step(ann_grad=True))
        torch.cuda.synchronize()
        # ---------------- CPU: the same front-ends in fp64, annotation VALUES as the GPU saw them (bf16), end-to-end autograd
        def twin(net):
            t = ConvFrontEnd(net.scope).double()
            t.load_tf_variables({k: v.cpu().double() for k, v in net.tf_variables().items()})
            return t
        fg64, fd64 = twin(ft.fg), twin(ft.fd)

        def straight_through(net64, seen16):
            a = net64(images.double()).reshape(B, R, 512)
            assert _rel(a, seen16.float()) < 1e-2               # the GPU front-end computed the same annotations (bf16-rounded)
            return a + (seen16.cpu().double() - a).detach()
        losses = O.wgan_gp_losses(gp64, dp64, ag16.cpu().double(), straight_through(fd64, ad16), real, noise.double(),
                                  alpha.double(), lam, T)
        ref_gp = losses["gp"].item()
        assert ref_gp > 0 and abs(gp_value - ref_gp) <= 1e-3 * ref_gp + 1e-5
        ref_d = torch.autograd.grad(losses["disc_cost"], list(fd64.live_parameters()))
        losses = O.wgan_gp_losses(gp64, dp64, straight_through(fg64, ag16), ad16.cpu().double(), real, noise.double(),
                                  alpha.double(), lam, T)
        ref_g = torch.autograd.grad(losses["gen_cost"], list(fg64.live_parameters()))
        worst = {"disc": max(_rel(a, b) for a, b in zip(got_d, ref_d)), "gen": max(_rel(a, b) for a, b in zip(got_g, ref_g))}
        os.makedirs("gpurun_out", exist_ok=True)
        with open(os.path.join("gpurun_out", "frontend_seam.json"), "w") as f:
            json.dump({"worst_relative_l2_of_conv_gradients": worst, "tolerance": TOL, "gp": gp_value}, f)
        assert worst["disc"] < TOL and worst["gen"] < TOL, worst
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32


def test_train_from_images_runs_the_reference_loop_on_pixels(tmp_path):
    """train:362-368 with the conv stacks in the loop at the reference's image size (221 x 221 -> 14 x 14 x 512): losses are
    finite, every live conv variable of both stacks moves, the dead ones (gen:59-62) do not, and the front-end travels
    through the checkpoint."""
    from sgg_b200.train import SceneGraphGAN
    B, V, nc = 2, 60, 2
    gan = SceneGraphGAN(str(tmp_path / "ck"), str(tmp_path / "logs"), None, None, None, None, None, critic_iters=nc, batch_size=B,
                        lambda_=10, resume=False, vocab_size=V)
    g = torch.Generator().manual_seed(1)
    batches = [(torch.randn(B, 221, 221, 3, generator=g), torch.randint(0, V, (B, 3), generator=g)) for _ in range(2)]
    f = gan._front()
    before = {k: v.clone() for net in (f.fg, f.fd) for k, v in net.tf_variables().items()}
    hot_before = gan.trainer.eng.d.theta.clone()
    logs = gan.train_from_images(batches)
    torch.cuda.synchronize()
    assert len(logs) == 2 and all(len(l["disc_cost"]) == nc for l in logs)
    for l in logs:
        assert all(map(lambda x: x == x and abs(x) < 1e6, l["disc_cost"] + l["gp"] + [l["gen_cost"]])), l
    after = {k: v for net in (f.fg, f.fd) for k, v in net.tf_variables().items()}
    dead = ("conv2d_10/", "conv2d_11/", "LayerNorm_10/", "LayerNorm_11/")
    for k, v in after.items():
        assert torch.equal(v, before[k]) == any(d in k for d in dead), k
    assert not torch.equal(gan.trainer.eng.d.theta, hot_before)
    assert f.adam_fd.t == 2 * nc and f.adam_fg.t == 2 and gan.trainer.eng.d.step == 2 * nc
    gan._saveModel()
    again = SceneGraphGAN(str(tmp_path / "ck"), str(tmp_path / "logs"), None, None, None, None, None, critic_iters=nc, batch_size=B,
                          lambda_=10, resume=True, vocab_size=V)
    assert again.front is not None and again.front.adam_fd.t == 2 * nc
    for net_a, net_b in ((f.fg, again.front.fg), (f.fd, again.front.fd)):
        for (k, a), b in zip(net_a.tf_variables().items(), net_b.tf_variables().values()):
            assert torch.equal(a, b), k
    for a, b in zip(f.adam_fd.m + f.adam_fd.v, again.front.adam_fd.m + again.front.adam_fd.v):
        assert torch.equal(a, b)
    gan.trainer.close(); again.trainer.close()


def test_training_from_a_dataset_directory_in_the_reference_formats(tmp_path):
    """vocab.json / ims_to_triples.json / image_means.txt / image_stds.txt / JPEG files (dataset_creation/*.py) ->
    SceneGraphGAN.train() -> R@k evaluation on the test split: the reference's own mode of operation (train.py:341-395)."""
    import random

    import numpy as np
    from PIL import Image

    from sgg_b200.train import SceneGraphGAN
    rng, nrng = random.Random(0), np.random.RandomState(0)
    vocab = {f"w{i}": i for i in range(30)}
    ims = {}
    for i in range(20):
        p = str(tmp_path / f"{i}.jpg")
        Image.fromarray(nrng.randint(0, 256, size=(48, 64, 3), dtype=np.uint8)).save(p)
        ims[p] = [[rng.randrange(30) for _ in range(3)] for _ in range(2 + i % 3)]
    (tmp_path / "vocab.json").write_text(json.dumps(vocab))
    (tmp_path / "ims.json").write_text(json.dumps(ims))
    (tmp_path / "means.txt").write_text("119.6\n115.1\n106.1\n")
    (tmp_path / "stds.txt").write_text("30.4\n30.5\n36.7\n")
    np.save(str(tmp_path / "emb.npy"), nrng.uniform(-0.1, 0.1, size=(30, 300)))
    gan = SceneGraphGAN(str(tmp_path / "ck"), str(tmp_path / "logs"), str(tmp_path / "ims.json"), str(tmp_path / "vocab.json"),
                        str(tmp_path / "emb.npy"), str(tmp_path / "means.txt"), str(tmp_path / "stds.txt"), critic_iters=1,
                        batch_size=2, lambda_=10, resume=False, allow_synthetic=False)
    assert gan.trainer.V == 30
    n = gan.train(max_iterations=2)
    assert n == 2 and gan.front is not None and gan.front.adam_fd.t == 2
    r50, r100 = gan.test_from_images(list(gan.dataset["test"])[:2], multiplier=2, out_path=str(tmp_path / "recalls.txt"))
    assert 0.0 <= r50 <= 1.0 and 0.0 <= r100 <= 1.0
    assert (tmp_path / "recalls.txt").read_text().count("\n") == 1
    gan.trainer.close()


@pytest.mark.parametrize("B,C,H,W", [(3, 32, 37, 41), (2, 512, 5, 7), (1, 4, 3, 3), (2, 64, 111, 111), (5, 128, 1, 1)])
def test_fused_layer_norm_elu_kernels(B, C, H, W):
    """sgg_ln_elu_forward / _backward (csrc/frontend.cu) against torch's group_norm + elu in fp64 and against the numpy
    restatement of tf.contrib.layers.layer_norm(activation_fn=elu) (gen:30): several chunks per sample, ragged tails,
    a single-pixel sample (variance of C values only), non-trivial gamma / beta, a mean far from zero."""
    import numpy as np

    from oracle import frontend_oracle as FO
    from sgg_b200._lib import kernel_counts
    from sgg_b200.frontend import layer_norm_elu
    g = torch.Generator().manual_seed(B * 1000 + C)
    x = (torch.randn(B, C, H, W, generator=g) * 1.7 + 3.0).cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    gamma = (1 + 0.3 * torch.randn(C, generator=g)).cuda().requires_grad_(True)
    beta = (0.3 * torch.randn(C, generator=g)).cuda().requires_grad_(True)
    dy = torch.randn(B, C, H, W, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    k0 = kernel_counts()
    y = layer_norm_elu(x, gamma, beta)
    gx, gg, gb = torch.autograd.grad(y, (x, gamma, beta), grad_outputs=dy)
    torch.cuda.synchronize()
    k1 = kernel_counts()
    for name in ("ln_elu_stats_kernel", "ln_elu_apply_kernel", "ln_elu_bwd_reduce_kernel", "ln_elu_bwd_apply_kernel"):
        assert k1.get(name, 0) - k0.get(name, 0) == 1, (name, k1)
    x64, g64, b64 = (t.detach().cpu().double().requires_grad_(True) for t in (x, gamma, beta))
    y64 = torch.nn.functional.elu(torch.nn.functional.group_norm(x64, 1, g64, b64, 1e-12))
    r = torch.autograd.grad(y64, (x64, g64, b64), grad_outputs=dy.cpu().double())
    assert _rel(y, y64) < 2e-6
    assert _rel(gx, r[0]) < 2e-5 and _rel(gg, r[1]) < 2e-5 and _rel(gb, r[2]) < 2e-5
    ref = FO.layer_norm_elu(x64.detach().permute(0, 2, 3, 1).numpy(), g64.detach().numpy(), b64.detach().numpy())
    assert np.abs(y.detach().permute(0, 2, 3, 1).cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()


def test_front_end_with_fused_norms_equals_the_library_norms():
    from sgg_b200.frontend import ConvFrontEnd
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        net = ConvFrontEnd(seed=4).cuda()
        images = torch.randn(2, 45, 61, 3, generator=torch.Generator().manual_seed(2)).cuda()
        outs, grads = [], []
        for fused in (None, False):
            net.fused_norm = fused
            a = net(images)
            outs.append(a.detach())
            grads.append(torch.autograd.grad(a.square().sum(), list(net.live_parameters())))
        assert _rel(outs[0], outs[1]) < 1e-5
        for a, b in zip(*grads):
            assert _rel(a, b) < 1e-4
    finally:
        torch.backends.cudnn.allow_tf32 = old
