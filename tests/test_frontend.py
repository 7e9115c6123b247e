"""The caller-side convolutional front-end (sgg_b200/frontend.py; library convolutions, not kernels of
this repo) against the numpy restatement of gen:29-68 in oracle/frontend_oracle.py, and the seam between the front-end and
the recurrent half: back-propagating the annotation adjoint through the stack equals end-to-end differentiation."""
import numpy as np
import pytest
import torch

from oracle import frontend_oracle as FO
from oracle import sgg_oracle as O
from sgg_b200.frontend import ConvFrontEnd, TFAdam, same_padding


def _randomised(scope, seed):
    net = ConvFrontEnd(scope, seed=seed).double()
    g = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():
        for p in list(net.gammas) + list(net.betas) + list(net.biases):
            p.add_(0.2 * torch.randn(p.shape, generator=g, dtype=torch.float64))
    return net


def test_same_padding_known_answers():
    """TensorFlow SAME: 221 -> 111 -> 56 -> 28 -> 14 with 5x5 stride-2 kernels (gen:35,50,65,68); the last two pad (1, 2)."""
    assert same_padding(221, 3, 1) == (1, 1) and same_padding(221, 5, 2) == (2, 2)
    assert same_padding(111, 5, 2) == (2, 2) and same_padding(56, 5, 2) == (1, 2) and same_padding(28, 5, 2) == (1, 2)
    n = 221
    for _ in range(4):
        n = -(-n // 2)
    assert n == 14


@pytest.mark.parametrize("hw", [(13, 13), (16, 11)])
def test_forward_matches_the_numpy_restatement(hw):
    net = _randomised("Generator/Generator", 1)
    g = torch.Generator().manual_seed(5)
    images = torch.randn(2, hw[0], hw[1], 3, generator=g, dtype=torch.float64)
    got = net(images)
    ref = FO.front_end({k: v.numpy() for k, v in net.tf_variables().items()}, "Generator/Generator", images.numpy())
    assert got.shape == ref.shape and got.shape[-1] == 512
    assert np.abs(got.detach().numpy() - ref).max() <= 1e-9 * np.abs(ref).max()


def test_full_size_image_gives_the_14x14x512_annotation_grid():
    net = ConvFrontEnd("Discriminator/Discriminator", seed=3)
    with torch.no_grad():
        out = net(torch.randn(1, 221, 221, 3))
    assert tuple(out.shape) == (1, 14, 14, 512) and torch.isfinite(out).all()


def test_variable_names_and_layouts_follow_the_reference_graph():
    net = ConvFrontEnd("Discriminator/Discriminator", seed=0)
    v = net.tf_variables()
    assert len(v) == 14 * 2 + 13 * 2
    assert tuple(v["Discriminator/Discriminator/conv2d/kernel"].shape) == (3, 3, 3, 32)            # HWIO, gen:29
    assert tuple(v["Discriminator/Discriminator/conv2d_2/kernel"].shape) == (5, 5, 32, 32)         # gen:35
    assert tuple(v["Discriminator/Discriminator/conv2d_12/kernel"].shape) == (5, 5, 256, 512)      # conv3_5 reads layernorm3_2
    assert tuple(v["Discriminator/Discriminator/conv2d_13/kernel"].shape) == (5, 5, 512, 512)      # self.downsampled
    assert tuple(v["Discriminator/Discriminator/LayerNorm_12/gamma"].shape) == (512,)
    assert "Discriminator/Discriminator/LayerNorm_13/gamma" not in v                               # gen:68 has no norm
    assert float(v["Discriminator/Discriminator/conv2d_5/bias"][0]) == pytest.approx(0.05)         # gen:21
    # he_normal: std sqrt(2 / fan_in), truncated at two (pre-scaling) standard deviations
    k = v["Discriminator/Discriminator/conv2d_9/kernel"]
    fan_in = 3 * 3 * 256
    assert float(k.std()) == pytest.approx((2.0 / fan_in) ** 0.5, rel=0.03)
    assert float(k.abs().max()) <= 2 * (2.0 / fan_in) ** 0.5 / 0.87962566103423978 + 1e-7
    # round trip through the TF names
    other = ConvFrontEnd("Discriminator/Discriminator", seed=9)
    used = other.load_tf_variables({k: t.clone() for k, t in v.items()})
    assert len(used) == len(v)
    for a, b in zip(net.parameters(), other.parameters()):
        assert torch.equal(a, b)


def test_dead_layers_get_no_gradient_and_adam_skips_them():
    """conv3_3 / conv3_4 (gen:59-62) feed nothing: tf.gradients gives None and AdamOptimizer leaves them alone."""
    net = ConvFrontEnd(seed=2)
    opt = TFAdam(net.parameters())
    before = [p.detach().clone() for p in net.parameters()]
    net(torch.randn(1, 9, 9, 3)).square().mean().backward()
    dead = {id(net.kernels[10]), id(net.biases[10]), id(net.kernels[11]), id(net.biases[11]),
            id(net.gammas[10]), id(net.betas[10]), id(net.gammas[11]), id(net.betas[11])}
    live = {id(p) for p in net.live_parameters()}
    assert not (dead & live) and len(dead) + len(live) == len(before)
    for p in net.parameters():
        assert (p.grad is None) == (id(p) in dead)
    opt.step()
    for p, b in zip(net.parameters(), before):
        assert torch.equal(p, b) == (id(p) in dead)
    # first Adam step moves every live weight by lr * g / (|g| + eps'): at most lr (+ fp32 rounding of the weight)
    for p, b in zip(net.parameters(), before):
        assert float((p.detach() - b).abs().max()) <= 1.001e-4


def test_annotation_adjoint_is_the_seam_between_front_end_and_hot_path():
    """d disc_cost / d conv variables by back-propagating the hot path's annotation adjoint (what sgg_disc_step returns in
    ann_d_grad, here from the oracle) through the stack == differentiating disc_cost end to end, gradient penalty included;
    likewise for the generator step."""
    B, T, V, R = 2, 2, 9, 4
    g = torch.Generator().manual_seed(7)
    images = torch.randn(B, 17, 17, 3, generator=g, dtype=torch.float64)        # 17 -> 9 -> 5 -> 3 -> 2: R = 4
    fg, fd = _randomised("Generator/Generator", 11), _randomised("Discriminator/Discriminator", 12)
    gp = {k: v.double() for k, v in O.init_generator_params(V, seed=1, R=R, C=512, H=512).items()}
    dp = {k: v.double() for k, v in O.init_discriminator_params(V, seed=2, R=R, C=512, H=512, E=8).items()}
    dp["Discriminator/W"] = dp["Discriminator/W"] * 300         # slopes above 1: the penalty is active
    labels = torch.randint(0, V, (B, T), generator=g)
    real = torch.nn.functional.one_hot(labels, V).double()
    noise = torch.randn(B, 512, generator=g, dtype=torch.float64)
    alpha = torch.rand(B, generator=g, dtype=torch.float64)
    # ---- D step: seam
    ann_g = fg(images).reshape(B, R, 512).detach()
    ann_d = fd(images).reshape(B, R, 512)
    step = O.disc_step_grads(gp, dp, ann_g, ann_d.detach(), real, noise, alpha, 10.0, T, ann_grad=True)
    assert float(step["gp"]) > 0
    seam = torch.autograd.grad(ann_d, list(fd.live_parameters()), grad_outputs=step["ann_grad"])
    # ---- D step: end to end
    losses = O.wgan_gp_losses(gp, dp, ann_g, fd(images).reshape(B, R, 512), real, noise, alpha, 10.0, T)
    e2e = torch.autograd.grad(losses["disc_cost"], list(fd.live_parameters()))
    for a, b in zip(seam, e2e):
        assert (a - b).norm() <= 1e-9 * b.norm() + 1e-14
    # ---- G step
    ann_g = fg(images).reshape(B, R, 512)
    step = O.gen_step_grads(gp, dp, ann_g.detach(), ann_d.detach(), noise, T, ann_grad=True)
    seam = torch.autograd.grad(ann_g, list(fg.live_parameters()), grad_outputs=step["ann_grad"])
    losses = O.wgan_gp_losses(gp, dp, fg(images).reshape(B, R, 512), ann_d.detach(), real, noise, alpha, 10.0, T)
    e2e = torch.autograd.grad(losses["gen_cost"], list(fg.live_parameters()))
    for a, b in zip(seam, e2e):
        assert (a - b).norm() <= 1e-9 * b.norm() + 1e-14


def test_conv_variables_travel_through_a_tensorflow_checkpoint(tmp_path):
    """A checkpoint in the TensorFlow V2 format with the reference graph's conv variables (HWIO kernels, their Adam slots,
    the hot-path variables next to them): the reader hands the conv variables to ConvFrontEnd under their TF names."""
    import numpy as np

    from sgg_b200 import tf_checkpoint as T
    src = _randomised("Generator/Generator", 21).float()
    tensors = {k: v.contiguous().numpy() for k, v in src.tf_variables().items()}
    tensors.update({k + "/Adam": np.zeros_like(v) for k, v in list(tensors.items())[:4]})
    tensors["Generator/Generator/decoder/bias"] = np.arange(7, dtype=np.float32)
    tensors["beta1_power"] = np.float32(0.5 ** 3)
    T.write_checkpoint(str(tmp_path / "model.ckpt"), tensors)
    parts = T.split_for_buckets(T.read_checkpoint(str(tmp_path / "model.ckpt")))
    assert parts["step"]["beta1_power"] == 3 and "Generator/Generator/decoder/bias" in parts["generator"]
    conv = {k: torch.from_numpy(v) for k, v in parts["other"].items()}
    assert "Generator/Generator/conv2d_13/kernel" in conv and conv["Generator/Generator/conv2d_13/kernel"].shape == (5, 5, 512, 512)
    dst = ConvFrontEnd("Generator/Generator", seed=0)
    used = dst.load_tf_variables(conv)
    assert len(used) == 54
    for a, b in zip(src.parameters(), dst.parameters()):
        assert torch.equal(a, b)
    with pytest.raises(ValueError):
        dst.load_tf_variables({"Generator/Generator/conv2d/kernel": torch.zeros(3, 3, 3, 31)}, strict=False)
    with pytest.raises(KeyError):
        dst.load_tf_variables({})


def test_golden_fixture_of_the_front_end():
    """tests/golden/frontend_17x13.json (written by tests/golden/make_golden_frontend.py from the numpy restatement): the
    restatement still produces it, and the torch module agrees with it."""
    import json
    import os

    from tests.golden import make_golden_frontend as G
    from tests.golden.make_golden import sample_index
    with open(os.path.join(os.path.dirname(G.__file__), "frontend_17x13.json")) as f:
        gold = json.load(f)
    again = G.compute()
    assert again["output_shape"] == gold["output_shape"] == [2, 2, 1, 512]
    assert again["downsampled"]["norm"] == pytest.approx(gold["downsampled"]["norm"], rel=1e-12)
    assert again["downsampled"]["samples"] == pytest.approx(gold["downsampled"]["samples"], rel=1e-10, abs=1e-12)
    cfg = gold["config"]
    net = _randomised(cfg["scope"], cfg["seed"])
    images = torch.randn(*cfg["shape"], generator=torch.Generator().manual_seed(cfg["image_seed"]), dtype=torch.float64)
    flat = net(images).detach().reshape(-1)
    idx = torch.from_numpy(sample_index("downsampled", flat.numel()))
    assert float(flat.norm()) == pytest.approx(gold["downsampled"]["norm"], rel=1e-9)
    assert flat[idx].tolist() == pytest.approx(gold["downsampled"]["samples"], rel=1e-8, abs=1e-10)
