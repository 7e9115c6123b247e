"""CPU oracle for the Scene-Graph-GAN training hot path.  TEST INFRASTRUCTURE ONLY.

This file is a literal PyTorch-CPU restatement of the reference arithmetic.  It is
NOT part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
path (``sgg_b200``) never imports anything from ``oracle/`` and fails loudly
when its CUDA library is missing.

PARITY UNPINNED: the reference (/root/reference, Python 2 + TensorFlow 1.x ``tf.contrib``)
cannot be executed in this image (no python2, no tensorflow, no network) and it ships no
tests, golden vectors or known-answer fixtures.  The arithmetic lives in
un-vendored, un-pinned TensorFlow 1.x (feature use implies ~1.9-1.12); the semantics of
each TF symbol are restated from its published algorithm and anchored on the
reference's own call sites, cited below as ``gen:`` =
architectures/generator_with_attention.py, ``disc:`` =
architectures/discriminator_with_attention.py, ``train:`` = train.py.

The restatement is deliberately un-optimised: attention uses the concat form and is
re-evaluated at every timestep exactly as TF executes gen:14-15; the gradient penalty uses
autograd with ``create_graph=True`` (the analogue of tf.gradients-of-tf.gradients).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor

LSTM_UNITS = 512        # gen:79 / disc:81  LayerNormBasicLSTMCell(512)
EMBED_DIM = 300         # train:63,70 word_embeddings.npy is [V, 300]
LN_EPS = 1e-12          # tf.contrib.layers.layer_norm variance_epsilon
FORGET_BIAS = 1.0       # LayerNormBasicLSTMCell default forget_bias
GP_EPS = 1e-10          # tfgan wasserstein_gradient_penalty epsilon
GP_TARGET = 1.0         # tfgan wasserstein_gradient_penalty target
LN_NAMES = ("input", "transform", "forget", "output", "state")


# --------------------------------------------------------------------------------------
# Parameter construction (TF variable names; gen:15,79,88  disc:15,81,90  train:70)
# --------------------------------------------------------------------------------------
def _glorot_uniform(shape, gen: torch.Generator, dtype) -> Tensor:
    """tf.layers.dense / LayerNormBasicLSTMCell default kernel init (glorot_uniform)."""
    fan_in, fan_out = shape
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2.0 - 1.0).mul_(limit).to(dtype)


def _network_params(prefix: str, in_extra: int, out_units: int, R: int, C: int, H: int,
                    gen: torch.Generator, dtype) -> "OrderedDict[str, Tensor]":
    p: "OrderedDict[str, Tensor]" = OrderedDict()
    # gen:15  dense(units = 14*14) on concat[flat(a) (R*C), c (H)]
    p[f"{prefix}/attention_perceptron/kernel"] = _glorot_uniform((R * C + H, R), gen, dtype)
    p[f"{prefix}/attention_perceptron/bias"] = torch.zeros(R, dtype=dtype)
    # gen:79  LayerNormBasicLSTMCell: bias-free kernel [in + H, 4H], rows [inputs | h]
    p[f"{prefix}/layer_norm_basic_lstm_cell/kernel"] = _glorot_uniform((C + in_extra + H, 4 * H), gen, dtype)
    for n in LN_NAMES:
        p[f"{prefix}/layer_norm_basic_lstm_cell/{n}/gamma"] = torch.ones(H, dtype=dtype)
        p[f"{prefix}/layer_norm_basic_lstm_cell/{n}/beta"] = torch.zeros(H, dtype=dtype)
    # gen:88 / disc:90  dense(name="decoder")
    p[f"{prefix}/decoder/kernel"] = _glorot_uniform((H, out_units), gen, dtype)
    p[f"{prefix}/decoder/bias"] = torch.zeros(out_units, dtype=dtype)
    return p


def init_generator_params(vocab_size: int, seed: int = 0, R: int = 196, C: int = 512,
                          H: int = LSTM_UNITS, dtype=torch.float32) -> "OrderedDict[str, Tensor]":
    """Hot-path variables of `Generator/Generator/*` (train:85-88 scope + gen:15,79,88)."""
    g = torch.Generator().manual_seed(seed)
    return _network_params("Generator/Generator", C, vocab_size, R, C, H, g, dtype)  # noise has z_hat's shape [B,C] (gen:81)


def init_discriminator_params(vocab_size: int, seed: int = 1, R: int = 196, C: int = 512,
                              H: int = LSTM_UNITS, E: int = EMBED_DIM,
                              dtype=torch.float32) -> "OrderedDict[str, Tensor]":
    """`Discriminator/Discriminator/*` plus `Discriminator/W` (train:70, OOV init U(-.1,.1)
    as in dataset_creation/map_files_to_triples.py:24)."""
    g = torch.Generator().manual_seed(seed)
    p = _network_params("Discriminator/Discriminator", E, 1, R, C, H, g, dtype)
    p["Discriminator/W"] = ((torch.rand((vocab_size, E), generator=g, dtype=torch.float64) * 0.2) - 0.1).to(dtype)
    return p


# --------------------------------------------------------------------------------------
# TF op restatements
# --------------------------------------------------------------------------------------
def layer_norm(x: Tensor, gamma: Tensor, beta: Tensor) -> Tensor:
    """tf.contrib.layers.layer_norm on a 2-D input: biased moments over the last axis,
    y = (x - mean) * rsqrt(var + 1e-12) * gamma + beta."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + LN_EPS) * gamma + beta


def ln_lstm_cell(p: Dict[str, Tensor], prefix: str, inputs: Tensor,
                 state: Tuple[Tensor, Tensor]) -> Tuple[Tensor, Tuple[Tensor, Tensor]]:
    """tf.contrib.rnn.LayerNormBasicLSTMCell.call (gen:79,87 / disc:81,89).
    state = LSTMStateTuple(c, h); gate order i, j, f, o; no bias when layer_norm=True."""
    c, h = state
    cell = f"{prefix}/layer_norm_basic_lstm_cell"
    args = torch.cat([inputs, h], dim=1)
    concat = args @ p[f"{cell}/kernel"]
    i, j, f, o = torch.chunk(concat, 4, dim=1)
    i = layer_norm(i, p[f"{cell}/input/gamma"], p[f"{cell}/input/beta"])
    j = layer_norm(j, p[f"{cell}/transform/gamma"], p[f"{cell}/transform/beta"])
    f = layer_norm(f, p[f"{cell}/forget/gamma"], p[f"{cell}/forget/beta"])
    o = layer_norm(o, p[f"{cell}/output/gamma"], p[f"{cell}/output/beta"])
    g = torch.tanh(j)
    new_c = c * torch.sigmoid(f + FORGET_BIAS) + torch.sigmoid(i) * g
    new_c = layer_norm(new_c, p[f"{cell}/state/gamma"], p[f"{cell}/state/beta"])
    new_h = torch.tanh(new_c) * torch.sigmoid(o)
    return new_h, (new_c, new_h)


def attention_mechanism(p: Dict[str, Tensor], prefix: str, flattened_context: Tensor,
                        partially_flattened_context: Tensor,
                        cell_state: Tuple[Tensor, Tensor]) -> Tuple[Tensor, Tensor]:
    """gen:13-18 / disc:13-18 (byte-identical).  Uses cell_state[0] == c (LSTMStateTuple
    order is (c, h)).  One dense layer, no tanh, no mask."""
    context_and_state = torch.cat([flattened_context, cell_state[0]], dim=1)               # gen:14
    e = context_and_state @ p[f"{prefix}/attention_perceptron/kernel"] \
        + p[f"{prefix}/attention_perceptron/bias"]                                         # gen:15
    alpha = torch.softmax(e, dim=-1)                                                       # gen:16
    z_hat = (partially_flattened_context * alpha.unsqueeze(2)).sum(dim=1)                  # gen:17
    return z_hat, alpha


def generator_forward(p: Dict[str, Tensor], annotations: Tensor, noise: Tensor,
                      n_steps: int = 3, return_aux: bool = False):
    """gen:74-91 starting at self.downsampled.  annotations [B,14,14,512] (NHWC) or
    [B,R,C]; noise [B,512] shared by all timesteps (gen:81,86)."""
    prefix = "Generator/Generator"
    B = annotations.shape[0]
    Cc = annotations.shape[-1]
    flattened_context = annotations.reshape(B, -1)                                         # gen:74
    partially_flattened_context = annotations.reshape(B, -1, Cc)                           # gen:75
    state0 = partially_flattened_context.mean(dim=1)                                       # gen:76
    state = (state0, state0)                                                               # gen:77
    decoded, alphas = [], []
    for _ in range(n_steps):                                                               # gen:85
        z_hat, alpha = attention_mechanism(p, prefix, flattened_context,
                                           partially_flattened_context, state)
        next_input = torch.cat([z_hat, noise], dim=1)                                      # gen:86
        output, state = ln_lstm_cell(p, prefix, next_input, state)                         # gen:87
        decoded.append(output @ p[f"{prefix}/decoder/kernel"] + p[f"{prefix}/decoder/bias"])  # gen:88
        alphas.append(alpha)
    logits = torch.stack(decoded, dim=1)                                                   # gen:90
    if return_aux:
        return logits, {"alpha": torch.stack(alphas, 1), "c": state[0], "h": state[1]}
    return logits


def discriminator_forward(p: Dict[str, Tensor], input_triples: Tensor, annotations: Tensor,
                          n_steps: int = 3, return_aux: bool = False):
    """disc:73-93.  input_triples [B,T,V] float (one-hot reals or raw fake logits)."""
    prefix = "Discriminator/Discriminator"
    B = annotations.shape[0]
    Cc = annotations.shape[-1]
    flattened_context = annotations.reshape(B, -1)                                         # disc:73
    partially_flattened_context = annotations.reshape(B, -1, Cc)                           # disc:74
    state0 = partially_flattened_context.mean(dim=1)                                       # disc:76
    state = (state0, state0)                                                               # disc:77
    outs, alphas = [], []
    for i in range(n_steps):                                                               # disc:85
        indices = input_triples[:, i, :]                                                   # disc:86
        embedding = indices @ p["Discriminator/W"]                                         # disc:87
        z_hat, alpha = attention_mechanism(p, prefix, flattened_context,
                                           partially_flattened_context, state)
        next_input = torch.cat([z_hat, embedding], dim=1)                                  # disc:88
        output, state = ln_lstm_cell(p, prefix, next_input, state)                         # disc:89
        outs.append(output @ p[f"{prefix}/decoder/kernel"] + p[f"{prefix}/decoder/bias"])  # disc:90
        alphas.append(alpha)
    scores = torch.stack(outs, dim=1)                                                      # disc:92
    if return_aux:
        return scores, {"alpha": torch.stack(alphas, 1), "c": state[0], "h": state[1]}
    return scores


# --------------------------------------------------------------------------------------
# tfgan losses (train:239-253)
# --------------------------------------------------------------------------------------
def wgan_gp_losses(gp: Dict[str, Tensor], dp: Dict[str, Tensor], ann_g: Tensor, ann_d: Tensor,
                   real: Tensor, noise: Tensor, gp_alpha: Tensor, lam: float, n_steps: int = 3,
                   create_graph: bool = True) -> Dict[str, Tensor]:
    """tfgan.gan_model + gan_loss(wasserstein_*, gradient_penalty_weight=lam,
    gradient_penalty_one_sided=True)  (train:239-253).

    gp_alpha is tfgan's `alpha = random_uniform([B,1,1])`, injected for determinism.
    Returns gen_cost, disc_cost and the pieces (w_disc, gp, slopes, fake).
    """
    fake = generator_forward(gp, ann_g, noise, n_steps)                       # G once
    d_fake = discriminator_forward(dp, fake, ann_d, n_steps)                  # D(fake)
    d_real = discriminator_forward(dp, real, ann_d, n_steps)                  # D(real), shared vars
    gen_cost = -d_fake.mean()                                                 # wasserstein_generator_loss
    w_disc = d_fake.mean() - d_real.mean()                                    # wasserstein_discriminator_loss
    # wasserstein_gradient_penalty: differences = generated - real; interpolates = real + alpha*diff
    differences = fake - real
    interpolates = real + gp_alpha.reshape(-1, 1, 1) * differences
    if not interpolates.requires_grad:
        interpolates = interpolates.clone().requires_grad_(True)
    d_interp = discriminator_forward(dp, interpolates, ann_d, n_steps)
    gradients = torch.autograd.grad(d_interp.sum(), interpolates, create_graph=create_graph)[0]
    gradient_squares = (gradients ** 2).sum(dim=(1, 2))
    slopes = torch.sqrt(gradient_squares + GP_EPS)
    penalties = torch.clamp(slopes / GP_TARGET - 1.0, min=0.0)                # one_sided=True
    penalty = (penalties ** 2).mean()                                         # mean over batch
    disc_cost = w_disc + lam * penalty
    return {"gen_cost": gen_cost, "disc_cost": disc_cost, "w_disc": w_disc, "gp": penalty,
            "slopes": slopes, "fake": fake, "d_fake": d_fake, "d_real": d_real,
            "gp_gradients": gradients}


def disc_step_grads(gp, dp, ann_g, ann_d, real, noise, gp_alpha, lam, n_steps=3, ann_grad=False):
    """Gradients of disc_cost w.r.t. every `Discriminator*` variable (train:263,266).
    G's forward is a constant here (var_list = disc_params).
    ann_grad: also return d disc_cost / d self.downsampled of the discriminator (disc:68) -- what TF
    back-propagates into the discriminator's conv variables, which are in disc_params (train:263)."""
    dp_req = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in dp.items())
    if ann_grad:
        ann_d = ann_d.detach().clone().requires_grad_(True)
    with torch.no_grad():
        fake = generator_forward(gp, ann_g, noise, n_steps)
    d_fake = discriminator_forward(dp_req, fake, ann_d, n_steps)
    d_real = discriminator_forward(dp_req, real, ann_d, n_steps)
    w_disc = d_fake.mean() - d_real.mean()
    interpolates = (real + gp_alpha.reshape(-1, 1, 1) * (fake - real)).detach().requires_grad_(True)
    d_interp = discriminator_forward(dp_req, interpolates, ann_d, n_steps)
    gradients = torch.autograd.grad(d_interp.sum(), interpolates, create_graph=True)[0]
    slopes = torch.sqrt((gradients ** 2).sum(dim=(1, 2)) + GP_EPS)
    penalty = (torch.clamp(slopes / GP_TARGET - 1.0, min=0.0) ** 2).mean()
    disc_cost = w_disc + lam * penalty
    grads = torch.autograd.grad(disc_cost, list(dp_req.values()) + ([ann_d] if ann_grad else []), allow_unused=True)
    gd = OrderedDict((k, (g if g is not None else torch.zeros_like(v)))
                     for (k, v), g in zip(dp_req.items(), grads))
    extra = {"ann_grad": grads[-1].detach()} if ann_grad else {}
    return {**extra, "disc_cost": disc_cost.detach(), "w_disc": w_disc.detach(), "gp": penalty.detach(),
            "slopes": slopes.detach(), "gp_gradients": gradients.detach(), "fake": fake,
            "d_fake": d_fake.detach(), "d_real": d_real.detach(), "d_interp": d_interp.detach(),
            "grads": gd}


def gen_step_grads(gp, dp, ann_g, ann_d, noise, n_steps=3, ann_grad=False):
    """Gradients of gen_cost = -mean(D(G(z))) w.r.t. every `Generator*` variable
    (train:262,265).  D's variables are constants here.
    ann_grad: also return d gen_cost / d self.downsampled of the generator (gen:68)."""
    gp_req = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in gp.items())
    if ann_grad:
        ann_g = ann_g.detach().clone().requires_grad_(True)
    fake = generator_forward(gp_req, ann_g, noise, n_steps)
    d_fake = discriminator_forward(dp, fake, ann_d, n_steps)
    gen_cost = -d_fake.mean()
    grads = torch.autograd.grad(gen_cost, list(gp_req.values()) + ([ann_g] if ann_grad else []), allow_unused=True)
    gd = OrderedDict((k, (g if g is not None else torch.zeros_like(v)))
                     for (k, v), g in zip(gp_req.items(), grads))
    extra = {"ann_grad": grads[-1].detach()} if ann_grad else {}
    return {**extra, "gen_cost": gen_cost.detach(), "fake": fake.detach(), "d_fake": d_fake.detach(), "grads": gd}


# --------------------------------------------------------------------------------------
# tf.train.AdamOptimizer (train:258-259)
# --------------------------------------------------------------------------------------
class TFAdam:
    """tf.train.AdamOptimizer(lr, beta1, beta2) update rule (epsilon OUTSIDE the bias
    correction: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps))."""

    def __init__(self, params: Dict[str, Tensor], lr=1e-4, beta1=0.5, beta2=0.9, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in params.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in params.items())

    @torch.no_grad()
    def step(self, params: Dict[str, Tensor], grads: Dict[str, Tensor]) -> None:
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k, th in params.items():
            g = grads[k]
            self.m[k].mul_(self.b1).add_(g, alpha=1.0 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
            th.sub_(lr_t * self.m[k] / (self.v[k].sqrt() + self.eps))


def train_iteration(gp, dp, adam_g: TFAdam, adam_d: TFAdam, ann_g, ann_d, real,
                    noises, gp_alphas, lam: float, critic_iters: int, n_steps: int = 3):
    """One pass of the reference loop body train:362-368: CRITIC_ITERS D steps then one G
    step, ALL ON THE SAME data batch (train:185-187), fresh noise / alpha per sess.run.
    `noises` has critic_iters+1 entries, `gp_alphas` has critic_iters entries."""
    log = {"disc_cost": [], "gp": []}
    for i in range(critic_iters):                                             # train:364-365
        r = disc_step_grads(gp, dp, ann_g, ann_d, real, noises[i], gp_alphas[i], lam, n_steps)
        adam_d.step(dp, r["grads"])
        log["disc_cost"].append(float(r["disc_cost"]))
        log["gp"].append(float(r["gp"]))
    r = gen_step_grads(gp, dp, ann_g, ann_d, noises[critic_iters], n_steps)   # train:368
    adam_g.step(gp, r["grads"])
    log["gen_cost"] = float(r["gen_cost"])
    return log


# --------------------------------------------------------------------------------------
# Synthetic inputs
# --------------------------------------------------------------------------------------
def greedy_triples(p: Dict[str, Tensor], annotations: Tensor, noise: Tensor, n_steps: int = 3) -> Tensor:
    """train:269-270: ``fake_triples = tf.argmax(self._Generator(images), axis=-1)`` -> [B, n_steps] token ids
    (tf.argmax returns the lowest index among equal maxima, as torch.argmax does)."""
    return generator_forward(p, annotations, noise, n_steps).argmax(dim=-1)


def recall(fake, real, N: float) -> float:
    """train:294-295 ``_recall``: |set(fake triples) & set(real triples)| / N."""
    return float(len(set(map(tuple, fake)).intersection(set(map(tuple, real))))) / N


def recall_at_k(fake: Tensor, scores: Tensor, real: Tensor, k: int) -> float:
    """train:320-327 with the ranking the comments there describe ("Sort by discriminator score"): the k fakes with
    the highest mean critic score against the real triples.  (The reference's ``score_accumulator.argsort()`` acts on
    an [N,1] array, i.e. sorts each 1-element row and yields zeros: a bug that is documented, not reproduced.)"""
    order = torch.argsort(-scores.reshape(-1), stable=True)[:k]
    return recall(fake[order].tolist(), real.tolist(), float(k))


def synthetic_batch(B: int, V: int, T: int = 3, R: int = 196, C: int = 512, seed: int = 1234,
                    dtype=torch.float32, bf16_exact: bool = False):
    """ann_g, ann_d ~ N(0,1) [B,R,C]; labels uniform ints -> one-hot [B,T,V] (train:173)."""
    g = torch.Generator().manual_seed(seed)
    ann_g = torch.randn((B, R, C), generator=g, dtype=torch.float32)
    ann_d = torch.randn((B, R, C), generator=g, dtype=torch.float32)
    if bf16_exact:
        ann_g = ann_g.bfloat16().float()
        ann_d = ann_d.bfloat16().float()
    labels = torch.randint(0, V, (B, T), generator=g)
    real = torch.nn.functional.one_hot(labels, V).to(dtype)
    return ann_g.to(dtype), ann_d.to(dtype), labels, real
