"""Oracle self-consistency (CPU).  The reference ships no tests or golden vectors, so
the oracle is pinned by: (1) an independent numpy restatement of the forward pass written from the
reference's call sites, (2) fp64 finite differences of its first- and second-order gradients,
(3) structural invariants of the reference arithmetic, (4) the committed golden fixtures."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import sgg_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SMALL = dict(R=6, C=8, H=8, E=5)


def _small_problem(B=3, T=3, V=11, seed=0, dtype=torch.float64):
    gp = O.init_generator_params(V, seed=seed, R=SMALL["R"], C=SMALL["C"], H=SMALL["H"], dtype=dtype)
    dp = O.init_discriminator_params(V, seed=seed + 1, R=SMALL["R"], C=SMALL["C"], H=SMALL["H"], E=SMALL["E"], dtype=dtype)
    g = torch.Generator().manual_seed(seed + 2)
    for p in (gp, dp):
        for k in p:  # non-trivial LN / bias values so every term of the gradients is exercised
            if "gamma" in k:
                p[k] = 1 + 0.3 * torch.randn(p[k].shape, generator=g, dtype=dtype)
            elif "beta" in k or "bias" in k:
                p[k] = 0.3 * torch.randn(p[k].shape, generator=g, dtype=dtype)
            elif "kernel" in k:
                p[k] = p[k] * 3.0
        if "Discriminator/W" in p:
            p["Discriminator/W"] = p["Discriminator/W"] * 20.0
    ann_g, ann_d, labels, real = O.synthetic_batch(B, V, T, SMALL["R"], SMALL["C"], seed=seed + 3, dtype=dtype)
    noise = torch.randn(B, SMALL["C"], generator=g, dtype=dtype)
    alpha = torch.rand(B, generator=g, dtype=dtype)
    return gp, dp, ann_g, ann_d, labels, real, noise, alpha


# ------------------------------------------------------------------ (1) independent numpy restatement
def _np_ln(x, g, b):
    m = x.mean(-1, keepdims=True)
    v = ((x - m) ** 2).mean(-1, keepdims=True)
    return (x - m) / np.sqrt(v + 1e-12) * g + b


def _np_sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def _np_forward(p, prefix, ann, u_of_t, T):
    """gen:74-91 / disc:73-93 in numpy, written independently of oracle/sgg_oracle.py."""
    P = {k: v.numpy() for k, v in p.items()}
    B, R, Cc = ann.shape
    flat = ann.reshape(B, R * Cc)
    c = h = ann.mean(1)
    cell = prefix + "/layer_norm_basic_lstm_cell"
    outs = []
    for t in range(T):
        e = np.concatenate([flat, c], 1) @ P[prefix + "/attention_perceptron/kernel"] + P[prefix + "/attention_perceptron/bias"]
        e = e - e.max(-1, keepdims=True)
        al = np.exp(e) / np.exp(e).sum(-1, keepdims=True)
        z = (ann * al[:, :, None]).sum(1)
        q = np.concatenate([z, u_of_t(t), h], 1) @ P[cell + "/kernel"]
        H = q.shape[1] // 4
        i, j, f, o = (q[:, k * H:(k + 1) * H] for k in range(4))
        i = _np_ln(i, P[cell + "/input/gamma"], P[cell + "/input/beta"])
        j = _np_ln(j, P[cell + "/transform/gamma"], P[cell + "/transform/beta"])
        f = _np_ln(f, P[cell + "/forget/gamma"], P[cell + "/forget/beta"])
        o = _np_ln(o, P[cell + "/output/gamma"], P[cell + "/output/beta"])
        c = c * _np_sig(f + 1.0) + _np_sig(i) * np.tanh(j)
        c = _np_ln(c, P[cell + "/state/gamma"], P[cell + "/state/beta"])
        h = np.tanh(c) * _np_sig(o)
        outs.append(h @ P[prefix + "/decoder/kernel"] + P[prefix + "/decoder/bias"])
    return np.stack(outs, 1)


def test_forward_matches_independent_numpy_restatement():
    gp, dp, ann_g, ann_d, labels, real, noise, alpha = _small_problem()
    fake = O.generator_forward(gp, ann_g, noise, 3)
    ref = _np_forward(gp, "Generator/Generator", ann_g.numpy(), lambda t: noise.numpy(), 3)
    np.testing.assert_allclose(fake.numpy(), ref, rtol=1e-10, atol=1e-12)
    W = dp["Discriminator/W"].numpy()
    for x in (fake, real):
        got = O.discriminator_forward(dp, x, ann_d, 3)
        ref = _np_forward(dp, "Discriminator/Discriminator", ann_d.numpy(), lambda t: x.numpy()[:, t] @ W, 3)
        np.testing.assert_allclose(got.numpy(), ref, rtol=1e-10, atol=1e-12)
        assert got.shape == (3, 3, 1)                                         # disc:92-93 [B,3,1]
    assert fake.shape == (3, 3, 11)                                           # gen:90-91 [B,3,V]


def test_nhwc_annotations_flatten_like_the_reference():
    """gen:74-75: [B,14,14,512] reshaped to [B,100352] and [B,196,512]; region r = h*14+w."""
    gp = O.init_generator_params(7, R=4, C=8, H=8, dtype=torch.float64)
    a = torch.randn(2, 2, 2, 8, dtype=torch.float64)
    n = torch.randn(2, 8, dtype=torch.float64)
    assert torch.equal(O.generator_forward(gp, a, n), O.generator_forward(gp, a.reshape(2, 4, 8), n))


# ------------------------------------------------------------------ (3) invariants
def test_attention_invariants_and_split_form():
    gp, dp, ann_g, *_ = _small_problem()
    B, R, Cc = ann_g.shape
    c = torch.randn(B, SMALL["H"], dtype=torch.float64)
    z, alpha = O.attention_mechanism(gp, "Generator/Generator", ann_g.reshape(B, -1), ann_g, (c, None))
    assert torch.allclose(alpha.sum(-1), torch.ones(B, dtype=torch.float64), atol=1e-14)
    assert (alpha > 0).all()
    lo, hi = ann_g.min(1).values, ann_g.max(1).values                          # z_hat inside the convex hull
    assert ((z >= lo - 1e-12) & (z <= hi + 1e-12)).all()
    # row-split of the single dense kernel: e = flat(a) W_a + c W_h + b, uses c (state[0]) not h
    W = gp["Generator/Generator/attention_perceptron/kernel"]
    e = ann_g.reshape(B, -1) @ W[:R * Cc] + c @ W[R * Cc:] + gp["Generator/Generator/attention_perceptron/bias"]
    assert torch.allclose(torch.softmax(e, -1), alpha, atol=1e-14)


def test_layer_norm_is_tf_contrib_layer_norm():
    x = torch.randn(5, 16, dtype=torch.float64) * 3 + 1
    g, b = torch.rand(16, dtype=torch.float64) + 0.5, torch.randn(16, dtype=torch.float64)
    y = O.layer_norm(x, g, b)
    n = (y - b) / g
    assert torch.allclose(n.mean(-1), torch.zeros(5, dtype=torch.float64), atol=1e-12)
    assert torch.allclose((n * n).mean(-1), torch.ones(5, dtype=torch.float64), atol=1e-9)   # biased variance
    ref = torch.nn.functional.layer_norm(x, (16,), g, b, eps=1e-12)
    assert torch.allclose(y, ref, atol=1e-12)


def test_lstm_cell_gate_order_and_forget_bias():
    """LayerNormBasicLSTMCell: columns [i|j|f|o], forget_bias 1.0 added after LN, no kernel bias."""
    H, Cin = 4, 3
    p = O._network_params("N", Cin - H if Cin > H else 0, 2, 1, H, H, torch.Generator().manual_seed(0), torch.float64)
    cell = "N/layer_norm_basic_lstm_cell"
    K = p[cell + "/kernel"]
    x = torch.randn(2, K.shape[0] - H, dtype=torch.float64)
    c, h = torch.randn(2, H, dtype=torch.float64), torch.randn(2, H, dtype=torch.float64)
    new_h, (new_c, new_h2) = O.ln_lstm_cell(p, "N", x, (c, h))
    assert new_h is new_h2
    q = torch.cat([x, h], 1) @ K
    ln = lambda v: torch.nn.functional.layer_norm(v, (H,), eps=1e-12)
    i, j, f, o = (ln(q[:, k * H:(k + 1) * H]) for k in range(4))
    c2 = ln(c * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j))
    assert torch.allclose(new_c, c2, atol=1e-12)
    assert torch.allclose(new_h, torch.tanh(c2) * torch.sigmoid(o), atol=1e-12)


def test_gradient_penalty_formula():
    """tfgan wasserstein_gradient_penalty(one_sided=True, target=1, eps=1e-10), mean over batch (train:245-250)."""
    gp, dp, ann_g, ann_d, labels, real, noise, alpha = _small_problem()
    r = O.wgan_gp_losses(gp, dp, ann_g, ann_d, real, noise, alpha, lam=10.0)
    g = r["gp_gradients"]
    s = torch.sqrt((g * g).sum((1, 2)) + 1e-10)
    assert torch.allclose(r["slopes"], s)
    assert torch.allclose(r["gp"], (torch.clamp(s - 1, min=0) ** 2).mean())
    assert torch.allclose(r["disc_cost"], r["d_fake"].mean() - r["d_real"].mean() + 10.0 * r["gp"])
    assert torch.allclose(r["gen_cost"], -r["d_fake"].mean())
    assert (s > 1).any(), "test problem must exercise the active side of the one-sided penalty"


# ------------------------------------------------------------------ (2) finite differences
def _fd(f, x, idxs, eps=1e-6):
    out = []
    flat = x.view(-1)
    for i in idxs:
        old = flat[i].item()
        flat[i] = old + eps
        fp = f()
        flat[i] = old - eps
        fm = f()
        flat[i] = old
        out.append((fp - fm) / (2 * eps))
    return torch.tensor(out, dtype=torch.float64)


def test_disc_step_gradients_by_finite_differences():
    """Second-order path: d(disc_cost)/d(theta_D) includes the gradient of the gradient penalty."""
    gp, dp, ann_g, ann_d, labels, real, noise, alpha = _small_problem()
    r = O.disc_step_grads(gp, dp, ann_g, ann_d, real, noise, alpha, 10.0, 3)
    assert float(r["gp"]) > 1e-4

    def cost():
        with torch.enable_grad():
            return float(O.wgan_gp_losses(gp, dp, ann_g, ann_d, real, noise, alpha, 10.0, 3, create_graph=False)["disc_cost"])

    rng = np.random.RandomState(0)
    for k, v in dp.items():
        idxs = rng.choice(v.numel(), size=min(4, v.numel()), replace=False)
        fd = _fd(cost, v, idxs)
        an = r["grads"][k].reshape(-1)[idxs]
        assert torch.allclose(fd, an, rtol=2e-5, atol=1e-8), (k, fd, an)


def test_gen_step_gradients_by_finite_differences():
    gp, dp, ann_g, ann_d, labels, real, noise, alpha = _small_problem()
    r = O.gen_step_grads(gp, dp, ann_g, ann_d, noise, 3)

    def cost():
        with torch.no_grad():
            return float(-O.discriminator_forward(dp, O.generator_forward(gp, ann_g, noise, 3), ann_d, 3).mean())

    rng = np.random.RandomState(1)
    for k, v in gp.items():
        idxs = rng.choice(v.numel(), size=min(4, v.numel()), replace=False)
        fd = _fd(cost, v, idxs)
        an = r["grads"][k].reshape(-1)[idxs]
        assert torch.allclose(fd, an, rtol=2e-5, atol=1e-9), (k, fd, an)
    # G's step leaves D untouched and vice versa (train:262-266 var_list split)
    assert set(r["grads"]) == set(gp)


def test_tf_adam_update_rule():
    """tf.train.AdamOptimizer(1e-4, 0.5, 0.9): epsilon outside the bias correction (train:258-259)."""
    p = {"w": torch.tensor([1.0, -2.0, 3.0], dtype=torch.float64)}
    opt = O.TFAdam(p, lr=1e-4, beta1=0.5, beta2=0.9, eps=1e-8)
    m = v = np.zeros(3)
    th = p["w"].numpy().copy()
    for t in range(1, 4):
        g = np.array([0.1 * t, -0.2, 1e-9])
        opt.step(p, {"w": torch.tensor(g)})
        m = 0.5 * m + 0.5 * g
        v = 0.9 * v + 0.1 * g * g
        lr_t = 1e-4 * math.sqrt(1 - 0.9 ** t) / (1 - 0.5 ** t)
        th = th - lr_t * m / (np.sqrt(v) + 1e-8)
        np.testing.assert_allclose(p["w"].numpy(), th, rtol=1e-14)


def test_train_iteration_schedule():
    """train:362-368 + train:185-187: critic_iters D steps then one G step on the same batch."""
    gp, dp, ann_g, ann_d, labels, real, noise, alpha = _small_problem()
    gp0 = {k: v.clone() for k, v in gp.items()}
    dp0 = {k: v.clone() for k, v in dp.items()}
    ag, ad = O.TFAdam(gp), O.TFAdam(dp)
    g = torch.Generator().manual_seed(5)
    noises = [torch.randn(noise.shape, generator=g, dtype=torch.float64) for _ in range(3)]
    alphas = [torch.rand(alpha.shape, generator=g, dtype=torch.float64) for _ in range(2)]
    log = O.train_iteration(gp, dp, ag, ad, ann_g, ann_d, real, noises, alphas, 10.0, 2, 3)
    assert len(log["disc_cost"]) == 2 and ad.t == 2 and ag.t == 1
    assert any(not torch.equal(gp[k], gp0[k]) for k in gp)
    assert any(not torch.equal(dp[k], dp0[k]) for k in dp)


# ------------------------------------------------------------------ (4) golden fixtures
def test_oracle_reproduces_committed_golden_vectors():
    """tests/golden/*.json were produced by tests/golden/make_golden.py from this oracle in fp64;
    this guards the oracle (and the fixtures the GPU parity tests read) against silent drift."""
    from tests.golden import make_golden as MG
    path = os.path.join(HERE, "golden", "step_B4_T3_V64_R12.json")
    with open(path) as f:
        gold = json.load(f)
    fresh = MG.compute(**gold["config"])
    for k in ("w_disc", "gp", "gen_cost"):
        assert abs(fresh[k] - gold[k]) <= 1e-12 * max(1.0, abs(gold[k])), k
    np.testing.assert_allclose(np.array(fresh["slopes"]), np.array(gold["slopes"]), rtol=1e-11)
    np.testing.assert_allclose(np.array(fresh["logits"]), np.array(gold["logits"]), rtol=1e-10, atol=1e-12)
    for net in ("d_grads", "g_grads"):
        for k, rec in gold[net].items():
            assert abs(fresh[net][k]["norm"] - rec["norm"]) <= 1e-10 * max(rec["norm"], 1e-30), k
            np.testing.assert_allclose(np.array(fresh[net][k]["samples"]), np.array(rec["samples"]), rtol=1e-8, atol=1e-14)


def test_greedy_triples_and_recall_known_answers():
    """train:269-270 argmax decoding (lowest index wins ties) and train:294-295 / 320-327 recall."""
    from oracle import sgg_oracle as O
    gp = O.init_generator_params(11, seed=5, R=6, dtype=torch.float64)
    ann, _, _, _ = O.synthetic_batch(3, 11, 2, 6, 512, seed=9, dtype=torch.float64)
    noise = torch.randn(3, 512, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    tok = O.greedy_triples(gp, ann, noise, 2)
    assert tok.shape == (3, 2) and tok.dtype == torch.int64
    assert torch.equal(tok, O.generator_forward(gp, ann, noise, 2).argmax(-1))
    assert torch.argmax(torch.tensor([1.0, 3.0, 3.0])).item() == 1            # tie -> lowest index, like tf.argmax
    assert O.recall([[1, 2, 3], [1, 2, 3], [4, 5, 0]], [[1, 2, 3], [9, 9, 9]], 50.0) == 1 / 50.0
    fake = torch.tensor([[1, 2, 3], [4, 5, 6], [7, 8, 9], [1, 1, 1]])
    scores = torch.tensor([0.1, 0.9, 0.5, 0.7])
    real = torch.tensor([[4, 5, 6], [1, 2, 3]])
    assert O.recall_at_k(fake, scores, real, 2) == 1 / 2                        # top-2 by score: rows 1, 3 -> one hit
    assert O.recall_at_k(fake, scores, real, 4) == 2 / 4


def test_u01_restatement_stays_inside_the_open_interval():
    """csrc/common.cuh u01(): ((x >> 9) + 0.5f) * 2^-23 in fp32.  Known answers for the extreme counter values: the
    largest draw must stay below 1 (the previous 24-bit form rounded (2^24 - 1) + 0.5 up to 2^24 and returned exactly
    1.0, which the Gumbel epilogue's -log(-log u) turned into +inf)."""
    import numpy as np

    def u01(x):
        return np.float32(np.float32(np.uint32(x) >> np.uint32(9)) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)

    def u01_old(x):
        return np.float32(np.float32(np.uint32(x) >> np.uint32(8)) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)

    assert u01(0xFFFFFFFF) == np.float32(1.0) - np.float32(2.0 ** -24) and u01(0xFFFFFFFF) < 1.0
    assert u01(0) == np.float32(2.0 ** -24) and u01(0) > 0.0
    assert u01_old(0xFFFFFFFF) == 1.0                        # the defect this form replaces
    g = -np.log(-np.log(np.float64(u01(0xFFFFFFFF))))
    assert np.isfinite(g)
