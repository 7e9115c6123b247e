"""tests/plan_mirror.py (the hoisted / pass-batched / reverse-over-tangent plan the CUDA kernels
implement) against the literal oracle (concat-form attention, autograd double backward), fp64, CPU.
Equality here proves the plan is an exact algebraic rewrite of the reference arithmetic; the GPU
tests then check every kernel against the oracle through the C ABI."""
import pytest
import torch

from oracle import sgg_oracle as O
from tests import plan_mirror as M
from tests.test_oracle import _small_problem


def _close(a, b, tol=1e-9):
    # absolute floor: d disc_cost / d decoder.bias is analytically 0 (mean fake - mean real, GP is bias-free)
    return (a - b).norm().item() <= tol * b.norm().item() + 1e-13


@pytest.mark.parametrize("B,T,V", [(3, 3, 11), (2, 5, 7), (1, 1, 4)])
def test_disc_step_plan_equals_oracle(B, T, V):
    gp, dp, ann_g, ann_d, labels, real, noise, alpha = _small_problem(B, T, V)
    ref = O.disc_step_grads(gp, dp, ann_g, ann_d, real, noise, alpha, 10.0, T, ann_grad=True)
    got = M.disc_step(gp, dp, ann_g, ann_d, labels, noise, alpha, 10.0, T)
    assert _close(got["fake"], ref["fake"])
    assert _close(got["ann_grad"], ref["ann_grad"], 1e-8)     # d disc_cost / d self.downsampled (disc:68)
    assert abs(float(got["w_disc"] - ref["w_disc"])) < 1e-12
    assert abs(float(got["gp"] - ref["gp"])) < 1e-11
    assert _close(got["slopes"], ref["slopes"])
    assert _close(got["gp_gradients"], ref["gp_gradients"])
    assert set(got["grads"]) == set(ref["grads"])
    for k, v in ref["grads"].items():
        assert _close(got["grads"][k], v, 1e-8), k


@pytest.mark.parametrize("B,T,V", [(3, 3, 11), (2, 4, 9)])
def test_gen_step_plan_equals_oracle(B, T, V):
    gp, dp, ann_g, ann_d, labels, real, noise, alpha = _small_problem(B, T, V)
    ref = O.gen_step_grads(gp, dp, ann_g, ann_d, noise, T, ann_grad=True)
    got = M.gen_step(gp, dp, ann_g, ann_d, noise, T)
    assert abs(float(got["gen_cost"] - ref["gen_cost"])) < 1e-12
    assert _close(got["ann_grad"], ref["ann_grad"], 1e-8)     # d gen_cost / d self.downsampled (gen:68)
    for k, v in ref["grads"].items():
        assert _close(got["grads"][k], v, 1e-8), k


def test_gp_inactive_side_gives_zero_penalty_gradient():
    """one_sided=True: slopes below the target contribute nothing (train:250)."""
    gp, dp, ann_g, ann_d, labels, real, noise, alpha = _small_problem(3, 3, 11)
    dp = {k: (v * 1e-3 if k == "Discriminator/W" else v) for k, v in dp.items()}   # tiny input gradients
    ref = O.disc_step_grads(gp, dp, ann_g, ann_d, real, noise, alpha, 10.0, 3)
    assert float(ref["gp"]) == 0.0 and (ref["slopes"] < 1).all()
    got = M.disc_step(gp, dp, ann_g, ann_d, labels, noise, alpha, 10.0, 3)
    assert float(got["gp"]) == 0.0
    for k, v in ref["grads"].items():
        assert _close(got["grads"][k], v, 1e-8), k
