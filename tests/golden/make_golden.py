"""Generates the golden fixtures of the parity tests from the CPU oracle (fp64).

The reference itself (Python 2 + TensorFlow 1.x) cannot be imported in this image, so these vectors
pin the ORACLE, not TensorFlow: `python -m tests.golden.make_golden` rewrites tests/golden/*.json;
tests/test_oracle.py re-derives them on every CPU run and tests/test_steps_gpu.py compares the CUDA
path with them on the B200.  Inputs are not stored: they are regenerated from the seeds in `config`
by tests.util.make_problem (torch CPU generator).  Large gradient tensors are stored as their L2 norm
plus 48 entries at fixed pseudo-random positions.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
N_SAMPLES = 48


def sample_index(name: str, numel: int) -> np.ndarray:
    seed = sum(ord(c) * (i + 1) for i, c in enumerate(name)) % (2 ** 31)
    return np.random.RandomState(seed).randint(0, numel, size=min(N_SAMPLES, numel))


def _rec(name, t):
    flat = t.detach().double().reshape(-1)
    idx = sample_index(name, flat.numel())
    return {"norm": float(flat.norm()), "samples": [float(x) for x in flat[torch.from_numpy(idx)]]}


def compute(B, T, V, R, seed, lam):
    from oracle import sgg_oracle as O
    from tests.util import make_problem
    p = make_problem(B, T, V, R=R, seed=seed, dtype=torch.float64)
    d = O.disc_step_grads(p["gp"], p["dp"], p["ann_g"], p["ann_d"], p["real"], p["noise"], p["alpha"], lam, T)
    g = O.gen_step_grads(p["gp"], p["dp"], p["ann_g"], p["ann_d"], p["noise"], T)
    return {
        "config": {"B": B, "T": T, "V": V, "R": R, "seed": seed, "lam": lam},
        "w_disc": float(d["w_disc"]), "gp": float(d["gp"]), "gen_cost": float(g["gen_cost"]),
        "slopes": [float(x) for x in d["slopes"]],
        "logits": [float(x) for x in d["fake"].reshape(-1)],
        "d_fake": [float(x) for x in d["d_fake"].reshape(-1)],
        "d_real": [float(x) for x in d["d_real"].reshape(-1)],
        "d_grads": {k: _rec(k, v) for k, v in d["grads"].items()},
        "g_grads": {k: _rec(k, v) for k, v in g["grads"].items()},
    }


CONFIGS = {
    "step_B4_T3_V64_R12.json": dict(B=4, T=3, V=64, R=12, seed=11, lam=10.0),
    "step_B3_T5_V40_R30.json": dict(B=3, T=5, V=40, R=30, seed=12, lam=10.0),
}

if __name__ == "__main__":
    for fn, cfg in CONFIGS.items():
        with open(os.path.join(HERE, fn), "w") as f:
            json.dump(compute(**cfg), f, indent=0)
        print("wrote", fn)
