#!/usr/bin/env python
"""Where do two runs of the same D step part ways?  Repeats the long-sequence parity case (B 3, T 30, V 70, R 24), classifies
every repeat by its attention-kernel gradient error against the fp64 oracle (the usual ~1.3e-4 or the rarer ~6e-4 outcome,
see step_spread.py) and prints, buffer by buffer in the order the step computes them, the relative distance between one
run of each class next to the distance between two runs of the same class."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    import __graft_entry__ as ge
    ge.build()
    from oracle import sgg_oracle as O
    from tests.util import make_engine, make_problem, rel
    B, T, V, R = 3, 30, 70, 24
    prob = make_problem(B, T, V, R=R, seed=0, dtype=torch.float64)
    eng = make_engine(prob, B, T, V, R=R, lam=10.0)
    ref = O.disc_step_grads(prob["gp"], prob["dp"], prob["ann_g"], prob["ann_d"], prob["real"], prob["noise"], prob["alpha"], 10.0, T)
    key = "Discriminator/Discriminator/attention_perceptron/kernel"
    NR, RP, KXP, H, VP = 4 * B, 64, 1344, 512, 128
    bufs = [("FAKE", (T * B, 2 * VP), torch.bfloat16), ("d.Y", (NR * T,), torch.float32), ("d.EA", (T, NR, RP), torch.float32),
            ("d.Q", (T, NR, 4 * H), torch.float32), ("DFAKE", (T * B, VP), torch.float32), ("slopes", (B,), torch.float32),
            ("coef", (B,), torch.float32), ("VHL", (T * B, 2 * VP), torch.bfloat16), ("d.ED", (T, B, RP), torch.float32),
            ("d.QB", (T, NR, 8 * H), torch.bfloat16), ("d.XB", (T, NR, KXP), torch.float32), ("d.EB", (T, NR, 2 * RP), torch.bfloat16),
            ("d.CB", (T, NR, H), torch.float32), ("d.PB", (B, RP), torch.float32)]
    snaps = {"A": [], "B": []}
    for i in range(60):
        eng.disc_step(); torch.cuda.synchronize()
        gv = eng.d.grad_views()
        e = rel(gv[key], ref["grads"][key])
        cls = "B" if e > 3.5e-4 else "A"
        if len(snaps[cls]) < 2:
            s = {n: eng.ws_view(n, shp, dt).float().clone() for n, shp, dt in bufs}
            for n in ("FAKE", "VHL"):            # hi/lo pairs: compare the values they carry
                s[n] = s[n][:, :VP] + s[n][:, VP:]
            x = eng.ws_view("d.X", (T + 1, NR, 2 * KXP), torch.bfloat16).float()
            x = x[:, :, :KXP] + x[:, :, KXP:]
            s["X.z"], s["X.u"], s["X.h"] = x[:T, :, :512].clone(), x[:T, :, 512:812].clone(), x[1:, :, 812:1324].clone()
            s["g.X.h"] = (lambda y: (y[1:, :, :1536] + y[1:, :, 1536:])[:, :, 1024:].clone())(
                eng.ws_view("g.X", (T + 1, B, 2 * 1536), torch.bfloat16).float())
            s["grad"] = eng.d.grad.clone(); s["err"] = e
            snaps[cls].append(s)
        if len(snaps["A"]) == 2 and len(snaps["B"]) >= 1:
            break
    out = {"errors": {c: [s["err"] for s in v] for c, v in snaps.items()}}
    if snaps["B"]:
        a0, a1, b0 = snaps["A"][0], snaps["A"][1], snaps["B"][0]
        d = lambda x, y: float((x - y).norm() / (y.norm() + 1e-30))
        out["buffers"] = {n: {"A_vs_A": d(a1[n], a0[n]), "B_vs_A": d(b0[n], a0[n])}
                          for n in [b[0] for b in bufs] + ["grad", "X.z", "X.u", "X.h", "g.X.h"]}
        blk = lambda s, n, k: s[n][:, k * B:(k + 1) * B]
        out["X_by_block"] = {n: {"A_vs_A": [d(blk(a1, n, k), blk(a0, n, k)) for k in range(4)],
                                 "B_vs_A": [d(blk(b0, n, k), blk(a0, n, k)) for k in range(4)]} for n in ("X.z", "X.u", "X.h")}
        out["gXh_by_t"] = [d(b0["g.X.h"][t], a0["g.X.h"][t]) for t in range(T)]
        out["gXh_by_t_AA"] = [d(a1["g.X.h"][t], a0["g.X.h"][t]) for t in range(T)]
        out["fake_by_t"] = [d(b0["FAKE"][t * B:(t + 1) * B], a0["FAKE"][t * B:(t + 1) * B]) for t in range(T)]
        out["fake_by_t_AA"] = [d(a1["FAKE"][t * B:(t + 1) * B], a0["FAKE"][t * B:(t + 1) * B]) for t in range(T)]
        # per-timestep view of the alphas and the gate pre-activations (where along the sequence does it start?)
        out["EA_by_t"] = [d(b0["d.EA"][t], a0["d.EA"][t]) for t in range(T)]
        out["EA_by_t_AA"] = [d(a1["d.EA"][t], a0["d.EA"][t]) for t in range(T)]
        out["EA_by_block_t1"] = [d(b0["d.EA"][1, k * B:(k + 1) * B], a0["d.EA"][1, k * B:(k + 1) * B]) for k in range(4)]
        out["Q_by_t"] = [d(b0["d.Q"][t], a0["d.Q"][t]) for t in range(T)]
    path = os.path.join(ROOT, "gpurun_out", os.environ.get("TAG", "r2") + "_step_bisect.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
