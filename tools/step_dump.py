#!/usr/bin/env python
"""Dumps the workspace buffers and gradients of one D step on a small case (CASE=B,T,V,R) to gpurun_out/<TAG>_dump.pt, hi/lo
pairs reconstructed, so that two configurations of the library can be compared buffer by buffer on the host."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    import __graft_entry__ as ge
    ge.build()
    from oracle import sgg_oracle as O
    from tests.util import make_engine, make_problem
    B, T, V, R = (int(x) for x in os.environ.get("CASE", "3,2,70,24").split(","))
    prob = make_problem(B, T, V, R=R, seed=0, dtype=torch.float64)
    eng = make_engine(prob, B, T, V, R=R, lam=10.0)
    eng.disc_step(); torch.cuda.synchronize()
    NR, RP, H, VP, EP = 4 * B, (R + 63) // 64 * 64, 512, (V + 63) // 64 * 64, 320
    KXP = 1344
    f = lambda n, shp, dt: eng.ws_view(n, shp, dt).float().cpu()
    hl = lambda x, w: x[..., :w] + x[..., w:]
    out = {"d.Y": f("d.Y", (NR, T), torch.float32), "d.EA": f("d.EA", (T, NR, RP), torch.float32), "d.Q": f("d.Q", (T, NR, 4 * H), torch.float32),
           "DFAKE": f("DFAKE", (T * B, VP), torch.float32), "slopes": f("slopes", (B,), torch.float32), "coef": f("coef", (B,), torch.float32),
           "VHL": hl(f("VHL", (T * B, 2 * VP), torch.bfloat16), VP), "d.ED": f("d.ED", (T, B, RP), torch.float32),
           "d.QB": hl(f("d.QB", (T, NR, 8 * H), torch.bfloat16), 4 * H), "d.XB": f("d.XB", (T, NR, KXP), torch.float32),
           "d.EB": hl(f("d.EB", (T, NR, 2 * RP), torch.bfloat16), RP), "d.CB": f("d.CB", (T, NR, H), torch.float32),
           "d.PB": f("d.PB", (B, RP), torch.float32), "d.X": hl(f("d.X", (T + 1, NR, 2 * KXP), torch.bfloat16), KXP),
           "d.CH": hl(f("d.CH", (T + 1, NR, 2 * H), torch.bfloat16), H), "d.Cf": f("d.Cf", (T + 1, NR, H), torch.float32),
           "UF": f("UF", (T * B, EP), torch.float32), "UBH": hl(f("UBH", (T * B, 2 * EP), torch.bfloat16), EP),
           "UDB": hl(f("UDB", (T * B, 2 * EP), torch.bfloat16), EP), "scalars": eng.scalars.cpu()}
    for k, v in eng.d.grad_views().items():
        out["grad/" + k] = v.float().cpu().clone()
    ref = O.disc_step_grads(prob["gp"], prob["dp"], prob["ann_g"], prob["ann_d"], prob["real"], prob["noise"], prob["alpha"], 10.0, T)
    for k, v in ref["grads"].items():
        out["ref/" + k] = v.float()
    torch.save(out, os.path.join(ROOT, "gpurun_out", os.environ.get("TAG", "r2") + "_dump.pt"))
    print("ok")


if __name__ == "__main__":
    main()
