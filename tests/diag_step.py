"""Diagnostic (not a pytest): run one D step and one G step on the GPU and print the relative error of
every intermediate against tests/plan_mirror.py, in execution order, to localise a broken kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import plan_mirror as M
from tests.util import make_problem, make_engine, rel

B, T, V, R, E = 8, 3, 96, 196, 300
lam = 10.0
prob = make_problem(B, T, V, R, E)
eng = make_engine(prob, B, T, V, R, E, lam)
m = M.disc_step(prob["gp"], prob["dp"], prob["ann_g"], prob["ann_d"], prob["labels"], prob["noise"], prob["alpha"], lam, T)
eng.disc_step()
torch.cuda.synchronize()
dbg = m["debug"]
RP, VP, EP, KXD, KXG = 256, (V + 63) // 64 * 64, 320, 1344, 1536
NR = 4 * B
def show(name, got, ref):
    print(f"{name:28s} rel={rel(got, ref):.3e}  |ref|={ref.double().norm().item():.3e}")
def hl(x, lo_off, n):
    return x[..., :n].float() + x[..., lo_off:lo_off + n].float()

fake = hl(eng.ws_view("FAKE", (T, B, 2 * VP), torch.bfloat16), VP, V)
show("G fake logits", fake.permute(1, 0, 2), m["fake"])
dP = eng.ws_view("d.P", (B, RP), torch.float32)[:, :R]
show("D P (no bias)", dP, dbg["P"] - prob["dp"]["Discriminator/Discriminator/attention_perceptron/bias"])
X = eng.ws_view("d.X", (T + 1, NR, 2 * KXD), torch.bfloat16)
Cf = eng.ws_view("d.Cf", (T + 1, NR, 512), torch.float32)
EA = eng.ws_view("d.EA", (T, NR, RP), torch.float32)
Q = eng.ws_view("d.Q", (T, NR, 2048), torch.float32)
Y = eng.ws_view("d.Y", (NR, T), torch.float32)
show("D c0", Cf[0, :B], dbg["c0"])
for t in range(T):
    s = dbg["steps"][t]
    xs = hl(X[t], KXD, 1324)
    show(f"t{t} u (x[:,512:812])", xs[:3 * B, 512:812], s["x"][:, 512:812])
    show(f"t{t} alpha", EA[t, :3 * B, :R], s["alpha"])
    show(f"t{t} z", xs[:3 * B, :512], s["x"][:, :512])
    show(f"t{t} h_in", xs[:3 * B, 812:1324], s["x"][:, 812:])
    show(f"t{t} q", Q[t, :3 * B], s["x"] @ M.split_params(prob["dp"], "Discriminator/Discriminator", R, 512)["K"])
    show(f"t{t} c_new", Cf[t + 1, :3 * B], s["cn"])
show("D y", Y[:3 * B], dbg["y"].squeeze(-1))
g = eng.ws_view("DFAKE", (T, B, VP), torch.float32)[:, :, :V]
show("gp gradient g", g.permute(1, 0, 2), m["gp_gradients"])
show("slopes", eng.ws_view("slopes", (B,), torch.float32), m["slopes"])
show("coef", eng.ws_view("coef", (B,), torch.float32), dbg["coef"])
for t in range(T):
    tn = dbg["tans"][t]
    xs = hl(X[t], KXD, 1324)
    show(f"t{t} xdot", xs[3 * B:], tn["x"])
    show(f"t{t} cdot_new", Cf[t + 1, 3 * B:], tn["cn"])
XB = eng.ws_view("d.XB", (T, NR, KXD), torch.float32)
for t in reversed(range(T)):
    show(f"t{t} ubar", XB[t, :3 * B, 512:812], dbg["rv"]["ubar"][t])
    show(f"t{t} udot_bar", XB[t, 3 * B:, 512:812], dbg["rv"]["udot_bar"][t])
show("Pbar", eng.ws_view("d.PB", (B, RP), torch.float32)[:, :R], dbg["rv"]["Pbar"])
sc = eng.scalars.cpu()
print("w_disc", sc[1].item(), float(m["w_disc"]), " gp", sc[2].item(), float(m["gp"]))
gv = eng.d.grad_views()
for k, v in m["grads"].items():
    show("grad " + k.split("/", 1)[1], gv[k], v)

print("---- G step")
mg = M.gen_step(prob["gp"], prob["dp"], prob["ann_g"], prob["ann_d"], prob["noise"], T)
eng.gen_step()
torch.cuda.synchronize()
print("gen_cost", eng.scalars[3].item(), float(mg["gen_cost"]))
df = eng.ws_view("DFAKE", (T, B, VP), torch.float32)[:, :, :V]
show("dfake", df.permute(1, 0, 2), torch.stack(mg["debug"]["dfake"], 1))
gv = eng.g.grad_views()
for k, v in mg["grads"].items():
    show("grad " + k.split("/", 1)[1], gv[k], v)
