"""Host-side mirror of the CUDA execution plan (TEST INFRASTRUCTURE, no autograd).

The CUDA path does not run the reference's graph literally: it hoists the loop-invariant
attention projection ``P = flat(a) W_a + b``, batches the discriminator's passes
(fake / real / interpolate) over one read of the annotations, and replaces the
"gradient of a gradient" of the WGAN-GP penalty by *reverse over a forward tangent*:

    d pen / d theta  =  d/d theta < grad_x D(x; theta), v >,   v = d pen / d g  held constant
                     =  d/d theta  [ d/d eps  sum D(x + eps v; theta) ]_{eps = 0}

so the double backward becomes: forward, input-gradient pass, tangent (JVP) forward, and
ONE reverse pass over the primal+tangent program.  Every function below corresponds to
one CUDA kernel (or one fused kernel family) in ``sgg_b200/csrc``; the unit
tests check each formula here against autograd on the oracle in fp64, and the GPU tests
check each kernel against the matching function here.

Shapes: a [B,R,C]; streams are stacked on the row axis (row n = s*B + b).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

LN_EPS = 1e-12
FORGET_BIAS = 1.0
GP_EPS = 1e-10


def split_params(p: Dict[str, torch.Tensor], prefix: str, R: int, C: int):
    """Row-split the reference's concat kernels: W_att -> (W_a, W_h),
    cell kernel -> (K_z, K_u, K_h)."""
    H = p[f"{prefix}/layer_norm_basic_lstm_cell/state/gamma"].shape[0]
    W_att = p[f"{prefix}/attention_perceptron/kernel"]
    K = p[f"{prefix}/layer_norm_basic_lstm_cell/kernel"]
    U = K.shape[0] - C - H
    out = {
        "Wa": W_att[: R * C], "Wh": W_att[R * C:], "b_att": p[f"{prefix}/attention_perceptron/bias"],
        "K": K, "Kz": K[:C], "Ku": K[C:C + U], "Kh": K[C + U:],
        "Wdec": p[f"{prefix}/decoder/kernel"], "bdec": p[f"{prefix}/decoder/bias"],
    }
    for n in ("input", "transform", "forget", "output", "state"):
        out[f"g_{n}"] = p[f"{prefix}/layer_norm_basic_lstm_cell/{n}/gamma"]
        out[f"b_{n}"] = p[f"{prefix}/layer_norm_basic_lstm_cell/{n}/beta"]
    return out


# ------------------------------------------------------------------ primitives
def ln_fwd(x):
    mu = x.mean(-1, keepdim=True)
    xc = x - mu
    r = torch.rsqrt((xc * xc).mean(-1, keepdim=True) + LN_EPS)
    return xc * r, r


def ln_proj(n, r, w):
    """J w with J = r (I - 11^T/N - n n^T/N): both the LN JVP and the LN VJP (J is symmetric)."""
    return r * (w - w.mean(-1, keepdim=True) - n * (n * w).mean(-1, keepdim=True))


def ln_rev(n, r, nbar, xdot=None, ndot=None, ndot_bar=None):
    """Reverse of (n = LNnorm(x), ndot = J(x) xdot) given adjoints nbar, ndot_bar.
    Returns (xbar, xdot_bar)."""
    xbar = ln_proj(n, r, nbar)
    if xdot is None:
        return xbar, None
    N = n.shape[-1]
    xdot_bar = ln_proj(n, r, ndot_bar)
    p = (n * ndot_bar).sum(-1, keepdim=True)
    q = (n * xdot).sum(-1, keepdim=True)
    s = (ndot_bar * ndot).sum(-1, keepdim=True)
    xbar = xbar - (r / N) * (n * s + q * xdot_bar + p * ndot)
    return xbar, xdot_bar


def softmax_rev(alpha, abar, edot=None, adot_bar=None):
    """Reverse of (alpha = softmax(e), adot = J edot), J = diag(alpha) - alpha alpha^T."""
    if edot is None:
        return alpha * (abar - (alpha * abar).sum(-1, keepdim=True)), None
    m_e = (alpha * edot).sum(-1, keepdim=True)
    m_t = (alpha * adot_bar).sum(-1, keepdim=True)
    edot_bar = alpha * (adot_bar - m_t)
    w = abar + adot_bar * (edot - m_e) - edot * m_t
    ebar = alpha * (w - (alpha * w).sum(-1, keepdim=True))
    return ebar, edot_bar


# ------------------------------------------------------------------ kernels (forward)
def attn_proj(a, w):
    """K1: P = flat(a) W_a + b (hoisted, once per network per step) and c0 = h0 = mean_r a."""
    B = a.shape[0]
    return a.reshape(B, -1) @ w["Wa"] + w["b_att"], a.mean(1)


def attn_step_fwd(a, P, c, w, S):
    """K2: e = P + c W_h ; alpha = softmax(e) ; z = sum_r alpha_r a_r.  rows n = s*B+b."""
    e = P.repeat(S, 1) + c @ w["Wh"]
    alpha = torch.softmax(e, -1)
    B = a.shape[0]
    z = torch.einsum("sbr,brc->sbc", alpha.reshape(S, B, -1), a).reshape(S * B, -1)
    return alpha, z


def attn_step_tan(a, alpha, cdot, w):
    """K2 tangent (interp stream only): edot = cdot W_h ; adot = J edot ; zdot = sum adot a."""
    edot = cdot @ w["Wh"]
    adot = alpha * (edot - (alpha * edot).sum(-1, keepdim=True))
    zdot = torch.einsum("br,brc->bc", adot, a)
    return edot, adot, zdot


def lstm_fwd(q, c, w):
    """K3 epilogue: 4x LN, gates, cell update, LN(c'), h.  Returns saves dict."""
    H = c.shape[-1]
    s = {"c_in": c}
    acts = {}
    for idx, (g, name) in enumerate(zip("ijfo", ("input", "transform", "forget", "output"))):
        n, r = ln_fwd(q[:, idx * H:(idx + 1) * H])
        s[f"n{g}"], s[f"r{g}"] = n, r
        acts[g] = n * w[f"g_{name}"] + w[f"b_{name}"]
    s["si"] = torch.sigmoid(acts["i"])
    s["tj"] = torch.tanh(acts["j"])
    s["sf"] = torch.sigmoid(acts["f"] + FORGET_BIAS)
    s["so"] = torch.sigmoid(acts["o"])
    cp = c * s["sf"] + s["si"] * s["tj"]
    s["nc"], s["rc"] = ln_fwd(cp)
    cn = s["nc"] * w["g_state"] + w["b_state"]
    s["tc"] = torch.tanh(cn)
    s["cn"] = cn
    s["h"] = s["tc"] * s["so"]
    return s


def lstm_tan(s, qdot, cdot, w):
    """K3 tangent: JVP of lstm_fwd along (qdot, cdot)."""
    H = cdot.shape[-1]
    t = {"c_in": cdot}
    ad = {}
    for idx, (g, name) in enumerate(zip("ijfo", ("input", "transform", "forget", "output"))):
        t[f"q{g}"] = qdot[:, idx * H:(idx + 1) * H]
        t[f"n{g}"] = ln_proj(s[f"n{g}"], s[f"r{g}"], t[f"q{g}"])
        ad[g] = t[f"n{g}"] * w[f"g_{name}"]
    t["ai"], t["aj"], t["af"], t["ao"] = ad["i"], ad["j"], ad["f"], ad["o"]
    t["si"] = s["si"] * (1 - s["si"]) * ad["i"]
    t["tj"] = (1 - s["tj"] ** 2) * ad["j"]
    t["sf"] = s["sf"] * (1 - s["sf"]) * ad["f"]
    t["so"] = s["so"] * (1 - s["so"]) * ad["o"]
    t["cp"] = cdot * s["sf"] + s["c_in"] * t["sf"] + t["si"] * s["tj"] + s["si"] * t["tj"]
    t["nc"] = ln_proj(s["nc"], s["rc"], t["cp"])
    t["cn"] = t["nc"] * w["g_state"]
    t["tc"] = (1 - s["tc"] ** 2) * t["cn"]
    t["h"] = t["tc"] * s["so"] + s["tc"] * t["so"]
    return t


def lstm_rev(s, w, hbar, cnbar, t=None, hdot_bar=None, cndot_bar=None):
    """K3 reverse (optionally over primal+tangent).  Returns dict with qbar [.,4H], cbar,
    (qdot_bar, cdot_bar) and the LN gamma/beta gradients (summed over rows)."""
    H = hbar.shape[-1]
    so, tc, si, tj, sf = s["so"], s["tc"], s["si"], s["tj"], s["sf"]
    dt = 1 - tc * tc
    tan = t is not None
    # h = tc*so
    tcbar = hbar * so
    sobar = hbar * tc
    if tan:
        tcbar = tcbar + hdot_bar * t["so"]
        sobar = sobar + hdot_bar * t["tc"]
        tcdot_bar = hdot_bar * so
        sodot_bar = hdot_bar * tc
    # tc = tanh(cn)
    cnb = cnbar + tcbar * dt
    if tan:
        cnb = cnb + tcdot_bar * (-2 * tc * dt) * t["cn"]
        cndb = cndot_bar + tcdot_bar * dt
    # cn = g*nc + b
    grads = {}
    ncbar = cnb * w["g_state"]
    grads["g_state"] = (cnb * s["nc"]).sum(0)
    grads["b_state"] = cnb.sum(0)
    if tan:
        ncdot_bar = cndb * w["g_state"]
        grads["g_state"] = grads["g_state"] + (cndb * t["nc"]).sum(0)
        cpbar, cpdot_bar = ln_rev(s["nc"], s["rc"], ncbar, t["cp"], t["nc"], ncdot_bar)
    else:
        cpbar, _ = ln_rev(s["nc"], s["rc"], ncbar)
    # cp = c*sf + si*tj
    c = s["c_in"]
    cbar = cpbar * sf
    sfbar = cpbar * c
    sibar = cpbar * tj
    tjbar = cpbar * si
    if tan:
        cbar = cbar + cpdot_bar * t["sf"]
        sfbar = sfbar + cpdot_bar * t["c_in"]
        sibar = sibar + cpdot_bar * t["tj"]
        tjbar = tjbar + cpdot_bar * t["si"]
        cdot_bar = cpdot_bar * sf
        sfdot_bar = cpdot_bar * c
        sidot_bar = cpdot_bar * tj
        tjdot_bar = cpdot_bar * si
    # nonlinearities
    d = {"i": si * (1 - si), "j": 1 - tj * tj, "f": sf * (1 - sf), "o": so * (1 - so)}
    d2 = {"i": d["i"] * (1 - 2 * si), "j": -2 * tj * d["j"], "f": d["f"] * (1 - 2 * sf), "o": d["o"] * (1 - 2 * so)}
    ybar = {"i": sibar, "j": tjbar, "f": sfbar, "o": sobar}
    if tan:
        ydot_bar = {"i": sidot_bar, "j": tjdot_bar, "f": sfdot_bar, "o": sodot_bar}
    qbar = torch.empty(hbar.shape[0], 4 * H, dtype=hbar.dtype)
    qdot_bar = torch.empty_like(qbar) if tan else None
    for idx, (g, name) in enumerate(zip("ijfo", ("input", "transform", "forget", "output"))):
        abar = ybar[g] * d[g]
        if tan:
            abar = abar + ydot_bar[g] * d2[g] * t[f"a{g}"]
            adot_bar = ydot_bar[g] * d[g]
        grads[f"g_{name}"] = (abar * s[f"n{g}"]).sum(0)
        grads[f"b_{name}"] = abar.sum(0)
        nbar = abar * w[f"g_{name}"]
        if tan:
            grads[f"g_{name}"] = grads[f"g_{name}"] + (adot_bar * t[f"n{g}"]).sum(0)
            xb, xdb = ln_rev(s[f"n{g}"], s[f"r{g}"], nbar, t[f"q{g}"], t[f"n{g}"], adot_bar * w[f"g_{name}"])
            qdot_bar[:, idx * H:(idx + 1) * H] = xdb
        else:
            xb, _ = ln_rev(s[f"n{g}"], s[f"r{g}"], nbar)
        qbar[:, idx * H:(idx + 1) * H] = xb
    out = {"qbar": qbar, "cbar": cbar, "grads": grads}
    if tan:
        out["qdot_bar"] = qdot_bar
        out["cdot_bar"] = cdot_bar
    return out


def attn_step_rev(a, alpha, zbar, S, edot=None, zdot_bar=None):
    """K2 reverse: abar_r = <zbar, a_r>; softmax reverse -> ebar (and edot_bar)."""
    B = a.shape[0]
    abar = torch.einsum("sbc,brc->sbr", zbar.reshape(S, B, -1), a).reshape(S * B, -1)
    if edot is None:
        return softmax_rev(alpha, abar)
    adot_bar = torch.einsum("bc,brc->br", zdot_bar, a)
    return softmax_rev(alpha, abar, edot, adot_bar)


# ------------------------------------------------------------------ network passes
def net_forward(a, w, u_list, S, P=None, c0=None):
    """Forward of the recurrent half for S stacked streams.  u_list[t]: [S*B, U] (noise for
    G, embeddings for D).  Returns y [S*B, T, out] and per-step saves."""
    if P is None:
        P, c0 = attn_proj(a, w)
    c = c0.repeat(S, 1)
    h = c
    steps, ys = [], []
    for u in u_list:
        alpha, z = attn_step_fwd(a, P, c, w, S)
        x = torch.cat([z, u, h], 1)
        q = x @ w["K"]
        s = lstm_fwd(q, c, w)
        s.update({"alpha": alpha, "x": x, "h_in": h})
        steps.append(s)
        c, h = s["cn"], s["h"]
        ys.append(h @ w["Wdec"] + w["bdec"])
    return torch.stack(ys, 1), steps, P, c0


def net_tangent(a, w, steps, udot_list):
    """Tangent forward (one stream) along input-space direction: udot_list[t] = v_t W_emb."""
    B = a.shape[0]
    H = w["Wh"].shape[0]
    cdot = torch.zeros(B, H, dtype=a.dtype)
    hdot = torch.zeros(B, H, dtype=a.dtype)
    tans = []
    for s, udot in zip(steps, udot_list):
        edot, adot, zdot = attn_step_tan(a, s["alpha"], cdot, w)
        xdot = torch.cat([zdot, udot, hdot], 1)
        t = lstm_tan(s, xdot @ w["K"], cdot, w)
        t.update({"edot": edot, "adot": adot, "x": xdot})
        tans.append(t)
        cdot, hdot = t["cn"], t["h"]
    return tans


def net_reverse(a, w, steps, S, ybar_list, tans=None, ydot_bar_list=None, tan_rows=None,
                need_weight_grads=True, C=None):
    """Reverse pass over T steps for S stacked streams.  Tangent data (if any) belongs to the
    rows `tan_rows` (a slice).  Returns dict: ubar[t] (adjoint of u_t), udot_bar[t], grads, and `abar`
    [B,R,C], the adjoint of the annotations (what the conv front-end gen:29-68 back-propagates): the tile enters
    through z_t = sum_r alpha_r a_r (and zdot_t = sum_r adot_r a_r), through P = flat(a) W_a and through
    c0 = h0 = mean_r a_r; the tangent's start state is the constant 0."""
    B, R, Cc = a.shape
    H = w["Wh"].shape[0]
    T = len(steps)
    n = S * B
    dt = a.dtype
    U = w["K"].shape[0] - Cc - H
    hbar = torch.zeros(n, H, dtype=dt)
    cnbar = torch.zeros(n, H, dtype=dt)
    hdb = cndb = None
    if tans is not None:
        hdb = torch.zeros(B, H, dtype=dt)
        cndb = torch.zeros(B, H, dtype=dt)
    G = {k: torch.zeros_like(v) for k, v in w.items() if k not in ("Kz", "Ku", "Kh")}
    Pbar = torch.zeros(B, R, dtype=dt)
    abar = torch.zeros(B, R, Cc, dtype=dt)
    ubar, udot_bar = [None] * T, [None] * T
    for ti in reversed(range(T)):
        s = steps[ti]
        ybar = ybar_list[ti]
        hb = hbar + ybar @ w["Wdec"].T
        if need_weight_grads:
            G["Wdec"] += s["h"].T @ ybar
            G["bdec"] += ybar.sum(0)
        # ---- LSTM pointwise reverse (per stream group: rows with tangents vs without)
        if tans is not None:
            t = tans[ti]
            ydb = ydot_bar_list[ti]
            hdb_t = hdb + ydb @ w["Wdec"].T
            if need_weight_grads:
                G["Wdec"] += t["h"].T @ ydb
            rows = tan_rows
            out_t = lstm_rev({k: v[rows] for k, v in s.items() if torch.is_tensor(v) and v.shape[0] == n},
                             w, hb[rows], cnbar[rows], t, hdb_t, cndb)
            qbar = torch.zeros(n, 4 * H, dtype=dt)
            cbar = torch.zeros(n, H, dtype=dt)
            other = torch.ones(n, dtype=torch.bool)
            other[rows] = False
            lg = dict(out_t["grads"])
            if other.any():
                out_o = lstm_rev({k: v[other] for k, v in s.items() if torch.is_tensor(v) and v.shape[0] == n},
                                 w, hb[other], cnbar[other])
                qbar[other] = out_o["qbar"]
                cbar[other] = out_o["cbar"]
                for k in lg:
                    lg[k] = lg[k] + out_o["grads"][k]
            qbar[rows] = out_t["qbar"]
            cbar[rows] = out_t["cbar"]
            qdb, cdb = out_t["qdot_bar"], out_t["cdot_bar"]
        else:
            out = lstm_rev(s, w, hb, cnbar)
            qbar, cbar, lg = out["qbar"], out["cbar"], out["grads"]
        if need_weight_grads:
            for k, v in lg.items():
                G[k] += v
            G["K"] += s["x"].T @ qbar
        xbar = qbar @ w["K"].T
        zbar, ubar[ti], hbar = xbar[:, :Cc], xbar[:, Cc:Cc + U], xbar[:, Cc + U:]
        if tans is not None:
            if need_weight_grads:
                G["K"] += t["x"].T @ qdb
            xdb = qdb @ w["K"].T
            zdb, udot_bar[ti], hdb = xdb[:, :Cc], xdb[:, Cc:Cc + U], xdb[:, Cc + U:]
            # ---- attention reverse
            ebar_o, _ = attn_step_rev(a, s["alpha"], zbar, S)          # first-order everywhere
            ebar_t, edb = softmax_rev(s["alpha"][rows],
                                      torch.einsum("bc,brc->br", zbar[rows], a),
                                      t["edot"], torch.einsum("bc,brc->br", zdb, a))
            ebar = ebar_o.clone()
            ebar[rows] = ebar_t
            cndb = cdb + edb @ w["Wh"].T
            if need_weight_grads:
                G["Wh"] += t["c_in"].T @ edb
        else:
            ebar, _ = attn_step_rev(a, s["alpha"], zbar, S)
        cnbar = cbar + ebar @ w["Wh"].T
        if need_weight_grads:
            G["Wh"] += s["c_in"].T @ ebar
            Pbar += ebar.reshape(S, B, R).sum(0)
            # ann_grad kernel: rank-(T * streams) outer-product sum over the saved alpha and the z_bar rows
            abar += torch.einsum("sbr,sbc->brc", s["alpha"].reshape(S, B, R), zbar.reshape(S, B, Cc))
            if tans is not None and ti > 0:       # adot_0 = 0
                abar += torch.einsum("br,bc->brc", t["adot"], zdb)
    if need_weight_grads:
        G["Wa"] = a.reshape(B, -1).T @ Pbar
        G["b_att"] = Pbar.sum(0)
        abar += (Pbar @ w["Wa"].T).reshape(B, R, Cc)                                   # K1 reverse
        abar += ((cnbar + hbar).reshape(S, B, H).sum(0) / R).reshape(B, 1, Cc)         # c0 = h0 = mean_r a_r
    return {"ubar": ubar, "udot_bar": udot_bar, "grads": G, "Pbar": Pbar, "abar": abar}


def pack_grads(G, prefix):
    """Re-assemble the reference's variable layout from the split gradients."""
    out = {
        f"{prefix}/attention_perceptron/kernel": torch.cat([G["Wa"], G["Wh"]], 0),
        f"{prefix}/attention_perceptron/bias": G["b_att"],
        f"{prefix}/layer_norm_basic_lstm_cell/kernel": G["K"],
        f"{prefix}/decoder/kernel": G["Wdec"],
        f"{prefix}/decoder/bias": G["bdec"],
    }
    for n in ("input", "transform", "forget", "output", "state"):
        out[f"{prefix}/layer_norm_basic_lstm_cell/{n}/gamma"] = G[f"g_{n}"]
        out[f"{prefix}/layer_norm_basic_lstm_cell/{n}/beta"] = G[f"b_{n}"]
    return out


# ------------------------------------------------------------------ full steps
def gen_forward(gp, ann_g, noise, T):
    R, C = ann_g.shape[1], ann_g.shape[2]
    wg = split_params(gp, "Generator/Generator", R, C)
    y, steps, P, c0 = net_forward(ann_g, wg, [noise] * T, 1)
    return y, steps, wg


def disc_step(gp, dp, ann_g, ann_d, labels, noise, gp_alpha, lam, T):
    """The D step as the CUDA plan runs it.  labels [B,T] int64 (one-hot reals)."""
    B, R, C = ann_d.shape
    V = dp["Discriminator/W"].shape[0]
    Wemb = dp["Discriminator/W"]
    wd = split_params(dp, "Discriminator/Discriminator", R, C)
    fake, _, _ = gen_forward(gp, ann_g, noise, T)                       # [B,T,V] constant
    al = gp_alpha.reshape(B, 1)
    # embeddings: fake dense GEMM, real gather, interp by linearity
    u_fake = [fake[:, t] @ Wemb for t in range(T)]
    u_real = [Wemb[labels[:, t]] for t in range(T)]
    u_int = [ur + al * (uf - ur) for uf, ur in zip(u_fake, u_real)]
    u_list = [torch.cat([uf, ur, ui], 0) for uf, ur, ui in zip(u_fake, u_real, u_int)]
    y, steps, P, c0 = net_forward(ann_d, wd, u_list, 3)
    d_fake, d_real = y[:B], y[B:2 * B]
    w_disc = d_fake.mean() - d_real.mean()
    rows = slice(2 * B, 3 * B)
    # ---- input-gradient pass on the interp stream (data path only)
    st_int = [{k: (v[rows] if torch.is_tensor(v) and v.shape[0] == 3 * B else v) for k, v in s.items()} for s in steps]
    one = [torch.ones(B, 1, dtype=ann_d.dtype)] * T
    ig = net_reverse(ann_d, wd, st_int, 1, one, need_weight_grads=False)
    g = torch.stack([ub @ Wemb.T for ub in ig["ubar"]], 1)               # [B,T,V]
    slopes = torch.sqrt((g * g).sum((1, 2)) + GP_EPS)
    pen = (torch.clamp(slopes - 1.0, min=0) ** 2).mean()
    coef = (2.0 / B) * torch.clamp(slopes - 1.0, min=0) / slopes          # d pen / d g = coef * g
    v = coef.reshape(B, 1, 1) * g
    # ---- tangent forward along v
    udot = [v[:, t] @ Wemb for t in range(T)]
    tans = net_tangent(ann_d, wd, st_int, udot)
    # ---- one reverse over everything
    n = 3 * B
    ybar = torch.zeros(n, 1, dtype=ann_d.dtype)
    ybar[:B] = 1.0 / (B * T)
    ybar[B:2 * B] = -1.0 / (B * T)
    ydb = torch.full((B, 1), float(lam), dtype=ann_d.dtype)
    rv = net_reverse(ann_d, wd, steps, 3, [ybar] * T, tans, [ydb] * T, rows)
    grads = pack_grads(rv["grads"], "Discriminator/Discriminator")
    # embedding gradient: fake^T (ubar_f + al*ubar_i) + scatter(labels, ubar_r + (1-al) ubar_i) + v^T udot_bar
    gW = torch.zeros_like(Wemb)
    for t in range(T):
        ub = rv["ubar"][t]
        uf, ur, ui = ub[:B], ub[B:2 * B], ub[2 * B:]
        gW += fake[:, t].T @ (uf + al * ui)
        gW.index_add_(0, labels[:, t], ur + (1 - al) * ui)
        gW += v[:, t].T @ rv["udot_bar"][t]
    grads["Discriminator/W"] = gW
    return {"disc_cost": w_disc + lam * pen, "w_disc": w_disc, "gp": pen, "slopes": slopes,
            "gp_gradients": g, "grads": grads, "fake": fake, "ann_grad": rv["abar"],
            "debug": {"steps": steps, "tans": tans, "rv": rv, "ig": ig, "P": P, "c0": c0, "y": y, "coef": coef,
                      "v": v, "u_list": u_list, "udot": udot}}


def gen_step(gp, dp, ann_g, ann_d, noise, T):
    """The G step as the CUDA plan runs it."""
    B, R, C = ann_d.shape
    Wemb = dp["Discriminator/W"]
    wd = split_params(dp, "Discriminator/Discriminator", R, C)
    fake, gsteps, wg = gen_forward(gp, ann_g, noise, T)
    u_list = [fake[:, t] @ Wemb for t in range(T)]
    y, dsteps, _, _ = net_forward(ann_d, wd, u_list, 1)
    gen_cost = -y.mean()
    ybar = torch.full((B, 1), -1.0 / (B * T), dtype=ann_d.dtype)
    rd = net_reverse(ann_d, wd, dsteps, 1, [ybar] * T, need_weight_grads=False)
    dfake = [ub @ Wemb.T for ub in rd["ubar"]]                           # [B,V] per t
    rg = net_reverse(ann_g, wg, gsteps, 1, dfake)
    return {"gen_cost": gen_cost, "grads": pack_grads(rg["grads"], "Generator/Generator"), "fake": fake,
            "ann_grad": rg["abar"],
            "debug": {"gsteps": gsteps, "dsteps": dsteps, "rd": rd, "rg": rg, "dfake": dfake, "y": y}}
