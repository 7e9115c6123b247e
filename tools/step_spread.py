#!/usr/bin/env python
"""Run-to-run spread of the D step against the fp64 oracle on the long-sequence parity case (B 3, T 30, V 70, R 24): REPEATS
runs per environment variant (PDL / auxiliary stream / side stream switched off one at a time), per-tensor relative L2
error of every repeat for the attention kernel's two row blocks (W_a, W_h) and the worst other tensor.  A bimodal
distribution that disappears under one of the switches points at an ordering bug; a unimodal one is summation-order noise.
Writes gpurun_out/<TAG>_step_spread.json."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker():
    import torch
    import __graft_entry__ as ge
    ge.build()
    from oracle import sgg_oracle as O
    from tests.util import make_engine, make_problem, rel
    B, T, V, R = (int(x) for x in os.environ.get("CASE", "3,30,70,24").split(","))
    prob = make_problem(B, T, V, R=R, seed=0, dtype=torch.float64)
    eng = make_engine(prob, B, T, V, R=R, lam=10.0)
    ref = O.disc_step_grads(prob["gp"], prob["dp"], prob["ann_g"], prob["ann_d"], prob["real"], prob["noise"], prob["alpha"], 10.0, T)
    key = "Discriminator/Discriminator/attention_perceptron/kernel"
    rows = []
    for _ in range(int(os.environ.get("REPEATS", "40"))):
        eng.disc_step(); torch.cuda.synchronize()
        gv = eng.d.grad_views()
        wa, wh = rel(gv[key][:R * 512], ref["grads"][key][:R * 512]), rel(gv[key][R * 512:], ref["grads"][key][R * 512:])
        other = max(rel(gv[k], v) for k, v in ref["grads"].items() if k != key and not k.endswith("decoder/bias"))
        rows.append([wa, wh, other, float(eng.scalars[2]), float(eng.d.grad.double().sum()), float(eng.ws_view("DFAKE", (T * B, (V + 63) // 64 * 64), torch.float32).double().sum())])
        if os.environ.get("WITH_G", "1") == "1":
            eng.gen_step(); torch.cuda.synchronize()
    print("RESULT " + json.dumps(rows))


def main():
    if os.environ.get("SPREAD_WORKER") == "1":
        return worker()
    out = {}
    variants = {"default": {}, "no_pdl": {"SGG_PDL": "0"}, "no_aux_stream": {"SGG_AUX_STREAM": "0"},
                "no_side_stream": {"SGG_SIDE_STREAM": "0"}, "d_steps_only": {"WITH_G": "0"},
                # no split-K chosen by the cost model: the fp32 reductions of the GEMM outputs keep one order
                "no_split_k": {"SGG_SPLIT_PENALTY": "1e15"}}
    only = os.environ.get("VARIANTS")
    if only:
        variants = {k: v for k, v in variants.items() if k in only.split(",")}
    for name, env in variants.items():
        e = dict(os.environ); e.update(env); e["SPREAD_WORKER"] = "1"
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], capture_output=True, text=True, env=e, timeout=280)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            out[name] = {"error": r.stderr[-500:]}
            continue
        rows = json.loads(line[0][7:])
        col = lambda i: sorted(x[i] for x in rows)
        out[name] = {"W_a_rows": {"min": col(0)[0], "median": col(0)[len(rows) // 2], "max": col(0)[-1], "top3": col(0)[-3:]},
                     "W_h_rows": {"min": col(1)[0], "median": col(1)[len(rows) // 2], "max": col(1)[-1], "top3": col(1)[-3:]},
                     "worst_other_tensor": {"median": col(2)[len(rows) // 2], "max": col(2)[-1]},
                     "gp_min_max": [col(3)[0], col(3)[-1]], "repeats": len(rows),
                     "distinct_gradient_checksums": len(set(x[4] for x in rows)),
                     "distinct_input_gradient_checksums": len(set(x[5] for x in rows))}
    out["case_B_T_V_R"] = os.environ.get("CASE", "3,30,70,24")
    path = os.path.join(ROOT, "gpurun_out", os.environ.get("TAG", "r2") + "_step_spread.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
