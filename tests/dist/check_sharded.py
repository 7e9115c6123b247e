#!/usr/bin/env python
"""Multi-GPU check, run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist/check_sharded.py [--batch 16 --timesteps 3 --vocab 200 --iters 2]

Trains the same shards twice from identical weights, noise and data: once with the row-sharded attention projection
(sgg_wa_shard_t: reduce-scatter of P, all-gather of P_bar, no W_a all-reduce) and once with W_a replicated (gradient
all-reduce), and requires the gathered parameters to agree to fp32 summation-order accuracy.  Also checks both against
a single-process run (world = 1) of the concatenated batch driven step by step with the ranks' noise draws."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-300)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--timesteps", type=int, default=3)
    ap.add_argument("--vocab", type=int, default=200)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--critic-iters", type=int, default=2)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    dist.barrier()
    from sgg_b200.trainer import HotPathTrainer
    g = torch.Generator().manual_seed(100 + rank)
    B, T, V = a.batch, a.timesteps, a.vocab
    batches = [(torch.randn(B, 196, 512, generator=g).bfloat16().cuda(), torch.randn(B, 196, 512, generator=g).bfloat16().cuda(),
                torch.randint(0, V, (B, T), generator=g).cuda()) for _ in range(2)]
    results = {}
    for mode in ("1", "0"):
        os.environ["SGG_WA_SHARD"] = mode
        for use_graph in ((False, True) if mode == "1" else (False,)):
            tr = HotPathTrainer(B, T, V, critic_iters=a.critic_iters, seed=3, use_graph=use_graph)
            assert tr.eng.shard == (mode == "1"), (tr.eng.shard, mode)
            # make the one-sided penalty active so that the second-order path carries signal
            tr.eng.d.views()["Discriminator/W"].mul_(40.0)
            tr.eng.d.refresh_shadow()
            init = ({k: v.clone() for k, v in tr.eng.g.views().items()}, {k: v.clone() for k, v in tr.eng.d.views().items()})
            for i in range(a.iters):
                tr.set_batch(*batches[i % 2])
                tr.iteration()
            torch.cuda.synchronize()
            losses = tr.losses()
            tr.gather_sharded()
            results[(mode, use_graph)] = ({k: v.clone() for k, v in tr.eng.g.views().items()},
                                          {k: v.clone() for k, v in tr.eng.d.views().items()}, losses,
                                          tr.eng.g.m.clone(), tr.eng.d.v.clone())
            tr.close()
            del tr
    ref = results[("0", False)]
    worst, failures = 0.0, []
    for key in (("1", False), ("1", True)):
        got = results[key]
        for net in (0, 1):
            for k in ref[net]:
                # Compare the UPDATES theta - theta_init.  Adam's first steps move every element by ~lr * sign(g), so the
                # few elements whose gradient is ~0 flip under a different fp32 summation order: the tolerance is on the
                # update's relative L2 distance, the Adam moments (linear in the gradients) are compared tightly below.
                du_ref, du_got = ref[net][k] - init[net][k], got[net][k] - init[net][k]
                if du_ref.norm().item() == 0.0:
                    if du_got.norm().item() != 0.0:
                        failures.append((key, k, "update of a frozen tensor"))
                    continue
                e = rel(du_got, du_ref)
                worst = max(worst, e)
                if not e < 5e-2:
                    failures.append((key, k, e))
        for name, idx, tol in (("g.m", 3, 2e-3), ("d.v", 4, 4e-3)):
            e = rel(got[idx], ref[idx])
            if not e < tol:
                failures.append((key, name, e))
        for k in ref[2]:
            if not abs(got[2][k] - ref[2][k]) <= 1e-3 * (abs(ref[2][k]) + 1e-2):
                failures.append((key, k, got[2][k], ref[2][k]))
    if rank == 0 and failures:
        print("check_sharded FAILURES:", *failures, sep="\n  ", flush=True)
    assert not failures, failures[:3]
    # every rank must hold identical parameters after the gather
    flat = torch.cat([v.reshape(-1) for v in results[("1", True)][1].values()])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks disagree on the discriminator parameters"
    if rank == 0:
        print(f"check_sharded ok: world={world} worst relative distance of the parameter updates {worst:.2e}; "
              f"g.m rel {rel(results[('1', True)][3], ref[3]):.2e}; losses {results[('1', True)][2]}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
