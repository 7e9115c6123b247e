"""World-size-2 gloo test (CPU) of the data-parallel scheme: batch shards, losses normalised by the
GLOBAL batch, one sum-allreduce of the gradient bucket per optimiser step == the single-process step on the
concatenated batch.  The per-rank arithmetic is the oracle's (the CUDA kernels normalise the same way through
sgg_step_args_t.world); the host plumbing under test is sgg_b200/dp.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sgg_oracle as O

B, T, V, R, LAM = 4, 3, 9, 6, 10.0


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    gp = O.init_generator_params(V, seed=3, R=R, dtype=torch.float64)
    dp_ = O.init_discriminator_params(V, seed=4, R=R, dtype=torch.float64)
    ann_g, ann_d, labels, real = O.synthetic_batch(B, V, T, R, 512, seed=5, dtype=torch.float64)
    g = torch.Generator().manual_seed(6)
    noise = torch.randn(B, 512, generator=g, dtype=torch.float64)
    # make the one-sided penalty active for some samples: scale the embedding up
    dp_["Discriminator/W"] = dp_["Discriminator/W"] * 40.0
    alpha = torch.rand(B, generator=g, dtype=torch.float64)
    return gp, dp_, ann_g, ann_d, real, noise, alpha


def _worker(rank: int, world: int, port: int, out_dir: str):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sgg_b200 import dp
        gp, dp_, ann_g, ann_d, real, noise, alpha = _problem()
        lo, hi = dp.shard_bounds(B, rank, world)
        assert hi - lo == B // world
        ident = dp.broadcast_comm_id(dist, None, rank, lambda: bytes(range(128)))
        assert ident == bytes(range(128))
        assert dp.rank_seed(0, 0) != dp.rank_seed(0, 1)
        sl = slice(lo, hi)
        # local step on the shard: the oracle's means are over the LOCAL batch -> rescale to the global batch
        rd = O.disc_step_grads(gp, dp_, ann_g[sl], ann_d[sl], real[sl], noise[sl], alpha[sl], LAM, T)
        rg = O.gen_step_grads(gp, dp_, ann_g[sl], ann_d[sl], noise[sl], T)
        res = {}
        for name, r in (("d", rd), ("g", rg)):
            bucket = torch.cat([g.reshape(-1) for g in r["grads"].values()]) / world
            dist.all_reduce(bucket, op=dist.ReduceOp.SUM)          # the one exchange of the step
            res[name] = bucket
        sc = torch.stack([rd["w_disc"], rd["gp"], rg["gen_cost"]]) / world
        dist.all_reduce(sc, op=dist.ReduceOp.SUM)
        res["scalars"] = sc
        torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_step_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    gp, dp_, ann_g, ann_d, real, noise, alpha = _problem()
    rd = O.disc_step_grads(gp, dp_, ann_g, ann_d, real, noise, alpha, LAM, T)
    rg = O.gen_step_grads(gp, dp_, ann_g, ann_d, noise, T)
    assert float(rd["gp"]) > 0, "the test problem must exercise the penalty"
    full_d = torch.cat([g.reshape(-1) for g in rd["grads"].values()])
    full_g = torch.cat([g.reshape(-1) for g in rg["grads"].values()])
    full_sc = torch.stack([rd["w_disc"], rd["gp"], rg["gen_cost"]])
    for r in range(world):
        res = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))
        assert torch.allclose(res["d"], full_d, rtol=1e-9, atol=1e-12)
        assert torch.allclose(res["g"], full_g, rtol=1e-9, atol=1e-12)
        assert torch.allclose(res["scalars"], full_sc, rtol=1e-10, atol=1e-13)


def test_shard_bounds_reject_ragged_batches():
    from sgg_b200 import dp
    import pytest
    assert dp.shard_bounds(2048, 3, 8) == (768, 1024)
    with pytest.raises(ValueError):
        dp.shard_bounds(10, 0, 4)
    with pytest.raises(ValueError):
        dp.shard_bounds(8, 4, 4)


def _shard_worker(rank: int, world: int, port: int, out_dir: str):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sgg_b200 import dp
        Bl, Rr, Cc, N = 3, 4, 64, 5            # local batch, regions, channels (one 64-wide k-block per region), outputs
        g = torch.Generator().manual_seed(11)
        a_all = torch.randn(world * Bl, Rr * Cc, generator=g, dtype=torch.float64)      # every rank's annotations
        Wa = torch.randn(Rr * Cc, N, generator=g, dtype=torch.float64)
        Pbar_all = torch.randn(world * Bl, N, generator=g, dtype=torch.float64)
        a_loc = a_all[rank * Bl:(rank + 1) * Bl]
        r0, r1 = dp.wa_shard_rows(Rr, Cc, rank, world)
        # all-to-all of column slabs (emulated with all_gather: gloo has no all_to_all on CPU tensors in every build)
        send = dp.slab_send_layout(a_loc, world)                                        # [world, Bl, Ks]
        gathered = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(gathered, send)
        slab = torch.cat([gathered[p][rank] for p in range(world)], dim=0)              # [world*Bl, Ks]
        assert torch.equal(slab, a_all[:, r0:r1])
        # K1: partial projection of the GLOBAL batch over this rank's rows, reduce-scatter (all_reduce + slice here)
        part = slab @ Wa[r0:r1]
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
        P_loc = part[rank * Bl:(rank + 1) * Bl]
        assert torch.allclose(P_loc, a_loc @ Wa, rtol=1e-12, atol=1e-12)
        # dW_a rows of this rank: all-gather of P_bar, contraction over the global batch -> equals the all-reduced gradient
        pb = [torch.empty(Bl, N, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(pb, Pbar_all[rank * Bl:(rank + 1) * Bl].contiguous())
        dWa_rows = slab.t() @ torch.cat(pb, dim=0)
        full = a_all.t() @ Pbar_all
        assert torch.allclose(dWa_rows, full[r0:r1], rtol=1e-12, atol=1e-12)
        torch.save({"ok": True}, os.path.join(out_dir, f"shard{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_row_sharded_projection_algebra_world2(tmp_path):
    """The exchange pattern of sgg_wa_shard_t (slab all-to-all, reduce-scatter of P, all-gather of P_bar) reproduces the
    replicated computation: host-side layout helpers of dp.py under gloo, world_size 2."""
    world = 2
    mp.spawn(_shard_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert torch.load(os.path.join(str(tmp_path), f"shard{r}.pt"))["ok"]


def test_wa_shard_rows_cover_the_contraction():
    from sgg_b200 import dp
    import pytest
    for world in (1, 2, 4, 8):
        rows = [dp.wa_shard_rows(196, 512, r, world) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == 196 * 512
        assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
        assert all((b - a) % 64 == 0 for a, b in rows)
    with pytest.raises(ValueError):
        dp.wa_shard_rows(196, 512, 0, 3)
