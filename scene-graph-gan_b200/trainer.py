"""WGAN-GP training schedule of the reference (train.py:362-368) over the CUDA step engine.

One iteration = ``critic_iters`` discriminator steps followed by one generator step, all on the
same data batch (train.py:185-187 repeats each batch CRITIC_ITERS+1 times), fresh noise
(gen:81) and fresh interpolation coefficients (tfgan gradient penalty) per step, two
tf.train.AdamOptimizer(1e-4, beta1=0.5, beta2=0.9) updates (train.py:258-266).

Data parallelism (the reference has none): one process per GPU, the batch is sharded
over ranks, the kernels normalise every loss by the GLOBAL batch, so the only exchange is one
sum-allreduce of the flat gradient bucket per optimiser step (NCCL over NVLink/NVSwitch).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Tuple

import torch

from .engine import Engine


class HotPathTrainer:
    def __init__(self, batch_size: int, n_steps: int = 3, vocab_size: int = 2000, critic_iters: int = 5,
                 lam: float = 10.0, regions: int = 196, embed_dim: int = 300, seed: int = 0,
                 embedding: Optional[torch.Tensor] = None, process_group=None, device=None):
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.pg = process_group
        self.world = self.dist.get_world_size(process_group) if self.dist else 1
        self.rank = self.dist.get_rank(process_group) if self.dist else 0
        self.B, self.T, self.V, self.R = batch_size, n_steps, vocab_size, regions
        self.critic_iters, self.lam = int(critic_iters), float(lam)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        # identical initial weights on every rank (same seed); decorrelated noise / alpha streams per rank
        self.eng = Engine(batch_size, n_steps, vocab_size, regions, embed_dim, lam=lam, world=self.world,
                          seed=(seed * 1000003 + 7919 * self.rank) & 0x7FFFFFFF, device=self.device)
        self.eng.g.init_reference(seed * 2 + 1)
        self.eng.d.init_reference(seed * 2 + 2, embedding=embedding)
        self.iterations = 0
        # double-buffered device staging for host batches
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._slots = [None, None]
        self._slot_ready = [None, None]
        self._cur = 0
        self._loss_host = torch.zeros(4, dtype=torch.float32).pin_memory()
        self.h2d_bytes_per_batch = 2 * batch_size * regions * 512 * 2 + batch_size * n_steps * 8
        self.d2h_bytes_per_iteration = 16

    # ------------------------------------------------------------------ device-resident path
    def set_batch(self, ann_g: torch.Tensor, ann_d: torch.Tensor, labels: torch.Tensor) -> None:
        self.eng.set_batch(ann_g, ann_d, labels)

    def _allreduce(self, bucket) -> None:
        if self.world > 1:
            self.dist.all_reduce(bucket.grad, op=self.dist.ReduceOp.SUM, group=self.pg)

    def disc_step(self) -> None:
        """train.py:365 sess.run(disc_train_op)."""
        e = self.eng
        e.sample_noise()
        e.sample_gp_alpha()
        e.disc_step()
        self._allreduce(e.d)
        e.d.adam_step()

    def gen_step(self) -> None:
        """train.py:368 sess.run(gen_train_op)."""
        e = self.eng
        e.sample_noise()
        e.gen_step()
        self._allreduce(e.g)
        e.g.adam_step()
        e._refresh = True          # generator weights changed: its hoisted projection is stale

    def iteration(self) -> None:
        """train.py:362-368 loop body on the batch given to set_batch()."""
        for _ in range(self.critic_iters):
            self.disc_step()
        self.gen_step()
        self.iterations += 1

    def losses(self) -> Dict[str, float]:
        """Host copy of the last step's scalars (forces a stream sync, 16 bytes D2H).  With world > 1
        these are this rank's shard of the global means (sum over ranks = the global value)."""
        self._loss_host.copy_(self.eng.scalars, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        w, gp, gc = (float(x) for x in self._loss_host[1:4])
        return {"w_disc": w, "gp": gp, "disc_cost": w + self.lam * gp, "gen_cost": gc}

    # ------------------------------------------------------------------ host-buffer path (end to end)
    def _alloc_slot(self):
        dev = self.device
        return (torch.empty(self.B, self.R, 512, dtype=torch.bfloat16, device=dev),
                torch.empty(self.B, self.R, 512, dtype=torch.bfloat16, device=dev),
                torch.empty(self.B, self.T, dtype=torch.int64, device=dev))

    def upload(self, ann_g_host: torch.Tensor, ann_d_host: torch.Tensor, labels_host: torch.Tensor) -> int:
        """Asynchronous H2D of one batch (pinned bf16 annotations [B,R,512] / [B,14,14,512], int64 labels
        [B,T]) into the staging slot that is NOT in use; returns the slot id."""
        slot = 1 - self._cur
        if self._slots[slot] is None:
            self._slots[slot] = self._alloc_slot()
        cs = self._copy_stream
        cs.wait_stream(torch.cuda.current_stream())   # the slot's previous consumer has been enqueued before
        with torch.cuda.stream(cs):
            for dst, src in zip(self._slots[slot], (ann_g_host, ann_d_host, labels_host)):
                dst.view(-1).copy_(src.view(-1), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        self._slot_ready[slot] = ev
        return slot

    def use_slot(self, slot: int) -> None:
        torch.cuda.current_stream().wait_event(self._slot_ready[slot])
        self._cur = slot
        self.eng.set_batch(*self._slots[slot])

    def fit(self, host_batches: Iterable[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]):
        """Trains one iteration per host batch; the upload of batch i+1 overlaps the compute of batch i.
        Yields the losses of every iteration (a 16-byte D2H read per iteration)."""
        it = iter(host_batches)
        try:
            nxt = self.upload(*next(it))
        except StopIteration:
            return
        while nxt is not None:
            self.use_slot(nxt)
            try:
                nxt = self.upload(*next(it))
            except StopIteration:
                nxt = None
            self.iteration()
            yield self.losses()
