"""WGAN-GP training schedule of the reference (train.py:362-368) over the CUDA step engine.

One iteration = ``critic_iters`` discriminator steps followed by one generator step, all on the
same data batch (train.py:185-187 repeats each batch CRITIC_ITERS+1 times), fresh noise
(gen:81) and fresh interpolation coefficients (tfgan gradient penalty) per step, two
tf.train.AdamOptimizer(1e-4, beta1=0.5, beta2=0.9) updates (train.py:258-266).  The whole
iteration is ONE C-ABI call (sgg_train_iteration) that is captured once per input-buffer set
into a CUDA graph and replayed.

Data parallelism (the reference has none): one process per GPU, the batch is sharded
over ranks, the kernels normalise every loss by the GLOBAL batch, so the only exchange is one
sum of the flat gradient bucket per optimiser step (NCCL over NVLink/NVSwitch, enqueued by the
library on the compute stream, inside the graph).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional, Tuple

import torch

from . import dp
from ._lib import check, lib
from .engine import Engine

COMM_ID_BYTES = 128


class HotPathTrainer:
    def __init__(self, batch_size: int, n_steps: int = 3, vocab_size: int = 2000, critic_iters: int = 5,
                 lam: float = 10.0, regions: int = 196, embed_dim: int = 300, seed: int = 0,
                 embedding: Optional[torch.Tensor] = None, process_group=None, device=None,
                 use_graph: bool = True, lr: float = 1e-4, beta1: float = 0.5, beta2: float = 0.9):
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.pg = process_group
        self.world = self.dist.get_world_size(process_group) if self.dist else 1
        self.rank = self.dist.get_rank(process_group) if self.dist else 0
        self.B, self.T, self.V, self.R = batch_size, n_steps, vocab_size, regions
        self.critic_iters, self.lam = int(critic_iters), float(lam)
        self.lr, self.beta1, self.beta2 = lr, beta1, beta2
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        # identical initial weights on every rank (same seed); decorrelated noise / alpha streams per rank
        self.eng = Engine(batch_size, n_steps, vocab_size, regions, embed_dim, lam=lam, world=self.world,
                          seed=dp.rank_seed(seed, self.rank), device=self.device,
                          critic_iters=self.critic_iters)
        self.eng.g.init_reference(seed * 2 + 1)
        self.eng.d.init_reference(seed * 2 + 2, embedding=embedding)
        self.comm = self._init_comm() if self.world > 1 else None
        self.iterations = 0
        self.use_graph = use_graph
        self._graphs: Dict[Tuple[int, int, int], Tuple[torch.cuda.CUDAGraph, int]] = {}
        self.kernel_launches = 0          # kernels of this library executed so far (graph replays included)
        # double-buffered device staging for host batches
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._slots = [None, None]
        self._slot_ready = [None, None]
        self._cur = 0
        self._loss_host = torch.zeros(self.critic_iters + 1, 4, dtype=torch.float32).pin_memory()
        self.h2d_bytes_per_batch = 2 * batch_size * regions * 512 * 2 + batch_size * n_steps * 8
        self.d2h_bytes_per_iteration = self._loss_host.numel() * 4

    # ------------------------------------------------------------------ communicator
    def _init_comm(self):
        def make_id() -> bytes:
            ident = (C.c_ubyte * COMM_ID_BYTES)()
            check(lib().sgg_comm_unique_id(ident), "sgg_comm_unique_id")
            return bytes(ident)
        ident = dp.broadcast_comm_id(self.dist, self.pg, self.rank, make_id)
        handle = C.c_void_p(0)
        check(lib().sgg_comm_init(ident, C.c_int32(self.rank), C.c_int32(self.world), C.byref(handle)), "sgg_comm_init")
        return handle

    def gather_sharded(self) -> None:
        """With world > 1 the annotation rows of attention_perceptron/kernel are row-sharded over the ranks (see
        sgg_wa_shard_t); call this before reading the full parameter / optimiser tensors."""
        if self.world > 1:
            torch.cuda.synchronize()
            self.eng.gather_sharded(self.dist, self.pg, self.rank)

    def close(self) -> None:
        if self.comm is not None:
            torch.cuda.synchronize()
            self._graphs.clear()
            lib().sgg_comm_destroy(self.comm)
            self.comm = None

    # ------------------------------------------------------------------ device-resident path
    def set_batch(self, ann_g: torch.Tensor, ann_d: torch.Tensor, labels: torch.Tensor) -> None:
        self.eng.set_batch(ann_g, ann_d, labels)

    def _launch_count(self) -> int:
        fn = lib().sgg_launch_count
        fn.restype = C.c_int64
        return int(fn())

    def _run_iteration_eager(self) -> None:
        self.eng.train_iteration(self.critic_iters, comm=self.comm, lr=self.lr, beta1=self.beta1, beta2=self.beta2)

    def iteration(self) -> None:
        """train.py:362-368 loop body on the batch given to set_batch()."""
        e = self.eng
        if not self.use_graph:
            n0 = self._launch_count()
            self._run_iteration_eager()
            self.kernel_launches += self._launch_count() - n0
        else:
            key = (e.ann_g.data_ptr(), e.ann_d.data_ptr(), e.labels.data_ptr())
            entry = self._graphs.get(key)
            if entry is None:
                if len(self._graphs) >= 8:
                    self._graphs.clear()
                # warm-up outside capture (lazy kernel attribute setup), on a side stream as capture requires
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    n0 = self._launch_count()
                    self._run_iteration_eager()
                    self.kernel_launches += self._launch_count() - n0
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                n0 = self._launch_count()
                with torch.cuda.graph(graph, stream=side):
                    self._run_iteration_eager()
                entry = (graph, self._launch_count() - n0)
                self._graphs[key] = entry
                # the eager warm-up above already performed this call's iteration
            else:
                entry[0].replay()
                self.kernel_launches += entry[1]
        e.d.step += self.critic_iters
        e.g.step += 1
        self.iterations += 1

    def losses(self) -> Dict[str, float]:
        """Host copy of the last iteration's per-step scalars (one small D2H + stream sync).  With world > 1
        these are this rank's shard of the global means (sum over ranks = the global value)."""
        self._loss_host.copy_(self.eng.scalars_all, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._losses_from(self._loss_host)

    # ------------------------------------------------------------------ host-buffer path (end to end)
    def _alloc_slot(self):
        dev = self.device
        return (torch.empty(self.B, self.R, 512, dtype=torch.bfloat16, device=dev),
                torch.empty(self.B, self.R, 512, dtype=torch.bfloat16, device=dev),
                torch.empty(self.B, self.T, dtype=torch.int64, device=dev))

    def upload(self, ann_g_host: torch.Tensor, ann_d_host: torch.Tensor, labels_host: torch.Tensor) -> int:
        """Asynchronous H2D of one batch (pinned bf16 annotations [B,R,512] / [B,14,14,512], int64 labels
        [B,T]) into the staging slot that is NOT in use; returns the slot id."""
        if labels_host.numel() and (int(labels_host.min()) < 0 or int(labels_host.max()) >= self.V):
            raise ValueError(f"label ids must lie in [0, {self.V}); got [{int(labels_host.min())}, {int(labels_host.max())}]")
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
        self.eng.set_batch(*self._slots[slot], validate=False)   # the host copy was range-checked in upload()

    def _losses_from(self, host: torch.Tensor) -> Dict[str, float]:
        nc = self.critic_iters
        out = {"gen_cost": float(host[nc, 3])}
        if nc > 0:
            w, gp = float(host[nc - 1, 1]), float(host[nc - 1, 2])
            out.update({"w_disc": w, "gp": gp, "disc_cost": w + self.lam * gp})
        return out

    def fit(self, host_batches: Iterable[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]):
        """Trains one iteration per host batch and yields the losses of every iteration, in order.  Software pipeline:
        the upload of batch i+1 (copy stream) overlaps the compute of batch i, and the 96-byte loss read of iteration
        i is enqueued behind it on the compute stream into one of two pinned buffers and only waited for after
        iteration i+1 has been launched, so the host never idles the GPU between iterations."""
        it = iter(host_batches)
        try:
            nxt = self.upload(*next(it))
        except StopIteration:
            return
        if not hasattr(self, "_loss_ring"):
            self._loss_ring = [torch.zeros_like(self._loss_host).pin_memory() for _ in range(2)]
        pending = None      # (event, pinned buffer) of the previous iteration
        k = 0
        while nxt is not None:
            self.use_slot(nxt)
            try:
                nxt = self.upload(*next(it))
            except StopIteration:
                nxt = None
            self.iteration()
            buf = self._loss_ring[k & 1]
            buf.copy_(self.eng.scalars_all, non_blocking=True)      # stream-ordered before the next iteration overwrites it
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            if pending is not None:
                pending[0].synchronize()
                yield self._losses_from(pending[1])
            pending = (ev, buf)
            k += 1
        if pending is not None:
            pending[0].synchronize()
            yield self._losses_from(pending[1])
