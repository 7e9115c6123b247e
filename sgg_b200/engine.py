"""Step engine: owns the workspace and drives the C-ABI step functions on the current stream."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from ._lib import FLAG_REFRESH_GEN_PROJ, Dims, IterArgs, StepArgs, check, lib, stream_ptr
from .params import DISC, GEN, ParamBucket, make_dims


def _p(t: Optional[torch.Tensor]):
    return t.data_ptr() if t is not None else None


class Engine:
    """One data-parallel rank of the WGAN-GP hot path (train.py:231-266, 362-368)."""

    def __init__(self, B: int, T: int = 3, V: int = 2000, R: int = 196, E: int = 300, lam: float = 10.0,
                 world: int = 1, seed: int = 0, device="cuda", critic_iters: int = 1):
        if not torch.cuda.is_available():
            raise RuntimeError("sgg_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.critic_iters = int(critic_iters)
        self.dims: Dims = make_dims(B, T, V, R, 512, 512, E, S=max(1, self.critic_iters))
        self.B, self.T, self.V, self.R, self.E = B, T, V, R, E
        self.lam, self.world, self.device = float(lam), int(world), device
        nbytes = lib().sgg_workspace_bytes
        nbytes.restype = C.c_int64
        self.ws_bytes = nbytes(C.byref(self.dims))
        if self.ws_bytes <= 0:
            raise RuntimeError("sgg_workspace_bytes: " + lib().sgg_last_error().decode())
        self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=device)
        self.g = ParamBucket(GEN, self.dims, device)
        self.d = ParamBucket(DISC, self.dims, device)
        self.scalars = torch.zeros(4, dtype=torch.float32, device=device)
        self.noise = torch.zeros(B, 512, dtype=torch.float32, device=device)
        self.gp_alpha = torch.zeros(B, dtype=torch.float32, device=device)
        self.logits = torch.zeros(B, T, V, dtype=torch.float32, device=device)
        self.ann_g = self.ann_d = self.labels = None
        self._refresh = True
        self._g_version = -1            # generator bucket version the cached projection P / c0 was computed from
        self._checked_labels = {}       # (data_ptr, tensor version) of label tensors already range-checked
        self.seed, self._rng_off = seed, 0
        # sgg_train_iteration state: device-side iteration counter, per-step randomness and losses
        nc = max(1, self.critic_iters)
        self.counters = torch.zeros(1, dtype=torch.int64, device=device)
        self.noise_all = torch.zeros(nc + 1, B, 512, dtype=torch.float32, device=device)
        self.gp_alpha_all = torch.zeros(nc, B, dtype=torch.float32, device=device)
        self.scalars_all = torch.zeros(nc + 1, 4, dtype=torch.float32, device=device)
        # row-sharded attention projection (include/sgg_b200.h: sgg_wa_shard_t): on by default when world > 1 and the
        # contraction splits evenly; SGG_WA_SHARD=0 keeps W_a replicated (all-reduce of its gradient) for comparison
        self.shard = (self.world > 1 and os.environ.get("SGG_WA_SHARD", "1") != "0" and (R * 512 // 64) % self.world == 0)
        self.slab_g = self.slab_d = self.shard_scratch = None
        if self.shard:
            L = lib()
            L.sgg_wa_shard_scratch_bytes.restype = C.c_int64
            L.sgg_wa_shard_slab_elems.restype = C.c_int64
            n_slab = L.sgg_wa_shard_slab_elems(C.byref(self.dims), C.c_int32(self.world))
            n_scr = L.sgg_wa_shard_scratch_bytes(C.byref(self.dims), C.c_int32(self.world))
            if n_slab <= 0 or n_scr <= 0:
                raise RuntimeError("sgg_wa_shard_*: " + L.sgg_last_error().decode())
            self.slab_g = torch.zeros(n_slab, dtype=torch.bfloat16, device=device)
            self.slab_d = torch.zeros(n_slab, dtype=torch.bfloat16, device=device)
            self.shard_scratch = torch.zeros(n_scr, dtype=torch.uint8, device=device)

    # ------------------------------------------------------------------ inputs
    def set_batch(self, ann_g: torch.Tensor, ann_d: torch.Tensor, labels: Optional[torch.Tensor],
                  validate: bool = True) -> None:
        """Annotations [B,R,512] (or [B,14,14,512]) bf16 on device; labels [B,T] int64.  The same batch
        serves the n_critic + 1 steps of an iteration (train.py:185-187).  Labels index rows of Discriminator/W
        (one-hot of train.py:173): ids outside [0, V) raise here (one device min/max per distinct label tensor; the
        kernels additionally clamp, so a skipped check cannot corrupt memory).  validate=False skips the check for
        callers that validated the host copy (HotPathTrainer.upload)."""
        for a in (ann_g, ann_d):
            assert a.dtype == torch.bfloat16 and a.is_cuda and a.is_contiguous()
            assert a.numel() == self.B * self.R * 512, "annotation shape mismatch"
        self.ann_g, self.ann_d = ann_g, ann_d
        if labels is not None:
            assert labels.dtype == torch.int64 and labels.shape == (self.B, self.T) and labels.is_contiguous()
            key = (labels.data_ptr(), labels._version)
            if validate and self._checked_labels.get(labels.data_ptr()) != key:
                lo, hi = int(labels.min()), int(labels.max())
                if lo < 0 or hi >= self.V:
                    raise ValueError(f"label ids must lie in [0, {self.V}); got [{lo}, {hi}]")
                if len(self._checked_labels) > 64:
                    self._checked_labels.clear()
                self._checked_labels[labels.data_ptr()] = key
        self.labels = labels
        self._refresh = True

    def sample_noise(self, stream=None) -> None:
        check(lib().sgg_rng_fill_normal(C.c_void_p(self.noise.data_ptr()), C.c_int64(self.noise.numel()),
                                        C.c_uint64(self.seed), C.c_uint64(self._rng_off), stream_ptr(stream)), "rng")
        self._rng_off += (self.noise.numel() + 3) // 4

    def sample_gp_alpha(self, stream=None) -> None:
        check(lib().sgg_rng_fill_uniform(C.c_void_p(self.gp_alpha.data_ptr()), C.c_int64(self.B),
                                         C.c_uint64(self.seed ^ 0x9E3779B97F4A7C15), C.c_uint64(self._rng_off),
                                         stream_ptr(stream)), "rng")
        self._rng_off += (self.B + 3) // 4

    def _ann_grad_buffer(self, which: str) -> torch.Tensor:
        name = "_ann_grad_" + which
        if getattr(self, name, None) is None:
            setattr(self, name, torch.empty(self.B, self.R, 512, dtype=torch.float32, device=self.device))
        return getattr(self, name)

    def _args(self, want_logits=False) -> StepArgs:
        a = StepArgs()
        a.dims, a.world, a.lam = self.dims, self.world, self.lam
        a.g_theta, a.g_shadow, a.g_grad = self.g.theta.data_ptr(), self.g.shadow.data_ptr(), self.g.grad.data_ptr()
        a.d_theta, a.d_shadow, a.d_grad = self.d.theta.data_ptr(), self.d.shadow.data_ptr(), self.d.grad.data_ptr()
        a.ann_g, a.ann_d, a.labels = _p(self.ann_g), _p(self.ann_d), _p(self.labels)
        a.noise, a.gp_alpha = self.noise.data_ptr(), self.gp_alpha.data_ptr()
        a.workspace, a.workspace_bytes = self.ws.data_ptr(), self.ws_bytes
        a.scalars = self.scalars.data_ptr()
        a.logits_out = self.logits.data_ptr() if want_logits else None
        # the hoisted projection P = flat(a_g) W_a and c0 are cached in the workspace: recompute them for a new batch
        # and whenever the generator's weights may have changed since (Adam step, load_state_dict, refresh_shadow)
        a.flags = FLAG_REFRESH_GEN_PROJ if (self._refresh or self._g_version != self.g.version) else 0
        return a

    def _proj_cached(self) -> None:
        self._refresh = False
        self._g_version = self.g.version

    # ------------------------------------------------------------------ steps
    def gen_forward(self, stream=None) -> torch.Tensor:
        a = self._args(want_logits=True)
        check(lib().sgg_gen_forward(C.byref(a), stream_ptr(stream)), "sgg_gen_forward")
        self._proj_cached()
        return self.logits

    def disc_forward(self, triples: torch.Tensor, stream=None) -> torch.Tensor:
        assert triples.dtype == torch.float32 and triples.shape == (self.B, self.T, self.V) and triples.is_contiguous()
        out = torch.empty(self.B, self.T, dtype=torch.float32, device=self.device)
        a = self._args()
        check(lib().sgg_disc_forward(C.byref(a), C.c_void_p(triples.data_ptr()), C.c_void_p(out.data_ptr()),
                                     stream_ptr(stream)), "sgg_disc_forward")
        return out

    def disc_step(self, stream=None, ann_grad: bool = False) -> Optional[torch.Tensor]:
        """Gradients of disc_cost into self.d.grad; scalars[1] = w_disc, scalars[2] = gp.
        ann_grad=True also returns d disc_cost / d ann_d [B,R,512] fp32 (what the discriminator's conv front-end
        disc:29-68 back-propagates; the buffer is reused by the next call)."""
        a = self._args()
        out = self._ann_grad_buffer("d") if ann_grad else None
        a.ann_d_grad = _p(out)
        check(lib().sgg_disc_step(C.byref(a), stream_ptr(stream)), "sgg_disc_step")
        self._proj_cached()
        return out

    def gen_step(self, stream=None, ann_grad: bool = False) -> Optional[torch.Tensor]:
        """Gradients of gen_cost into self.g.grad; scalars[3] = gen_cost.
        ann_grad=True also returns d gen_cost / d ann_g [B,R,512] fp32 (gen:29-68's upstream gradient)."""
        a = self._args()
        out = self._ann_grad_buffer("g") if ann_grad else None
        a.ann_g_grad = _p(out)
        check(lib().sgg_gen_step(C.byref(a), stream_ptr(stream)), "sgg_gen_step")
        self._proj_cached()
        return out

    def train_iteration(self, critic_iters: Optional[int] = None, comm=None, lr=1e-4, beta1=0.5, beta2=0.9,
                        eps=1e-8, stream=None) -> None:
        """One reference loop body (train.py:362-368): critic_iters D steps + 1 G step with their Adam
        updates, randomness drawn on the device.  Per-step losses land in self.scalars_all.  Safe to
        capture into a CUDA graph (no host-dependent state)."""
        nc = self.critic_iters if critic_iters is None else int(critic_iters)
        it = IterArgs()
        it.step = self._args()
        it.critic_iters = nc
        it.g_m, it.g_v, it.d_m, it.d_v = (self.g.m.data_ptr(), self.g.v.data_ptr(), self.d.m.data_ptr(),
                                          self.d.v.data_ptr())
        it.lr, it.beta1, it.beta2, it.eps = lr, beta1, beta2, eps
        it.seed = self.seed
        it.counters = self.counters.data_ptr()
        it.noise_all, it.gp_alpha_all = self.noise_all.data_ptr(), self.gp_alpha_all.data_ptr()
        it.scalars_all = self.scalars_all.data_ptr()
        it.comm = comm
        if self.shard and comm is not None:
            it.shard.enabled = 1
            it.shard.slab_g, it.shard.slab_d = self.slab_g.data_ptr(), self.slab_d.data_ptr()
            it.shard.scratch, it.shard.scratch_bytes = self.shard_scratch.data_ptr(), self.shard_scratch.numel()
        check(lib().sgg_train_iteration(C.byref(it), stream_ptr(stream)), "sgg_train_iteration")
        self._refresh = True   # the generator was updated inside the library
        self.g.version += 1
        self.d.version += 1

    def wa_rows(self, rank: int):
        """Rows of attention_perceptron/kernel rank `rank` maintains under the row-sharded projection."""
        ks = (self.R * 512 // 64) // self.world * 64
        return rank * ks, (rank + 1) * ks

    def gather_sharded(self, dist, group=None, rank: int = 0) -> None:
        """Row-sharded projection: every rank keeps only its rows of W_a (theta, Adam moments, shadow) current.  This
        all-gathers the rows so that the full tensors can be read (checkpointing, evaluation, tests)."""
        if not self.shard:
            return
        for bucket in (self.g, self.d):
            name, off, rows, cols, soff, pitch = next(e for e in bucket.entries if e[0].endswith("attention_perceptron/kernel"))
            r0, r1 = self.wa_rows(rank)
            n = (r1 - r0) * cols
            for flat in (bucket.theta, bucket.m, bucket.v):
                block = flat[off:off + self.world * n]
                dist.all_gather_into_tensor(block, block[rank * n:(rank + 1) * n].clone(), group=group)
            bucket.refresh_shadow()

    def ws_view(self, name: str, shape, dtype) -> torch.Tensor:
        """Test accessor: a view of a named workspace buffer."""
        off, eb = C.c_int64(0), C.c_int64(0)
        check(lib().sgg_ws_lookup(C.byref(self.dims), name.encode(), C.byref(off), C.byref(eb)), "sgg_ws_lookup")
        n = 1
        for s in shape:
            n *= s
        return self.ws[off.value:off.value + n * eb.value].view(dtype).view(*shape)
