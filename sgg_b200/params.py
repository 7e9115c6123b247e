"""Flat parameter buckets with the reference's TF variable names.

One fp32 master bucket per network (+ gradient and Adam moment buckets of the same layout)
and one bf16 shadow bucket holding the tensor-core operands.  The name -> view table
reproduces `Generator/Generator/*`, `Discriminator/Discriminator/*`, `Discriminator/W`
(train.py:262-263 splits variables by these prefixes).
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from typing import Dict

import torch

from ._lib import Dims, ParamEntry, check, lib, stream_ptr

GEN, DISC = 0, 1


def make_dims(B, T, V, R=196, C_=512, H=512, E=300, S=1) -> Dims:
    return Dims(B=B, T=T, V=V, R=R, C=C_, H=H, E=E, S=S)


class ParamBucket:
    def __init__(self, net: int, dims: Dims, device="cuda"):
        self.net, self.dims = net, dims
        n, nf, ns = C.c_int(0), C.c_int64(0), C.c_int64(0)
        check(lib().sgg_param_table(net, C.byref(dims), None, 0, C.byref(n), C.byref(nf), C.byref(ns)), "sgg_param_table")
        ents = (ParamEntry * n.value)()
        check(lib().sgg_param_table(net, C.byref(dims), ents, n.value, C.byref(n), C.byref(nf), C.byref(ns)), "sgg_param_table")
        self.entries = [(e.name.decode(), e.offset, e.rows, e.cols, e.shadow_offset, e.shadow_pitch) for e in ents]
        self.shadow_rows = {e.name.decode(): e.shadow_rows for e in ents}
        self.n_floats, self.n_shadow = nf.value, ns.value
        self.theta = torch.zeros(self.n_floats, dtype=torch.float32, device=device)
        self.grad = torch.zeros_like(self.theta)
        self.m = torch.zeros_like(self.theta)
        self.v = torch.zeros_like(self.theta)
        self.shadow = torch.zeros(self.n_shadow, dtype=torch.bfloat16, device=device)
        self.step = 0
        # bumped whenever theta may have changed through this object (Adam, load, shadow refresh after an in-place
        # edit of a view): consumers that cache products of the weights (Engine: the hoisted projection P) compare it
        self.version = 0

    # ---- named views in the reference's (TF) shapes
    def _views(self, flat: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
        out = OrderedDict()
        for name, off, rows, cols, _, _ in self.entries:
            v = flat[off:off + rows * cols]
            out[name] = v.view(rows, cols) if rows > 1 or name.endswith("kernel") or name.endswith("/W") else v.view(cols)
        return out

    def views(self):
        return self._views(self.theta)

    def grad_views(self):
        return self._views(self.grad)

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        views = self.views()
        missing = set(views) - set(sd)
        if missing:
            raise KeyError(f"missing variables: {sorted(missing)}")
        for k, v in views.items():
            v.copy_(sd[k].to(device=v.device, dtype=torch.float32).reshape(v.shape))
        self.refresh_shadow()

    def state_dict(self):
        return OrderedDict((k, v.detach().clone()) for k, v in self.views().items())

    def refresh_shadow(self, stream=None) -> None:
        self.version += 1
        check(lib().sgg_refresh_shadow(self.net, C.byref(self.dims), C.c_void_p(self.theta.data_ptr()),
                                       C.c_void_p(self.shadow.data_ptr()), stream_ptr(stream)), "sgg_refresh_shadow")

    def init_reference(self, seed: int, embedding: torch.Tensor | None = None) -> None:
        """The reference's initialisers: glorot_uniform kernels (tf.layers.dense / LSTM cell
        defaults), zero biases, LN gamma=1 beta=0; Discriminator/W from the caller's embedding
        matrix (train:69-72) or U(-0.1,0.1) (dataset_creation/map_files_to_triples.py:24)."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        sd = {}
        for name, _, rows, cols, _, _ in self.entries:
            if name.endswith("/W"):
                sd[name] = embedding if embedding is not None else torch.rand(rows, cols, generator=g) * 0.2 - 0.1
            elif name.endswith("kernel"):
                lim = math.sqrt(6.0 / (rows + cols))
                sd[name] = (torch.rand(rows, cols, generator=g) * 2 - 1) * lim
            elif name.endswith("gamma"):
                sd[name] = torch.ones(cols)
            else:
                sd[name] = torch.zeros(cols)
        self.load_state_dict(sd)

    def adam_step(self, lr=1e-4, beta1=0.5, beta2=0.9, eps=1e-8, grad_scale=1.0, stream=None) -> None:
        """tf.train.AdamOptimizer(1e-4, beta1=0.5, beta2=0.9) of train.py:258-259."""
        self.step += 1
        self.version += 1
        check(lib().sgg_adam_step(self.net, C.byref(self.dims), C.c_void_p(self.theta.data_ptr()),
                                  C.c_void_p(self.grad.data_ptr()), C.c_void_p(self.m.data_ptr()),
                                  C.c_void_p(self.v.data_ptr()), C.c_void_p(self.shadow.data_ptr()),
                                  C.c_int64(self.step), C.c_float(lr), C.c_float(beta1), C.c_float(beta2),
                                  C.c_float(eps), C.c_float(grad_scale), stream_ptr(stream)), "sgg_adam_step")
