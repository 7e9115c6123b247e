"""Shared plumbing of the Generator / Discriminator front classes."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from .. import ops
from .._lib import check, lib, stream_ptr
from ..engine import Engine

REGION_CHANNELS = 512


def as_annotations(images: torch.Tensor, device) -> torch.Tensor:
    """The hot path starts at ``self.downsampled`` (gen:68 / disc:68): ``images`` must already be the
    [B, 14, 14, 512] annotation grid (any [B, h, w, 512] or [B, R, 512] is accepted).  The reference's conv
    front-end that produces it from 221x221x3 images is outside this path; a caller-side version on
    library convolutions is ``sgg_b200.frontend.ConvFrontEnd``."""
    if images.dim() not in (3, 4) or images.shape[-1] != REGION_CHANNELS:
        raise ValueError(
            f"expected annotations [B,14,14,512] (the reference's self.downsampled), got {tuple(images.shape)}; "
            "the convolutional front-end is not part of the B200 hot path (sgg_b200.frontend.ConvFrontEnd produces "
            "the grid from pixels with library convolutions)")
    return images.to(device=device, dtype=torch.bfloat16).contiguous()


class AttentionNet:
    """State shared by both networks: the engine (parameters + workspace) and the reference's tensor attributes."""

    def __init__(self, vocab_size: int, n_steps: int = 3):
        self.vocab_size = int(vocab_size)
        self.n_steps = int(n_steps)
        self._engine: Optional[Engine] = None
        self._owns_engine = False
        self.downsampled = self.flattened_context = self.partially_flattened_context = self.alpha = None

    # the trainer shares one engine between G and D; a stand-alone network creates its own
    def _attach(self, engine: Engine) -> None:
        self._engine = engine

    def _ensure_engine(self, B: int, R: int, embedding=None, seed: int = 0) -> Engine:
        e = self._engine
        if e is None or (self._owns_engine and (e.B != B or e.R != R)):
            old = e
            e = Engine(B, self.n_steps, self.vocab_size, R, embedding.shape[1] if embedding is not None else 300)
            if old is None:
                e.g.init_reference(2 * seed + 1)
                e.d.init_reference(2 * seed + 2, embedding=embedding)
            else:  # batch size changed: keep the trained variables
                e.g.load_state_dict(old.g.state_dict())
                e.d.load_state_dict(old.d.state_dict())
            self._engine, self._owns_engine = e, True
        elif e.B != B or e.R != R:
            raise ValueError(f"engine was built for batch {e.B} x {e.R} regions, got {B} x {R}")
        return e

    def _set_context(self, ann: torch.Tensor) -> None:
        B = ann.shape[0]
        self.downsampled = ann if ann.dim() == 4 else ann.view(B, -1, 1, REGION_CHANNELS)
        self.flattened_context = ann.view(B, -1)                              # gen:74
        self.partially_flattened_context = ann.view(B, -1, REGION_CHANNELS)   # gen:75

    def _attention(self, bucket, cell_state: Tuple[torch.Tensor, torch.Tensor]) -> torch.Tensor:
        """gen:13-18 on the current context: e = [flat(a), c] W + b (uses cell_state[0] = c), softmax, context."""
        if self.partially_flattened_context is None:
            raise RuntimeError("attentionMechanism needs a context: call the build method first")
        a = self.partially_flattened_context
        B, R, _ = a.shape
        c = cell_state[0].to(device=a.device, dtype=torch.float32)
        v = bucket.views()
        prefix = "Generator/Generator" if bucket.net == 0 else "Discriminator/Discriminator"
        W = v[f"{prefix}/attention_perceptron/kernel"]
        bias = v[f"{prefix}/attention_perceptron/bias"]
        # flat(a) W_a on the tensor cores (W_a as a hi/lo pair from the shadow bucket), c W_h in fp32 (tiny)
        name, off, rows, cols, soff, pitch = next(e for e in bucket.entries if e[0].endswith("attention_perceptron/kernel"))
        srows = bucket.shadow_rows[name]
        Wa = bucket.shadow[soff:soff + 2 * srows * pitch].view(2 * srows, pitch)[:, :R]
        e = torch.empty(B, 256, dtype=torch.float32, device=a.device)
        ops.gemm(a.view(B, R * REGION_CHANNELS), Wa, B, R, b_mn=True,
                 segs=[(0, 0, 0, 0, R * REGION_CHANNELS), (0, 0, srows, 0, R * REGION_CHANNELS)], out=e[:, :R], splits=0)
        e[:, :R] += c @ W[R * REGION_CHANNELS:] + bias
        alpha = torch.empty_like(e)
        z = torch.zeros(B, 2 * REGION_CHANNELS, dtype=torch.bfloat16, device=a.device)
        check(lib().sgg_attn_forward(C.c_void_p(a.data_ptr()), C.c_int32(B), C.c_int32(R), C.c_int32(1),
                                     C.c_void_p(e.data_ptr()), C.c_void_p(alpha.data_ptr()), C.c_int64(256),
                                     C.c_void_p(z.data_ptr()), C.c_int64(2 * REGION_CHANNELS), C.c_int64(REGION_CHANNELS),
                                     stream_ptr()), "sgg_attn_forward")
        self.alpha = alpha[:, :R]
        return z[:, :REGION_CHANNELS].float() + z[:, REGION_CHANNELS:].float()
