"""Generator-only inference: forward pass + test-time decoding (train.py:269-270: ``tf.argmax(generator_output,
axis=2)``), or Gumbel-max sampling as an extension.  Thin host wrapper over ``sgg_gen_sample``."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from ._lib import SAMPLE_GREEDY, SAMPLE_GUMBEL, SampleArgs, check, lib, stream_ptr
from .params import GEN, ParamBucket, make_dims

MODES = {"greedy": SAMPLE_GREEDY, "argmax": SAMPLE_GREEDY, "gumbel": SAMPLE_GUMBEL}


class GeneratorSampler:
    """Decodes ``T`` tokens per image for batches of ``B`` annotation grids with the generator variables of
    ``bucket`` (a ``ParamBucket`` of net GEN; its vocabulary fixes V)."""

    def __init__(self, bucket: ParamBucket, B: int, T: int = 3, R: int = 196, chunk: int = 0, seed: int = 0, device="cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("sgg_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        assert bucket.net == GEN
        self.bucket, self.B, self.T, self.R, self.V = bucket, int(B), int(T), int(R), int(bucket.dims.V)
        self.chunk, self.seed, self.device = int(chunk), int(seed), device
        self.dims = make_dims(self.B, self.T, self.V, self.R, 512, 512, int(bucket.dims.E), S=1)
        f = lib().sgg_sample_workspace_bytes
        f.restype = C.c_int64
        self.ws_bytes = f(C.byref(self.dims), C.c_int32(self.chunk))
        if self.ws_bytes <= 0:
            raise RuntimeError("sgg_sample_workspace_bytes: " + lib().sgg_last_error().decode())
        self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=device)
        self.tokens = torch.zeros(self.B, self.T, dtype=torch.int32, device=device)
        self._calls = 0

    def sample(self, ann: torch.Tensor, mode: str = "greedy", noise: Optional[torch.Tensor] = None,
               want_logits: bool = False, stream=None):
        """ann: [B,R,512] (or [B,14,14,512]) bf16 on the device.  Returns tokens [B,T] int32 (and the raw logits
        [B,T,V] fp32 when asked).  ``noise`` [B,512] may be injected; otherwise N(0,1) is drawn on the device."""
        assert ann.dtype == torch.bfloat16 and ann.is_cuda and ann.is_contiguous()
        assert ann.numel() == self.B * self.R * 512, "annotation shape mismatch"
        logits = torch.empty(self.B, self.T, self.V, dtype=torch.float32, device=self.device) if want_logits else None
        a = SampleArgs()
        a.dims = self.dims
        a.g_theta, a.g_shadow = self.bucket.theta.data_ptr(), self.bucket.shadow.data_ptr()
        a.ann_g = ann.data_ptr()
        if noise is not None:
            assert noise.dtype == torch.float32 and noise.shape == (self.B, 512) and noise.is_cuda and noise.is_contiguous()
            a.noise = noise.data_ptr()
        a.mode, a.chunk = MODES[mode], self.chunk
        a.seed = self.seed
        a.offset = self._calls * (self.B * (512 // 4 + self.T))   # a fresh Philox range per call
        a.workspace, a.workspace_bytes = self.ws.data_ptr(), self.ws_bytes
        a.tokens_out = self.tokens.data_ptr()
        a.logits_out = logits.data_ptr() if logits is not None else None
        check(lib().sgg_gen_sample(C.byref(a), stream_ptr(stream)), "sgg_gen_sample")
        self._calls += 1
        return (self.tokens, logits) if want_logits else self.tokens
