"""``Generator`` with the reference's interface (architectures/generator_with_attention.py:9-91).

    g = Generator(vocab_size)
    logits = g.build_generator(annotations)      # [B, 3, V] raw logits (gen:88-91: no softmax anywhere)
    g.alpha                                       # [B, 196] attention weights of the last timestep (gen:16)

Differences from the TensorFlow original, all outside the hot path: ``images`` is the annotation grid
``self.downsampled`` [B,14,14,512] rather than 221x221x3 pixels, the call is eager (it runs the kernels and
returns a CUDA tensor) and the variables live in a flat bucket whose named views carry the TF variable names.
"""
from __future__ import annotations

from typing import Optional

import torch

from ._common import AttentionNet, as_annotations


class Generator(AttentionNet):
    def __init__(self, vocab_size, n_steps: int = 3):
        super().__init__(vocab_size, n_steps)

    @property
    def variables(self):
        """Named views ``Generator/Generator/...`` of the parameter bucket (train:262 splits by this prefix)."""
        return self._engine.g.views()

    def attentionMechanism(self, cell_state):
        """gen:13-18: ``cell_state`` is the LSTMStateTuple (c, h); only c enters the scores."""
        return self._attention(self._engine.g, cell_state)

    def build_generator(self, images, is_training=True, noise: Optional[torch.Tensor] = None):
        """gen:20,74-91 from ``self.downsampled``.  ``is_training`` is accepted and ignored, as in the reference
        (gen:20 never reads it).  ``noise`` [B,512] may be injected; otherwise N(0,1) is drawn on the device
        (gen:81: one draw per call, shared by all timesteps)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        ann = as_annotations(images, dev)
        B, R = ann.shape[0], ann.numel() // (ann.shape[0] * 512)
        e = self._ensure_engine(B, R)
        self._set_context(ann)
        e.ann_g, e._refresh = ann.view(B, R, 512), True
        if noise is None:
            e.sample_noise()
        else:
            e.noise.copy_(noise.to(device=dev, dtype=torch.float32))
        logits = e.gen_forward().clone()
        self.alpha = e.ws_view("g.EA", (e.T, B, 256), torch.float32)[e.T - 1, :, :R].clone()
        return logits

    def sample(self, images, mode: str = "greedy", noise: Optional[torch.Tensor] = None, chunk: int = 0):
        """Test-time decoding of train:269-270 (``tf.argmax(generator_output, axis=2)``) without materialising the
        logits; ``mode="gumbel"`` draws token ~ softmax(logits) instead (extension).  Returns tokens [B, T] int32."""
        from ..sampling import GeneratorSampler
        dev = torch.device("cuda", torch.cuda.current_device())
        ann = as_annotations(images, dev)
        B, R = ann.shape[0], ann.numel() // (ann.shape[0] * 512)
        e = self._ensure_engine(B, R)
        key = (B, R, chunk)
        if getattr(self, "_sampler_key", None) != key:
            self._sampler = GeneratorSampler(e.g, B, self.n_steps, R, chunk=chunk)
            self._sampler_key = key
        self._set_context(ann)
        return self._sampler.sample(ann.view(B, R, 512), mode=mode, noise=noise).clone()
