"""``Discriminator`` with the reference's interface (architectures/discriminator_with_attention.py:7-93).

    d = Discriminator(vocab_size, embedding_matrix)            # embedding_matrix [V, 300], trainable (train:70)
    scores = d.build_discriminator(input_triples, annotations)  # [B, 3, 1]

``input_triples`` is [B, 3, V] float: one-hot reals (train:173) or the generator's raw logits used as a soft
one-hot through ``matmul(indices, W)`` (disc:86-87).
"""
from __future__ import annotations

import torch

from ._common import AttentionNet, as_annotations


class Discriminator(AttentionNet):
    def __init__(self, vocab_size, embedding_matrix, n_steps: int = 3):
        super().__init__(vocab_size, n_steps)
        self.embedding_matrix = embedding_matrix        # disc:11 keeps the caller's variable

    @property
    def variables(self):
        """Named views ``Discriminator/Discriminator/...`` and ``Discriminator/W`` of the parameter bucket."""
        return self._engine.d.views()

    def attentionMechanism(self, cell_state):
        """disc:13-18 (byte-identical to the generator's)."""
        return self._attention(self._engine.d, cell_state)

    def build_discriminator(self, input_triples, images, is_training=True):
        """disc:20,73-93 from ``self.downsampled``; ``is_training`` is ignored as in the reference."""
        dev = torch.device("cuda", torch.cuda.current_device())
        ann = as_annotations(images, dev)
        B, R = ann.shape[0], ann.numel() // (ann.shape[0] * 512)
        emb = self.embedding_matrix
        emb = torch.as_tensor(emb, dtype=torch.float32) if emb is not None else None
        e = self._ensure_engine(B, R, embedding=emb)
        self._set_context(ann)
        e.ann_d = ann.view(B, R, 512)
        tri = torch.as_tensor(input_triples).to(device=dev, dtype=torch.float32).contiguous()
        if tri.shape != (B, e.T, self.vocab_size):
            raise ValueError(f"input_triples must be [B,{e.T},{self.vocab_size}], got {tuple(tri.shape)}")
        scores = e.disc_forward(tri)
        self.alpha = e.ws_view("d.EA", (e.T, B, 256), torch.float32)[e.T - 1, :, :R].clone()   # NR = B in this call
        return scores.unsqueeze(-1)                                                            # disc:92 [B,3,1]
