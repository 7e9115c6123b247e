"""CPU restatement of the reference's convolutional front-end (TEST INFRASTRUCTURE -- never imported by the product).

Follows architectures/generator_with_attention.py:21-68 (the discriminator's disc:21-68 is byte-identical) in plain
numpy, NHWC / HWIO as TensorFlow lays the tensors out, independent of torch's convolution and group-norm kernels:

  tf.layers.conv2d(padding="same", strides=s)      gen:29,31,35,...   -> conv2d_same
  tf.contrib.layers.layer_norm(activation_fn=elu)  gen:30,32,36,...   -> layer_norm_elu (moments over H, W, C per sample,
                                                                         epsilon 1e-12, gamma / beta over C)
  layer wiring                                      gen:29-68          -> front_end (conv3_3 / conv3_4, gen:59-62, are dead)

Parity unpinned at the TensorFlow boundary (no TF in this image, the reference holds no fixtures for this path): the
semantics above are restated from the TF 1.x documentation of the two ops, as SURVEY 8c does for the recurrent half.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

LN_EPS = 1e-12

# (conv index in creation order, kernel, stride, source layer (-1 = images), layer norm index or -1)   gen:29-68
WIRING: List[Tuple[int, int, int, int, int]] = [
    (0, 3, 1, -1, 0), (1, 3, 1, 0, 1), (2, 5, 2, 1, 2),                                   # block 1      gen:29-36
    (3, 3, 1, 2, 3), (4, 3, 1, 3, 4), (5, 3, 1, 4, 5), (6, 3, 1, 5, 6), (7, 5, 2, 6, 7),  # block 2      gen:39-51
    (8, 3, 1, 7, 8), (9, 3, 1, 8, 9),                                                     # block 3      gen:54-57
    (12, 5, 2, 9, 12),                                                                    # conv3_5      gen:65-66
    (13, 5, 2, 12, -1),                                                                   # downsampled  gen:68
]


def suffix(i: int) -> str:
    return "" if i == 0 else f"_{i}"


def conv2d_same(x: np.ndarray, kernel: np.ndarray, bias: np.ndarray, stride: int) -> np.ndarray:
    """x [B,H,W,Cin], kernel [kh,kw,Cin,Cout] (HWIO), TensorFlow SAME padding (extra pixel at the bottom / right)."""
    B, H, W, _ = x.shape
    kh, kw, _, cout = kernel.shape
    oh, ow = -(-H // stride), -(-W // stride)
    ph = max((oh - 1) * stride + kh - H, 0)
    pw = max((ow - 1) * stride + kw - W, 0)
    xp = np.pad(x, ((0, 0), (ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0)))
    y = np.zeros((B, oh, ow, cout), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, i:i + (oh - 1) * stride + 1:stride, j:j + (ow - 1) * stride + 1:stride, :]
            y += patch @ kernel[i, j]
    return y + bias


def layer_norm_elu(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray) -> np.ndarray:
    mean = x.mean(axis=(1, 2, 3), keepdims=True)
    var = ((x - mean) ** 2).mean(axis=(1, 2, 3), keepdims=True)          # tf.nn.moments: biased
    y = (x - mean) / np.sqrt(var + LN_EPS) * gamma + beta
    return np.where(y > 0, y, np.expm1(np.minimum(y, 0)))                # tf.nn.elu


def front_end(variables: Dict[str, np.ndarray], scope: str, images: np.ndarray) -> np.ndarray:
    """images [B,H,W,3] -> self.downsampled [B,h,w,512] with the variables of `scope` under their TF names."""
    outs = {-1: images}
    for conv, _, stride, src, ln in WIRING:
        y = conv2d_same(outs[src], variables[f"{scope}/conv2d{suffix(conv)}/kernel"],
                        variables[f"{scope}/conv2d{suffix(conv)}/bias"], stride)
        if ln >= 0:
            y = layer_norm_elu(y, variables[f"{scope}/LayerNorm{suffix(ln)}/gamma"], variables[f"{scope}/LayerNorm{suffix(ln)}/beta"])
        outs[conv] = y
    return outs[13]
