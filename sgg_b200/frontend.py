"""The caller-side producer of the annotations: the convolutional front-end of gen:29-68 / disc:29-68.

NOT part of the B200 hot path.  The convolutions and their backward pass are PyTorch library calls (cuDNN on the GPU); the
HBM-bound half of the stack -- the H*W*C layer norm with its ELU, forward and reverse -- runs on this repo's kernels
(``csrc/frontend.cu``, ``sgg_ln_elu_forward`` / ``sgg_ln_elu_backward``) for fp32 CUDA tensors.  What this module adds to the hot path is the seam: ``annotations =
front_end(images)`` feeds ``Engine.set_batch``, and the annotation adjoint the step functions return
(``sgg_step_args_t.ann_g_grad`` / ``ann_d_grad``: d gen_cost / d ann_g, d disc_cost / d ann_d incl. the gradient
penalty's second-order path) is back-propagated through the stack with ``annotations.backward(adjoint)``, so that the
conv variables -- which ARE in the var_lists of train:262-263 -- train as in the reference.

Restated semantics (TensorFlow 1.x, un-vendored; see SURVEY 8c for the pinning caveat):
  * ``tf.layers.conv2d(padding="same")`` (gen:29...): output size ceil(n / stride); total padding
    max((out - 1) * stride + k - n, 0), split floor / ceil between before / after (asymmetric for 56 -> 28 -> 14);
    kernel HWIO ``he_normal`` (VarianceScaling(2, fan_in, truncated normal)), bias constant 0.05 (gen:21-22);
  * ``tf.contrib.layers.layer_norm(activation_fn=tf.nn.elu)`` (gen:30...): moments over all of (H, W, C) per sample,
    variance epsilon 1e-12, gamma / beta of shape [C] (begin_params_axis = -1), then ELU;
  * layer order and widths of gen:29-68; conv3_3 / conv3_4 and their norms (gen:59-62) feed nothing -- their variables are
    created for name parity and take no part in the computation;
  * variable names in creation order: ``conv2d``, ``conv2d_1`` ... ``conv2d_13`` (kernel, bias) and ``LayerNorm`` ...
    ``LayerNorm_12`` (beta, gamma) under ``Generator/Generator/`` resp. ``Discriminator/Discriminator/``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.nn.functional as F

LN_EPS = 1e-12          # tf.contrib.layers.layer_norm variance_epsilon
BIAS_INIT = 0.05        # gen:21

# (name in the reference, in channels, out channels, kernel, stride, input layer (-1 = images), has LayerNorm + ELU, live)
_LAYERS: List[Tuple[str, int, int, int, int, int, bool, bool]] = [
    ("conv1_1", 3, 32, 3, 1, -1, True, True),      # gen:29-30
    ("conv1_2", 32, 32, 3, 1, 0, True, True),      # gen:31-32
    ("conv1_3", 32, 32, 5, 2, 1, True, True),      # gen:35-36  221 -> 111
    ("conv2_1", 32, 64, 3, 1, 2, True, True),      # gen:39-40
    ("conv2_2", 64, 64, 3, 1, 3, True, True),      # gen:41-42
    ("conv2_3", 64, 128, 3, 1, 4, True, True),     # gen:44-45
    ("conv2_4", 128, 128, 3, 1, 5, True, True),    # gen:46-47
    ("conv2_5", 128, 128, 5, 2, 6, True, True),    # gen:50-51  111 -> 56
    ("conv3_1", 128, 256, 3, 1, 7, True, True),    # gen:54-55
    ("conv3_2", 256, 256, 3, 1, 8, True, True),    # gen:56-57
    ("conv3_3", 256, 512, 3, 1, 9, True, False),   # gen:59-60  dead
    ("conv3_4", 512, 512, 3, 1, 10, True, False),  # gen:61-62  dead
    ("conv3_5", 256, 512, 5, 2, 9, True, True),    # gen:65-66  56 -> 28 (reads layernorm3_2)
    ("downsampled", 512, 512, 5, 2, 12, False, True),   # gen:68  28 -> 14, no norm / activation
]


def _tf_suffix(i: int) -> str:
    return "" if i == 0 else f"_{i}"


def same_padding(n: int, k: int, stride: int) -> Tuple[int, int]:
    """TensorFlow "SAME": (before, after) padding of one spatial axis."""
    out = -(-n // stride)
    total = max((out - 1) * stride + k - n, 0)
    return total // 2, total - total // 2


def he_normal_(w: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """tf.keras.initializers.he_normal on an OIHW kernel: truncated normal (|x| <= 2 sigma) whose standard deviation
    after truncation is sqrt(2 / fan_in) (VarianceScaling divides by 0.87962566103423978)."""
    fan_in = w.shape[1] * w.shape[2] * w.shape[3]
    std = math.sqrt(2.0 / fan_in) / 0.87962566103423978
    with torch.no_grad():
        return torch.nn.init.trunc_normal_(w, 0.0, std, -2 * std, 2 * std, generator=generator)


class _LayerNormELU(torch.autograd.Function):
    """tf.contrib.layers.layer_norm(activation_fn=tf.nn.elu) on a channels-last fp32 CUDA tensor through the C ABI
    (csrc/frontend.cu): two passes forward, two passes reverse, x is the only saved activation."""

    @staticmethod
    def forward(ctx, x, gamma, beta):
        import ctypes as C

        from ._lib import check, lib, stream_ptr
        B, Cc, H, W = x.shape
        xs = x.contiguous(memory_format=torch.channels_last)          # storage order [B, H, W, C]
        y = torch.empty_like(xs, memory_format=torch.channels_last)
        stats = torch.empty(B, 2, dtype=torch.float32, device=x.device)
        L = lib()
        L.sgg_ln_elu_scratch_floats.restype = C.c_int64
        n_scr = L.sgg_ln_elu_scratch_floats(C.c_int64(B), C.c_int64(H * W), C.c_int32(Cc))
        if n_scr < 0:
            raise RuntimeError("sgg_ln_elu_scratch_floats: " + L.sgg_last_error().decode())
        scratch = torch.empty(n_scr, dtype=torch.float32, device=x.device)
        g, b = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        check(L.sgg_ln_elu_forward(C.c_void_p(xs.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(b.data_ptr()), C.c_int64(B),
                                   C.c_int64(H * W), C.c_int32(Cc), C.c_float(LN_EPS), C.c_void_p(y.data_ptr()),
                                   C.c_void_p(stats.data_ptr()), C.c_void_p(scratch.data_ptr()), stream_ptr()), "sgg_ln_elu_forward")
        ctx.save_for_backward(xs, g, b, stats)
        ctx.n_scr = n_scr
        return y

    @staticmethod
    def backward(ctx, dy):
        import ctypes as C

        from ._lib import check, lib, stream_ptr
        xs, g, b, stats = ctx.saved_tensors
        B, Cc, H, W = xs.shape
        dys = dy.float().contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(xs, memory_format=torch.channels_last)
        dg, db = torch.empty_like(g), torch.empty_like(b)
        scratch = torch.empty(ctx.n_scr, dtype=torch.float32, device=xs.device)
        check(lib().sgg_ln_elu_backward(C.c_void_p(xs.data_ptr()), C.c_void_p(dys.data_ptr()), C.c_void_p(g.data_ptr()),
                                        C.c_void_p(b.data_ptr()), C.c_void_p(stats.data_ptr()), C.c_int64(B), C.c_int64(H * W),
                                        C.c_int32(Cc), C.c_void_p(dx.data_ptr()), C.c_void_p(dg.data_ptr()), C.c_void_p(db.data_ptr()),
                                        C.c_void_p(scratch.data_ptr()), stream_ptr()), "sgg_ln_elu_backward")
        return dx, dg, db


def layer_norm_elu(y: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, fused: Optional[bool] = None) -> torch.Tensor:
    """gen:30 on an NCHW-shaped tensor: moments over (C, H, W) per sample, per-channel gamma / beta, ELU.  fp32 CUDA tensors go
    through this repo's kernels; anything else (CPU tensors of the tests, autocast dtypes) through torch's group_norm + elu,
    which compute the same function."""
    if fused is None:
        fused = y.is_cuda and y.dtype == torch.float32 and gamma.dtype == torch.float32
    if fused:
        return _LayerNormELU.apply(y, gamma, beta)
    return F.elu(F.group_norm(y, 1, gamma.to(y.dtype), beta.to(y.dtype), LN_EPS))


class ConvFrontEnd(torch.nn.Module):
    """images [B, H, W, 3] (NHWC, standardised as in train:170) -> self.downsampled [B, h, w, 512] (gen:29-68)."""

    def __init__(self, scope: str = "Generator/Generator", seed: Optional[int] = None):
        super().__init__()
        self.scope = scope
        self.fused_norm: Optional[bool] = None      # None: this repo's kernels for fp32 CUDA tensors; False: torch ops
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        self.kernels = torch.nn.ParameterList()
        self.biases = torch.nn.ParameterList()
        self.gammas = torch.nn.ParameterList()
        self.betas = torch.nn.ParameterList()
        for (_, cin, cout, k, _, _, norm, _) in _LAYERS:
            self.kernels.append(torch.nn.Parameter(he_normal_(torch.empty(cout, cin, k, k), g)))
            self.biases.append(torch.nn.Parameter(torch.full((cout,), BIAS_INIT)))
            if norm:
                self.gammas.append(torch.nn.Parameter(torch.ones(cout)))
                self.betas.append(torch.nn.Parameter(torch.zeros(cout)))
        self.to(memory_format=torch.channels_last)

    # ------------------------------------------------------------------ forward
    def _norm_index(self, layer: int) -> int:
        return sum(1 for l in _LAYERS[:layer] if l[6])

    def forward(self, images: torch.Tensor) -> torch.Tensor:
        if images.dim() != 4 or images.shape[-1] != 3:
            raise ValueError(f"expected NHWC images [B, H, W, 3], got {tuple(images.shape)}")
        x = images.permute(0, 3, 1, 2)                       # NCHW view of the NHWC data = channels_last
        outs: Dict[int, torch.Tensor] = {-1: x}
        for i, (_, _, _, k, stride, src, norm, live) in enumerate(_LAYERS):
            if not live:
                continue
            h = outs[src]
            pt, pb = same_padding(h.shape[2], k, stride)
            pl, pr = same_padding(h.shape[3], k, stride)
            if pt == pb and pl == pr:
                y = F.conv2d(h, self.kernels[i].to(h.dtype), self.biases[i].to(h.dtype), stride=stride, padding=(pt, pl))
            else:
                y = F.conv2d(F.pad(h, (pl, pr, pt, pb)), self.kernels[i].to(h.dtype), self.biases[i].to(h.dtype), stride=stride)
            if norm:
                j = self._norm_index(i)
                y = layer_norm_elu(y, self.gammas[j], self.betas[j], self.fused_norm)
            outs[i] = y
        return outs[len(_LAYERS) - 1].permute(0, 2, 3, 1)    # NHWC [B, h, w, 512]

    # ------------------------------------------------------------------ TF variable names / layouts
    def tf_variables(self) -> "OrderedDict[str, torch.Tensor]":
        """Views under the reference's variable names; conv kernels transposed to TF's HWIO layout."""
        out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        for i in range(len(_LAYERS)):
            out[f"{self.scope}/conv2d{_tf_suffix(i)}/kernel"] = self.kernels[i].detach().permute(2, 3, 1, 0)
            out[f"{self.scope}/conv2d{_tf_suffix(i)}/bias"] = self.biases[i].detach()
        for j in range(len(self.gammas)):
            out[f"{self.scope}/LayerNorm{_tf_suffix(j)}/beta"] = self.betas[j].detach()
            out[f"{self.scope}/LayerNorm{_tf_suffix(j)}/gamma"] = self.gammas[j].detach()
        return out

    def load_tf_variables(self, tensors: Dict[str, torch.Tensor], strict: bool = True) -> List[str]:
        """Loads variables named as in a reference checkpoint (``tf_checkpoint.read_checkpoint``); returns the names used."""
        used = []
        with torch.no_grad():
            for name, view in self.tf_variables().items():
                if name not in tensors:
                    if strict:
                        raise KeyError(name)
                    continue
                src = torch.as_tensor(tensors[name], dtype=torch.float32)
                if tuple(src.shape) != tuple(view.shape):
                    raise ValueError(f"{name}: shape {tuple(src.shape)} != {tuple(view.shape)}")
                view.copy_(src)            # a permuted view of the parameter: writes through
                used.append(name)
        return used

    def live_parameters(self) -> Iterable[torch.nn.Parameter]:
        """The variables that receive a gradient (tf.gradients returns None for conv3_3 / conv3_4 and their norms, and
        AdamOptimizer skips those)."""
        for i, l in enumerate(_LAYERS):
            if l[7]:
                yield self.kernels[i]
                yield self.biases[i]
                if l[6]:
                    j = self._norm_index(i)
                    yield self.gammas[j]
                    yield self.betas[j]


class TFAdam:
    """tf.train.AdamOptimizer(1e-4, 0.5, 0.9) (train:258-259) on a list of torch parameters: epsilon outside the bias
    correction, variables without a gradient are skipped.  Same rule as sgg_adam_step; this one serves the front-end's
    variables, which live outside the flat buckets."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-4, beta1=0.5, beta2=0.9, eps=1e-8):
        self.params = list(params)
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]

    @torch.no_grad()
    def step(self) -> None:
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for p, m, v in zip(self.params, self.m, self.v):
            if p.grad is None:
                continue
            m.mul_(self.b1).add_(p.grad, alpha=1.0 - self.b1)
            v.mul_(self.b2).addcmul_(p.grad, p.grad, value=1.0 - self.b2)
            p.addcdiv_(m, v.sqrt().add_(self.eps), value=-lr_t)
            p.grad = None

    def state(self):
        return {"t": self.t, "m": [m.cpu().clone() for m in self.m], "v": [v.cpu().clone() for v in self.v]}

    def load_state(self, st) -> None:
        self.t = int(st["t"])
        for dst, src in zip(self.m + self.v, list(st["m"]) + list(st["v"])):
            dst.copy_(src)


class FrontEndTrainer:
    """train:362-368 with the conv front-ends in the loop: ``CRITIC_ITERS`` x {D step, Adam on Discriminator*} then
    {G step, Adam on Generator*}, all on one (images, labels) batch (train:185-187).  The recurrent half of every step is
    the hot path (``HotPathTrainer``'s engine, step-level C ABI); the convolutions around it are library calls."""

    def __init__(self, trainer, seed: int = 0, compute_dtype: torch.dtype = torch.float32):
        self.tr = trainer                       # HotPathTrainer
        dev = trainer.eng.device
        self.fg = ConvFrontEnd("Generator/Generator", seed=2 * seed + 11).to(dev)
        self.fd = ConvFrontEnd("Discriminator/Discriminator", seed=2 * seed + 12).to(dev)
        self.adam_fg = TFAdam(self.fg.live_parameters(), trainer.lr, trainer.beta1, trainer.beta2, 1e-8)
        self.adam_fd = TFAdam(self.fd.live_parameters(), trainer.lr, trainer.beta1, trainer.beta2, 1e-8)
        self.compute_dtype = compute_dtype

    def _annotations(self, net: ConvFrontEnd, images: torch.Tensor, grad: bool) -> torch.Tensor:
        with torch.set_grad_enabled(grad):
            if self.compute_dtype == torch.float32:
                return net(images)
            with torch.autocast("cuda", dtype=self.compute_dtype):
                return net(images).float()

    def annotations(self, images: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(ann_g, ann_d) [B,R,512] bf16 of a batch of images, no gradient (evaluation, train.py:269-275)."""
        B = images.shape[0]
        return tuple(self._annotations(net, images, grad=False).reshape(B, -1, 512).to(torch.bfloat16).contiguous()
                     for net in (self.fg, self.fd))

    def validation_cost(self, images: torch.Tensor, labels: torch.Tensor) -> float:
        """disc_cost on a held-out batch (train.py:380), nothing is updated."""
        eng = self.tr.eng
        ann_g, ann_d = self.annotations(images)
        eng.set_batch(ann_g, ann_d, labels)
        eng.sample_noise(); eng.sample_gp_alpha()
        eng.disc_step()
        sc = eng.scalars.tolist()
        return sc[1] + self.tr.lam * sc[2]

    def iteration(self, images: torch.Tensor, labels: torch.Tensor) -> Dict[str, list]:
        """One pass of the reference loop body on device tensors ``images`` [B,221,221,3] fp32, ``labels`` [B,T] int64."""
        tr, eng = self.tr, self.tr.eng
        B = images.shape[0]
        log = {"disc_cost": [], "gp": [], "gen_cost": None}
        ann_g = self._annotations(self.fg, images, grad=False)         # the generator is a constant of the D steps
        ann_g16 = ann_g.reshape(B, -1, 512).to(torch.bfloat16).contiguous()
        for _ in range(tr.critic_iters):
            ann_d = self._annotations(self.fd, images, grad=True)      # every sess.run recomputes the stack (train:365)
            eng.set_batch(ann_g16, ann_d.detach().reshape(B, -1, 512).to(torch.bfloat16).contiguous(), labels)
            eng.sample_noise(); eng.sample_gp_alpha()
            a_bar = eng.disc_step(ann_grad=True)
            ann_d.backward(a_bar.view_as(ann_d))
            eng.d.adam_step(tr.lr, tr.beta1, tr.beta2, 1e-8)
            self.adam_fd.step()
            sc = eng.scalars.tolist()
            log["disc_cost"].append(sc[1] + tr.lam * sc[2]); log["gp"].append(sc[2])
        ann_g = self._annotations(self.fg, images, grad=True)
        ann_d = self._annotations(self.fd, images, grad=False)
        eng.set_batch(ann_g.detach().reshape(B, -1, 512).to(torch.bfloat16).contiguous(),
                      ann_d.reshape(B, -1, 512).to(torch.bfloat16).contiguous(), labels)
        eng.sample_noise()
        a_bar = eng.gen_step(ann_grad=True)
        ann_g.backward(a_bar.view_as(ann_g))
        eng.g.adam_step(tr.lr, tr.beta1, tr.beta2, 1e-8)
        self.adam_fg.step()
        log["gen_cost"] = float(eng.scalars[3])
        return log
