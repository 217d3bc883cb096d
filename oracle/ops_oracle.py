"""CPU oracle for the StyleGAN custom ops (bias_act, upfirdn2d, filtered_lrelu).

TEST INFRASTRUCTURE ONLY (see oracle/renderer_oracle.py for the rules).  Plain torch restatements of the
reference's own slow reference implementations:

* ``bias_act``        <- /root/reference/_train/eg3dc/src/torch_utils/ops/bias_act.py:93-122  (_bias_act_ref)
* ``upfirdn2d``       <- .../ops/upfirdn2d.py:169-213                                         (_upfirdn2d_ref)
* ``setup_filter``    <- .../ops/upfirdn2d.py:72-116
* ``filtered_lrelu``  <- .../ops/filtered_lrelu.py:123-155                                    (_filtered_lrelu_ref)

They are differentiable torch code, so first- and second-order gradients of the CUDA ops are checked against
``torch.autograd`` through these.  Pinned against the reference's ``impl='ref'`` outputs by
``tests/golden/make_golden_ops.py`` -> ``tests/golden/ops_*.npz`` (checked in tests/test_ops_oracle_golden.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

_SQRT2 = math.sqrt(2)
ACTS = {
    'linear': (lambda x, a: x, 0.0, 1.0),
    'relu': (lambda x, a: F.relu(x), 0.0, _SQRT2),
    'lrelu': (lambda x, a: F.leaky_relu(x, a), 0.2, _SQRT2),
    'tanh': (lambda x, a: torch.tanh(x), 0.0, 1.0),
    'sigmoid': (lambda x, a: torch.sigmoid(x), 0.0, 1.0),
    'elu': (lambda x, a: F.elu(x), 0.0, 1.0),
    'selu': (lambda x, a: F.selu(x), 0.0, 1.0),
    'softplus': (lambda x, a: F.softplus(x), 0.0, 1.0),
    'swish': (lambda x, a: torch.sigmoid(x) * x, 0.0, _SQRT2),
}


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    fn, def_alpha, def_gain = ACTS[act]
    alpha = float(def_alpha if alpha is None else alpha)
    gain = float(def_gain if gain is None else gain)
    if b is not None:
        x = x + b.reshape([-1 if i == dim else 1 for i in range(x.ndim)])
    x = fn(x, alpha)
    if gain != 1:
        x = x * gain
    if clamp is not None and clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


def setup_filter(f, normalize=True, flip_filter=False, gain=1, separable=None):
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    if f.ndim == 0:
        f = f[None]
    if separable is None:
        separable = f.ndim == 1 and f.numel() >= 8
    if f.ndim == 1 and not separable:
        f = torch.outer(f, f)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    return f * (gain ** (f.ndim / 2))


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _pad4(p):
    if isinstance(p, int):
        p = [p, p]
    p = list(p)
    if len(p) == 2:
        p = [p[0], p[0], p[1], p[1]]
    return p


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    """zero-insert upsample -> pad/crop -> FIR (true convolution unless flip_filter) -> decimate."""
    N, C, H, W = x.shape
    upx, upy = _pair(up)
    dnx, dny = _pair(down)
    px0, px1, py0, py1 = _pad4(padding)
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32)
    y = x.reshape(N, C, H, 1, W, 1)
    y = F.pad(y, [0, upx - 1, 0, 0, 0, upy - 1]).reshape(N, C, H * upy, W * upx)
    y = F.pad(y, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    y = y[:, :, max(-py0, 0): y.shape[2] - max(-py1, 0), max(-px0, 0): y.shape[3] - max(-px1, 0)]
    f = (f * (gain ** (f.ndim / 2))).to(x.dtype)
    if not flip_filter:
        f = f.flip(list(range(f.ndim)))
    if f.ndim == 2:
        y = F.conv2d(y, f[None, None].repeat(C, 1, 1, 1), groups=C)
    else:
        y = F.conv2d(y, f[None, None, None, :].repeat(C, 1, 1, 1), groups=C)
        y = F.conv2d(y, f[None, None, :, None].repeat(C, 1, 1, 1), groups=C)
    return y[:, :, ::dny, ::dnx]


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=_SQRT2, slope=0.2, clamp=None, flip_filter=False):
    px0, px1, py0, py1 = _pad4(padding)
    y = bias_act(x, b)
    y = upfirdn2d(y, fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    y = bias_act(y, act='lrelu', alpha=slope, gain=gain, clamp=clamp)
    return upfirdn2d(y, fd, down=down, flip_filter=flip_filter)
