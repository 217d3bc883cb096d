"""GPU parity of the StyleGAN ops (through the C-ABI) vs the reference impl='ref' fixtures and vs the CPU oracle,
forward + first/second-order gradients (R1 regularisation double-backward goes through bias_act and upfirdn2d)."""
import os

import numpy as np
import pytest
import torch

from oracle import ops_oracle as oo
from tests.golden.cases_ops import BIAS_ACT_CASES, UPFIRDN_CASES, FLRELU_CASES, make_input, make_filter

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ops_golden.npz'))
DEV = 'cuda:0'


def _ba():
    from panic3d_b200.torch_utils.ops import bias_act
    return bias_act


def _f32_scalars(kw, act):
    """alpha / gain / clamp cross the C ABI as fp32 (as in the reference plugin: bias_act.h:27-29); give the fp64
    oracle the same rounded scalars so fp64 comparisons can be tight."""
    _, def_alpha, def_gain = oo.ACTS[act]
    out = dict(kw)
    out['alpha'] = float(np.float32(def_alpha if kw.get('alpha') is None else kw['alpha']))
    out['gain'] = float(np.float32(def_gain if kw.get('gain') is None else kw['gain']))
    if kw.get('clamp') is not None:
        out['clamp'] = float(np.float32(kw['clamp']))
    return out


# ------------------------------------------------------------------------------------------ bias_act
@pytest.mark.parametrize('name', sorted(BIAS_ACT_CASES))
def test_bias_act_forward_matches_reference_fixture(name):
    c = BIAS_ACT_CASES[name]
    x, b = make_input(c)
    kw = dict(dim=c.get('dim', 1), act=c['act'], alpha=c.get('alpha'), gain=c.get('gain'), clamp=c.get('clamp'))
    y = _ba().bias_act(x.to(DEV), None if b is None else b.to(DEV), **kw)
    assert y.shape == x.shape and y.dtype == x.dtype
    assert np.abs(y.cpu().numpy() - G['ba_' + name]).max() < 1e-5


@pytest.mark.parametrize('dtype,tol', [(torch.float16, 2e-3), (torch.bfloat16, 2e-2), (torch.float64, 1e-12)])
@pytest.mark.parametrize('name', ['lrelu_conv', 'linear_torgb', 'swish', 'softplus'])
def test_bias_act_dtypes(name, dtype, tol):
    c = BIAS_ACT_CASES[name]
    x, b = make_input(c)
    kw = dict(dim=c.get('dim', 1), act=c['act'], alpha=c.get('alpha'), gain=c.get('gain'), clamp=c.get('clamp'))
    xd, bd = x.to(dtype), b.to(dtype)
    y = _ba().bias_act(xd.to(DEV), bd.to(DEV), **kw)
    ref = oo.bias_act(xd.double(), bd.double(), **_f32_scalars(kw, c['act']))
    assert y.dtype == dtype
    assert (y.cpu().double() - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize('name', ['lrelu_conv', 'linear_torgb', 'tanh'])
def test_bias_act_channels_last_and_unaligned(name):
    c = BIAS_ACT_CASES[name]
    x, b = make_input(c)
    kw = dict(dim=1, act=c['act'], clamp=c.get('clamp'))
    ref = oo.bias_act(x, b, **kw)
    xc = x.to(DEV).contiguous(memory_format=torch.channels_last)
    y = _ba().bias_act(xc, b.to(DEV), **kw)
    assert y.stride() == xc.stride()
    assert (y.cpu() - ref).abs().max().item() < 1e-5
    # misaligned storage offset -> scalar path
    buf = torch.zeros(x.numel() + 1, device=DEV)
    xs = buf[1:].view(x.shape)
    xs.copy_(x)
    y2 = _ba().bias_act(xs, b.to(DEV), **kw)
    assert (y2.cpu() - ref).abs().max().item() < 1e-5


@pytest.mark.parametrize('name', sorted(BIAS_ACT_CASES))
def test_bias_act_gradients(name):
    """dx, db and (where the activation has one) the double-backward, vs torch.autograd on the oracle (fp64 inputs)."""
    c = BIAS_ACT_CASES[name]
    x, b = make_input(c)
    kw = dict(dim=c.get('dim', 1), act=c['act'], alpha=c.get('alpha'), gain=c.get('gain'), clamp=c.get('clamp'))
    rng = np.random.default_rng(c['seed'] + 99)
    dy = torch.from_numpy(rng.standard_normal(c['shape']))
    ddx = torch.from_numpy(rng.standard_normal(c['shape']))

    def run(fn, xin, bin_, dy_, ddx_):
        xin = xin.clone().requires_grad_(True)
        bin_ = None if bin_ is None else bin_.clone().requires_grad_(True)
        y = fn(xin, bin_, **kw)
        ins = [xin] + ([bin_] if bin_ is not None else [])
        g1 = torch.autograd.grad(y, ins, dy_, create_graph=True)
        # R1-style second order: differentiate <dx, ddx> w.r.t. the incoming gradient path and x
        s = (g1[0] * ddx_).sum()
        g2 = torch.autograd.grad(s, ins, allow_unused=True) if s.requires_grad else [None] * len(ins)
        return [y] + list(g1) + [g for g in g2]

    kw_gpu, kw = kw, _f32_scalars(kw, c['act'])
    ref = run(oo.bias_act, x.double(), None if b is None else b.double(), dy, ddx)
    kw = kw_gpu
    got = run(_ba().bias_act, x.double().to(DEV), None if b is None else b.double().to(DEV), dy.to(DEV), ddx.to(DEV))
    for r, g_ in zip(ref, got):
        if r is None:
            assert g_ is None or g_.abs().max().item() == 0
            continue
        assert g_ is not None
        assert (g_.cpu() - r).abs().max().item() < 1e-9 * max(1.0, r.abs().max().item()), name


def test_bias_act_empty_and_errors():
    ba = _ba()
    y = ba.bias_act(torch.zeros(0, 4, device=DEV), torch.zeros(4, device=DEV), act='lrelu')
    assert y.shape == (0, 4)
    with pytest.raises(RuntimeError):
        ba.bias_act(torch.zeros(2, 4), torch.zeros(4))                       # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        ba.bias_act(torch.zeros(2, 4, device=DEV), torch.zeros(5, device=DEV))  # wrong bias length
