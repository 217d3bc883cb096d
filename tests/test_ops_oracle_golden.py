"""CPU: pin oracle/ops_oracle.py against the reference's impl='ref' outputs (tests/golden/ops_golden.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import ops_oracle as oo
from tests.golden.cases_ops import BIAS_ACT_CASES, UPFIRDN_CASES, FLRELU_CASES, make_input, make_filter

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ops_golden.npz'))
TOL = 1e-5


@pytest.mark.parametrize('name', sorted(BIAS_ACT_CASES))
def test_bias_act(name):
    c = BIAS_ACT_CASES[name]
    x, b = make_input(c)
    y = oo.bias_act(x, b, dim=c.get('dim', 1), act=c['act'], alpha=c.get('alpha'), gain=c.get('gain'), clamp=c.get('clamp'))
    assert np.abs(y.numpy() - G['ba_' + name]).max() < TOL


@pytest.mark.parametrize('name', sorted(UPFIRDN_CASES))
def test_upfirdn2d(name):
    c = UPFIRDN_CASES[name]
    x, _ = make_input(c)
    f = make_filter(c, oo.setup_filter)
    y = oo.upfirdn2d(x, f, up=c.get('up', 1), down=c.get('down', 1), padding=c.get('padding', 0), flip_filter=c.get('flip', False), gain=c.get('gain', 1))
    assert y.shape == G['up_' + name].shape
    assert np.abs(y.numpy() - G['up_' + name]).max() < TOL


@pytest.mark.parametrize('name', sorted(FLRELU_CASES))
def test_filtered_lrelu(name):
    c = FLRELU_CASES[name]
    x, b = make_input(c)
    fu = make_filter(dict(c, filter=c.get('fu')), oo.setup_filter)
    fd = make_filter(dict(c, filter=c.get('fd')), oo.setup_filter)
    y = oo.filtered_lrelu(x, fu, fd, b, up=c.get('up', 1), down=c.get('down', 1), padding=c.get('padding', 0),
                          gain=c.get('gain', np.sqrt(2)), slope=c.get('slope', 0.2), clamp=c.get('clamp'), flip_filter=c.get('flip', False))
    assert y.shape == G['fl_' + name].shape
    assert np.abs(y.numpy() - G['fl_' + name]).max() < 2e-5
