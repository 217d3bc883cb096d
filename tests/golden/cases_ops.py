"""Seeded cases for the StyleGAN ops (shared by the golden generator, oracle tests and GPU parity tests)."""
import numpy as np
import torch

BIAS_ACT_CASES = {
    # the variants on panic3d's path: lrelu sqrt2 clamp 256 on conv outputs, linear+clamp in ToRGB, lrelu on FCs
    'lrelu_conv': dict(seed=1, shape=(2, 8, 6, 10), act='lrelu', clamp=256.0, bias=True),
    'lrelu_small_clamp': dict(seed=2, shape=(2, 8, 6, 10), act='lrelu', clamp=0.5, bias=True),
    'linear_torgb': dict(seed=3, shape=(3, 3, 5, 7), act='linear', clamp=0.8, bias=True),
    'lrelu_fc': dict(seed=4, shape=(5, 24), act='lrelu', bias=True),
    'linear_nobias': dict(seed=5, shape=(2, 4, 3, 3), act='linear', gain=2.0, bias=False),
    'relu': dict(seed=6, shape=(2, 5, 4, 4), act='relu', bias=True),
    'tanh': dict(seed=7, shape=(2, 5, 4, 4), act='tanh', bias=True, clamp=0.9),
    'sigmoid': dict(seed=8, shape=(2, 5, 4, 4), act='sigmoid', bias=True),
    'elu': dict(seed=9, shape=(2, 5, 4, 4), act='elu', bias=True),
    'selu': dict(seed=10, shape=(2, 5, 4, 4), act='selu', bias=True),
    'softplus': dict(seed=11, shape=(2, 5, 4, 4), act='softplus', bias=True),
    'swish': dict(seed=12, shape=(2, 5, 4, 4), act='swish', bias=True, clamp=1.5),
    'lrelu_dim2_alpha': dict(seed=13, shape=(2, 3, 7, 5), act='lrelu', alpha=0.1, gain=0.7, dim=2, bias=True),
    'odd_numel': dict(seed=14, shape=(1, 3, 5, 7), act='lrelu', bias=True),
}

F1331 = [1, 3, 3, 1]
UPFIRDN_CASES = {
    # 4x4 [1,3,3,1] outer product at up1/down1 (after transposed conv), up2 (skip image), down2 (D skip)
    'blur_after_tconv': dict(seed=21, shape=(2, 6, 17, 17), filter=F1331, padding=[1, 1, 1, 1], gain=4.0),
    'up2_skip': dict(seed=22, shape=(2, 3, 8, 8), filter=F1331, up=2, padding=[2, 1, 2, 1], gain=4.0),
    'down2_skip': dict(seed=23, shape=(2, 4, 16, 16), filter=F1331, down=2, padding=[1, 1, 1, 1]),
    'blur_before_sconv': dict(seed=24, shape=(1, 5, 16, 16), filter=F1331, padding=[2, 2, 2, 2]),
    'gauss_separable_1d': dict(seed=25, shape=(1, 3, 20, 22), filter='gauss13', padding=[6, 6, 6, 6]),
    'asym_updown_flip': dict(seed=26, shape=(2, 2, 7, 9), filter=[1, 2, 4, 3, 1], up=[3, 2], down=[2, 3], padding=[3, 2, 1, 4], flip=True, gain=1.7),
    'crop_negative_pad': dict(seed=27, shape=(1, 2, 12, 12), filter=F1331, padding=[-1, 2, 3, -2]),
    'identity_filter': dict(seed=28, shape=(1, 2, 5, 6), filter=None, up=2),
    'nonsquare_2d': dict(seed=29, shape=(1, 3, 9, 11), filter='rand3x5', up=2, down=1, padding=[2, 2, 1, 1]),
}

FLRELU_CASES = {
    # StyleGAN3 SynthesisLayer-like settings (networks_stylegan3.py:357): up 2/4, down 2, 12-tap kaiser-ish filters
    'up2_down2_sep': dict(seed=41, shape=(2, 4, 12, 12), fu='rand12', fd='rand12', up=2, down=2, padding=[9, 10, 9, 10], clamp=256.0, bias=True),
    'up4_down2_sep': dict(seed=42, shape=(1, 3, 8, 8), fu='rand24', fd='rand12', up=4, down=2, padding=[17, 18, 17, 18], clamp=1.0, bias=True),
    'up1_down1_none': dict(seed=43, shape=(2, 3, 6, 7), fu=None, fd=None, up=1, down=1, padding=0, bias=True),
    'up2_down1_full2d': dict(seed=44, shape=(1, 2, 7, 6), fu='rand4x4', fd='rand3x3', up=2, down=1, padding=[3, 2, 2, 3], gain=1.3, slope=0.1, bias=True),
    'up1_down2_flip': dict(seed=45, shape=(1, 2, 10, 10), fu=[1, 2, 1], fd=F1331, up=1, down=2, padding=[2, 2, 2, 2], flip=True, bias=False),
}


def make_input(c):
    rng = np.random.default_rng(c['seed'])
    x = torch.from_numpy(rng.standard_normal(c['shape']).astype(np.float32))
    b = None
    if c.get('bias'):
        b = torch.from_numpy(rng.standard_normal((c['shape'][c.get('dim', 1)],)).astype(np.float32))
    return x, b


def make_filter(c, setup_filter):
    f = c.get('filter')
    if f is None:
        return None
    rng = np.random.default_rng(c['seed'] + 500)
    if f == 'gauss13':
        t = np.arange(-6, 7, dtype=np.float64)
        return setup_filter(np.exp(-t ** 2 / (2 * 2.0 ** 2)))               # 13 taps -> separable
    if isinstance(f, str) and f.startswith('rand') and 'x' in f:
        h, w = (int(v) for v in f[4:].split('x'))
        return setup_filter(rng.uniform(0.1, 1.0, (h, w)).astype(np.float32))
    if isinstance(f, str) and f.startswith('rand'):
        return setup_filter(rng.uniform(0.1, 1.0, (int(f[4:]),)).astype(np.float32))   # >= 8 taps -> separable
    return setup_filter(f)
