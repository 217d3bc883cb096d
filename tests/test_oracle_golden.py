"""CPU: pin the oracle (oracle/renderer_oracle.py) against outputs of the unmodified
reference (tests/golden/*.npz, made by tests/golden/make_golden.py).  The reference
has no test vectors of its own; these are the pins."""
import json

import numpy as np
import pytest
import torch

from oracle import renderer_oracle as orc
from tests.golden.cases import RENDER_CASES, POINT_CASES, build_case_inputs
from tests.helpers import load_golden, oracle_render, input_checksum

TOL = 2e-5   # fp32 CPU vs fp32 CPU, different op order


@pytest.mark.parametrize('name', sorted(RENDER_CASES))
def test_render_matches_reference(name):
    g = load_golden('render', name)
    assert g['case'] == json.loads(json.dumps(RENDER_CASES[name])), 'fixture is stale: rerun tests/golden/make_golden.py'
    assert abs(input_checksum(g['case']) - float(g['input_checksum'])) < 1e-6 * max(1.0, abs(float(g['input_checksum'])))
    rgb, depth, wsum, xyz = oracle_render(g['case'])
    st = int(g['case'].get('store_stride', 1))          # big cases keep every st-th ray of the reference output
    for got, key in ((rgb, 'rgb'), (depth, 'depth'), (wsum, 'wsum'), (xyz, 'xyz')):
        err = (got[:, ::st] - g[key]).abs().max().item()
        assert err < TOL, f'{name}:{key} max abs err {err}'


@pytest.mark.parametrize('name', ['small_plain', 'small_batch3', 'small_black_eg3dplanes'])
def test_aten_gather_equals_manual_gather(name):
    a = oracle_render(RENDER_CASES[name], gather='aten')
    b = oracle_render(RENDER_CASES[name], gather='manual')
    for x, y in zip(a, b):
        assert (x - y).abs().max().item() < TOL


@pytest.mark.parametrize('name', sorted(POINT_CASES))
def test_run_model_matches_reference(name):
    g = load_golden('points', name)
    case = g['case']
    planes, dec, *_rest, opts = build_case_inputs(case)
    rgb, sigma = orc.run_model(planes, dec, g['pts'], opts, case.get('use_triplane', True))
    assert (rgb - g['rgb']).abs().max().item() < TOL
    assert (sigma - g['sigma']).abs().max().item() < TOL


def test_cameras_and_rays_match_reference():
    z = np.load(__import__('os').path.join(__import__('tests.helpers', fromlist=['GOLDEN']).GOLDEN, 'cameras.npz'))
    for (e, a, d, f), c2w_ref, K_ref in zip(z['cams'], z['c2w'], z['K']):
        c2w, K = orc.camera_params_to_matrix(e, a, d, f)
        assert np.abs(c2w.numpy() - c2w_ref).max() < 1e-6
        assert np.abs(K.numpy() - K_ref).max() < 1e-6
    for i, (e, a, d) in enumerate(z['ortho']):
        ro, rd = orc.rays_ortho(e, a, d, 0.7, 8)
        assert np.abs(ro[0].numpy() - z['ortho_ro'][i]).max() < 1e-6
        assert np.abs(rd[0].numpy() - z['ortho_rd'][i]).max() < 1e-6
    # eval table cam60[spin12]: elev 0, azim -180..150 step 30 reordered (lustrous_renders_v1.py:14-30)
    assert z['spin12'].shape == (12, 2) and np.all(z['spin12'][:, 0] == 0)
    assert sorted(z['spin12'][:, 1].tolist()) == [float(a) for a in range(-180, 180, 30)]


def test_pinhole_rays_in_fixture():
    g = load_golden('render', 'small_plain')
    ro, rd = orc.ray_sampler(g['c2w'], g['K'], g['case']['R'])
    assert (ro - g['ro']).abs().max().item() < 1e-6
    assert (rd - g['rd']).abs().max().item() < 1e-6
