"""CPU: the volume-query restatement (oracle/volume_oracle.py) against fixtures produced by the reference's own
get_eg3d_volume / create_samples / sigma2density (tests/golden/make_golden.py:run_volume_case)."""
import numpy as np
import pytest
import torch

from oracle import volume_oracle as vo
from tests.golden.cases import VOLUME_CASES, build_case_inputs
from tests.helpers import load_golden


@pytest.mark.parametrize('name', sorted(VOLUME_CASES))
def test_volume_oracle_matches_reference_fixture(name):
    g = load_golden('volume', name)
    case = g['case']
    g = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in g.items()}
    planes, dec, _, _, _, _, opts = build_case_inputs(case)
    out = vo.volume(planes, dec, opts, case['res'], triplane_crop=case.get('triplane_crop'), cull_clouds=case.get('cull_clouds'),
                    use_triplane=case.get('use_triplane', True))
    assert np.array_equal(out['coordinates'].numpy(), g['coordinates'])         # the sheared lattice, bit for bit
    for key in ('sigmas', 'rgbs'):
        assert out[key].shape == g[key].shape
        assert np.abs(out[key].numpy() - g[key]).max() < 2e-5, key
    d, dg = out['densities'].numpy(), g['densities']
    assert ((d == -1e3) == (dg == -1e3)).mean() > 0.999                         # masks agree (threshold flips aside)
    keep = (d != -1e3) & (dg != -1e3)
    assert np.abs(d[keep] - dg[keep]).max() < 2e-5


def test_lattice_is_sheared_like_the_reference():
    """y/x 'indices' are float quotients, never floored (eg3d_metrics3d.py:80-82): consecutive points differ in y."""
    s = vo.create_samples(5, 1.0)[0]
    assert s.shape == (125, 3)
    assert torch.allclose(s[:5, 2], torch.linspace(-0.5, 0.5, 5))
    assert not torch.equal(s[0, 1], s[1, 1]) and s[1, 1] - s[0, 1] == pytest.approx(0.25 / 5, abs=1e-6)


def test_volume_module_helpers_match_oracle_on_cpu():
    import panic3d_b200.volume as pv
    a, origin, size = pv.create_samples(7, cube_length=0.7)
    assert torch.equal(a, vo.create_samples(7, 0.7)) and size == pytest.approx(0.7 / 6) and origin[0] == pytest.approx(-0.35)
    x = torch.linspace(-5, 5, 11)
    assert torch.equal(pv.sigma2density(x), vo.sigma2density(x))
    with pytest.raises(RuntimeError, match='no CPU path'):
        pv.query_volume(torch.zeros(1, 3, 32, 8, 8), None, {'box_warp': 0.7})
