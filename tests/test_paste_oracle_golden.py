"""CPU: the paste-front oracle (oracle/paste_oracle.py) against outputs of the reference's own ``paste_front``
(tests/golden/paste_*.npz, written by tests/golden/make_golden_paste.py from the unmodified reference)."""
import os

import numpy as np
import pytest
import torch

from oracle import paste_oracle as po
from tests.golden.cases_paste import PASTE_CASES, OUT_KEYS, build_paste_inputs

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    return np.load(os.path.join(HERE, 'golden', f'paste_{name}.npz'))


def run_oracle(case, inp):
    p = {k: v for k, v in case['params'].items() if k not in ('offset_occ', 'mode')}
    return po.paste_front(inp['image'], inp['image_xyz'], inp['image_weights'], inp['front_rgb'], inp['occ'], inp['ro'], inp['rd'],
                          inp['box_warp'], frontw=inp['frontw'], normalize_images=case['normalize_images'], **p)


@pytest.mark.parametrize('name', sorted(PASTE_CASES))
def test_oracle_equals_reference_outputs(name):
    case = PASTE_CASES[name]
    inp = build_paste_inputs(case)
    gold = load(name)
    chk = sum(float(v.double().sum()) for v in inp.values() if torch.is_tensor(v))
    assert abs(chk - float(gold['input_checksum'])) < 1e-6 * max(1.0, abs(chk)), 'seeded inputs drifted'
    out = run_oracle(case, inp)
    for k in OUT_KEYS:
        # same torch CPU ops in the same order as the reference: bit-identical
        assert np.array_equal(out[k].numpy(), gold[k]), k
    ro, rd = po.occlusion_rays(inp['image_xyz'], inp['ray_start'], case['params'].get('offset_occ', 0.01))
    assert np.array_equal(ro.numpy(), gold['occ_ro']) and np.array_equal(rd.numpy(), gold['occ_rd'])


def test_every_mask_has_both_values():
    """The seeded inputs must exercise both sides of every threshold, or the parity tests prove nothing."""
    case = PASTE_CASES['erode4_odd']
    out = run_oracle(case, build_paste_inputs(case))
    for k in ('mask_weights', 'mask_edges', 'mask_occ', 'mask_dxyz', 'mask_frontweight', 'mask'):
        m = out[k]
        assert 0.02 < float((m > 0.5).float().mean()) < 0.99, (k, float(m.mean()))


def test_kornia_restatements_known_answers():
    """Hand-computed cases for the two restated kornia functions (no kornia in this image to compare with)."""
    x = torch.arange(5, dtype=torch.float32)[None, None, None, :].repeat(1, 1, 4, 1) * 2          # ramp: d/dx = 2 per pixel
    m = po.sobel_magnitude(x)
    # normalised Sobel of a ramp with slope 2: gx = 2 * 2 * (1+2+1) / 8 = 2 in the interior, half at the replicated border
    assert torch.allclose(m[0, 0, :, 1:4], torch.full((4, 3), float(np.sqrt(4 + 1e-6))))
    assert torch.allclose(m[0, 0, :, 0], torch.full((4,), float(np.sqrt(1 + 1e-6))))
    b = torch.ones(1, 1, 6, 6); b[0, 0, 2, 3] = 0
    e3 = po.erosion_ones(b, 3)
    want = torch.ones(6, 6); want[1:4, 2:5] = 0                                                  # 3x3 anchored at the centre
    assert torch.equal(e3[0, 0], want)
    e2 = po.erosion_ones(b, 2)                                                                   # 2x2: anchor (1,1) -> window rows y-1..y
    want = torch.ones(6, 6); want[2:4, 3:5] = 0
    assert torch.equal(e2[0, 0], want)
    assert torch.equal(po.erosion_ones(torch.ones(1, 1, 4, 4), 3), torch.ones(1, 1, 4, 4))       # geodesic border never erodes
