"""CPU: panic3d_b200's math_utils mirror against outputs of the reference module (tests/golden/math_utils.npz)."""
import os

import numpy as np
import torch

from tests.golden.make_golden_math import math_inputs

HERE = os.path.dirname(os.path.abspath(__file__))


def test_math_utils_equal_the_reference_bit_for_bit():
    from panic3d_b200.training.volumetric_rendering import math_utils as mu
    g = np.load(os.path.join(HERE, 'golden', 'math_utils.npz'))
    o, d, start, stop = math_inputs()
    for box in (0.7, 1.0):
        t0, t1 = mu.get_ray_limits_box(o, d, box_side_length=box)
        assert tuple(t0.shape) == (3, 40, 1) and tuple(t1.shape) == (3, 40, 1)
        assert np.array_equal(t0.numpy(), g[f't0_{box}'], equal_nan=True) and np.array_equal(t1.numpy(), g[f't1_{box}'], equal_nan=True)
    miss = g['t0_0.7'] == -1
    assert 0 < miss.sum() < miss.size and (g['t1_0.7'][miss] == -2).all()                    # both outcomes are present
    assert np.array_equal(mu.linspace(start, stop, 7).numpy(), g['lin7']) and np.array_equal(mu.linspace(start[0], stop[0], 2).numpy(), g['lin2'])
    m = torch.arange(16, dtype=torch.float32).reshape(4, 4) * 0.25 - 1
    assert np.array_equal(mu.transform_vectors(m, torch.cat([o[0], torch.ones(40, 1)], -1)).numpy(), g['transform'])
    assert np.array_equal(mu.normalize_vecs(o[2]).numpy(), g['normalize']) and np.array_equal(mu.torch_dot(o[1], d[1]).numpy(), g['dot'])
    o_req = o.clone().requires_grad_(True)
    assert not mu.get_ray_limits_box(o_req, d, 0.7)[0].requires_grad                         # detached, like the reference
