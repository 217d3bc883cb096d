"""Golden vectors of ``training/volumetric_rendering/math_utils.py`` from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_math.py

Inputs are regenerated from the seed by ``math_inputs()`` (tests import it); only reference outputs are stored."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def math_inputs():
    g = np.random.Generator(np.random.PCG64(31))
    o = g.uniform(-1.2, 1.2, (3, 40, 3)).astype(np.float32)
    d = g.normal(0, 1, (3, 40, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o[0, :4] = [[0.0, 0.0, -1.0], [0.1, 0.2, -1.0], [0.36, 0.0, -1.0], [0.0, 0.35, 1.0]]       # axis-parallel rays: inside, inside, outside, on a face
    d[0, :4] = [[0.0, 0.0, 1.0], [0.0, 0.0, 1.0], [0.0, 0.0, 1.0], [0.0, 0.0, -1.0]]
    o[1, :2] = [[0.0, 0.0, 0.0], [2.0, 2.0, 2.0]]                                              # origin inside the box; pointing away
    d[1, 1] = [0.57735026, 0.57735026, 0.57735026]
    start = g.uniform(0.3, 0.7, (2, 5, 1)).astype(np.float32)
    stop = start + g.uniform(0.5, 1.0, (2, 5, 1)).astype(np.float32)
    return torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(start), torch.from_numpy(stop)


def main():
    sys.path[:0] = [REF + '/_train/eg3dc/src']
    from training.volumetric_rendering import math_utils as ref
    assert ref.__file__.startswith(REF)
    o, d, start, stop = math_inputs()
    out = {}
    for box in (0.7, 1.0):
        t0, t1 = ref.get_ray_limits_box(o, d, box_side_length=box)
        out[f't0_{box}'], out[f't1_{box}'] = t0.numpy(), t1.numpy()
    out['lin7'] = ref.linspace(start, stop, 7).numpy()
    out['lin2'] = ref.linspace(start[0], stop[0], 2).numpy()
    m = torch.arange(16, dtype=torch.float32).reshape(4, 4) * 0.25 - 1
    out['transform'] = ref.transform_vectors(m, torch.cat([o[0], torch.ones(40, 1)], -1)).numpy()
    out['normalize'] = ref.normalize_vecs(o[2]).numpy()
    out['dot'] = ref.torch_dot(o[1], d[1]).numpy()
    np.savez_compressed(os.path.join(HERE, 'math_utils.npz'), **out)
    print({k: v.shape for k, v in out.items()}, 'misses', int((out['t0_0.7'] == -1).sum()), 'nan', int(np.isnan(out['t0_0.7']).sum()))


if __name__ == '__main__':
    main()
