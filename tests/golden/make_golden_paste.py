"""Golden vectors of the front-view paste from the UNMODIFIED reference (build container only):

    PROJECT_DN=/root/reference python tests/golden/make_golden_paste.py

Runs the reference's own ``paste_front`` (``training/triplane.py:608-691``, with its helpers ``sample_orthofront``,
``get_front_occlusion``, ``get_front_weights``, ``get_xyz_discrepancy``) on a stand-in ``G`` whose ``f`` returns fixed
``image_weights`` for the two extra renders, and stores its outputs.  ``kornia`` (0.6.5 in the reference's Dockerfile) is
not installed here: the module the reference imports is ``oracle.paste_oracle.kornia_shim()`` - a restatement of the two
kornia functions it calls (see that file's header: those two are "parity unpinned").  The rays the occlusion render
receives are stored too, pinning ``get_front_occlusion``'s ray construction.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
os.environ.setdefault('PROJECT_DN', REF)
sys.path[:0] = [ROOT, REF, REF + '/_train/eg3dc/src']

from oracle import paste_oracle as po                                        # noqa: E402
_k = po.kornia_shim()
sys.modules['kornia'] = _k
sys.modules['kornia.filters'] = _k.filters
sys.modules['kornia.morphology'] = _k.morphology

import training.triplane as ref_tp                                           # noqa: E402
from tests.golden.cases_paste import PASTE_CASES, OUT_KEYS, build_paste_inputs  # noqa: E402


class StandInG:
    """What paste_front needs of the generator: rendering_kwargs and f() -> {'image_weights'} for its two renders."""

    def __init__(self, inp):
        self.rendering_kwargs = {'ray_start': inp['ray_start'], 'box_warp': inp['box_warp']}
        self.inp = inp
        self.seen = {}

    def f(self, xin, return_more=False):
        if xin.get('force_rays') is not None:                                # get_front_occlusion (triplane.py:565-580)
            assert xin['paste_params'] is None
            self.seen['occ_ro'] = xin['force_rays']['ray_origins'].clone()
            self.seen['occ_rd'] = xin['force_rays']['ray_directions'].clone()
            return {'image_weights': self.inp['occ']}
        assert float(xin['fovs'][0]) == -1 and 'paste_params' not in xin     # get_front_weights (:581-600)
        return {'image_weights': self.inp['frontw']}


def main():
    for name, case in PASTE_CASES.items():
        inp = build_paste_inputs(case)
        G = StandInG(inp)
        x = {'cond': {'image_ortho_front': inp['front_rgb']}, 'normalize_images': case['normalize_images'],
             'force_rays': {'ray_origins': inp['ro'], 'ray_directions': inp['rd']}, 'paste_params': dict(case['params'])}
        out = {'image': inp['image'].clone(), 'image_xyz': inp['image_xyz'].clone(), 'image_weights': inp['image_weights'].clone()}
        with torch.no_grad():
            res = ref_tp.paste_front(G, x, out, **case['params'])
        store = {k: res[k].numpy() for k in OUT_KEYS}
        store['occ_ro'] = G.seen['occ_ro'].numpy(); store['occ_rd'] = G.seen['occ_rd'].numpy()
        store['input_checksum'] = np.float64(sum(float(v.double().sum()) for v in inp.values() if torch.is_tensor(v)))
        # fp16 would lose the borderline information; masks are 0/1 or bilinear blends of 0/1 -> compress well
        np.savez_compressed(os.path.join(HERE, f'paste_{name}.npz'), **store)
        print(name, {k: tuple(v.shape) for k, v in store.items() if k in ('image', 'mask')},
              'mask mean', float(res['mask'].mean()), [round(float(res[k].mean()), 3) for k in OUT_KEYS[3:]])


if __name__ == '__main__':
    main()
