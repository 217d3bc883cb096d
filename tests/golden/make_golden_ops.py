"""Golden vectors for the three StyleGAN ops: outputs of the reference's own impl='ref' code paths
(ops/bias_act.py:93, ops/upfirdn2d.py:169, ops/filtered_lrelu.py:123) on seeded inputs.

    PROJECT_DN=/root/reference python tests/golden/make_golden_ops.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
os.environ.setdefault('PROJECT_DN', REF)
sys.path[:0] = [ROOT, REF + '/_train/eg3dc/src']

from torch_utils.ops import bias_act as ref_ba, upfirdn2d as ref_up, filtered_lrelu as ref_fl   # noqa: E402
from tests.golden.cases_ops import BIAS_ACT_CASES, UPFIRDN_CASES, FLRELU_CASES, make_input, make_filter  # noqa: E402


def main():
    out = {}
    for name, c in BIAS_ACT_CASES.items():
        x, b = make_input(c)
        y = ref_ba.bias_act(x, b, dim=c.get('dim', 1), act=c['act'], alpha=c.get('alpha'), gain=c.get('gain'), clamp=c.get('clamp'), impl='ref')
        out['ba_' + name] = y.numpy()
    for name, c in UPFIRDN_CASES.items():
        x, _ = make_input(c)
        f = make_filter(c, ref_up.setup_filter)
        y = ref_up.upfirdn2d(x, f, up=c.get('up', 1), down=c.get('down', 1), padding=c.get('padding', 0),
                             flip_filter=c.get('flip', False), gain=c.get('gain', 1), impl='ref')
        out['up_' + name] = y.numpy()
    for name, c in FLRELU_CASES.items():
        x, b = make_input(c)
        fu = make_filter(dict(c, filter=c.get('fu')), ref_up.setup_filter)
        fd = make_filter(dict(c, filter=c.get('fd')), ref_up.setup_filter)
        y = ref_fl.filtered_lrelu(x, fu=fu, fd=fd, b=b, up=c.get('up', 1), down=c.get('down', 1), padding=c.get('padding', 0),
                                  gain=c.get('gain', np.sqrt(2)), slope=c.get('slope', 0.2), clamp=c.get('clamp'),
                                  flip_filter=c.get('flip', False), impl='ref')
        out['fl_' + name] = y.numpy()
    # helpers built on upfirdn2d (upfirdn2d.py:279-389)
    x, _ = make_input(dict(seed=77, shape=(2, 3, 9, 8)))
    f = ref_up.setup_filter([1, 3, 3, 1])
    out['hl_filter2d'] = ref_up.filter2d(x, f, impl='ref').numpy()
    out['hl_upsample2d'] = ref_up.upsample2d(x, f, up=2, impl='ref').numpy()
    out['hl_downsample2d'] = ref_up.downsample2d(x, f, down=2, impl='ref').numpy()
    np.savez_compressed(os.path.join(HERE, 'ops_golden.npz'), **out)
    print('wrote', len(out), 'arrays')


if __name__ == '__main__':
    main()
