"""Paste-front parity cases.  Inputs are regenerated from the seed by ``oracle.paste_oracle.synth_paste_inputs``;
only reference OUTPUTS are stored (``paste_<name>.npz``, written by ``make_golden_paste.py``)."""

# thresholds of the eval script (_scripts/eval/generate.py:59-65) and of the training modes 'A' / 'Agrad'
# (loss_orthocondA.py:131-150): thresh_dxyz 5e-6; paste_front's own default is 0.01 (triplane.py:613)
EVAL = dict(thresh_weight=0.95, thresh_edges=0.02, thresh_occ=0.05, offset_occ=0.01, thresh_dxyz=0.000005)

PASTE_CASES = {
    # name: seed, N views, R = neural rendering resolution, S = front image / output side, kwargs of paste_front
    'eval_small': dict(seed=11, N=2, R=24, S=96, normalize_images=False, params=dict(EVAL)),
    'normalized': dict(seed=12, N=1, R=16, S=80, normalize_images=True, params=dict(EVAL)),
    'defaults_erode3': dict(seed=13, N=2, R=32, S=64, normalize_images=False, params=dict(front_weight_erosion=3)),
    'erode4_odd': dict(seed=14, N=1, R=20, S=50, normalize_images=False, params=dict(EVAL, front_weight_erosion=4)),
    'same_res': dict(seed=15, N=1, R=48, S=48, normalize_images=False, params=dict(EVAL, thresh_dxyz=0.01)),
    'eval_mid': dict(seed=16, N=1, R=64, S=160, normalize_images=False, params=dict(EVAL)),
}

OUT_KEYS = ['image', 'paste', 'mask', 'mask_weights', 'mask_edges', 'mask_occ', 'mask_dxyz', 'mask_frontweight']


def build_paste_inputs(case):
    from oracle.paste_oracle import synth_paste_inputs
    return synth_paste_inputs(case['seed'], case['N'], case['R'], case['S'])
