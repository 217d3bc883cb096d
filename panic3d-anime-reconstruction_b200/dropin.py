"""Plug the B200 hot path into an unmodified checkout of the reference.

The reference instantiates the renderer by import name
(`from training.volumetric_rendering.renderer import ImportanceRenderer`, triplane.py:14-15, and again when
`load_eg3dc_model` rebuilds G from current source, _train/eg3dc/util/eg3dc_v0.py:46-52) and its ops by
`from torch_utils.ops import bias_act, upfirdn2d, ...`.  `install()` registers this package's modules under
those names in `sys.modules` BEFORE the reference imports them, so `G.synthesis`, `G.f`, `G.sample_mixed`,
`_scripts.eval.generate` and the training loop call into the CUDA library with no caller edits.

    import panic3d_b200.dropin as dropin
    dropin.install()                    # before `import training.triplane` / `load_eg3dc_model(...)`
"""
from __future__ import annotations

import importlib
import sys

_RENDER = {
    'training.volumetric_rendering.renderer': '.training.volumetric_rendering.renderer',
    'training.volumetric_rendering.ray_sampler': '.training.volumetric_rendering.ray_sampler',
    'training.volumetric_rendering.ray_marcher': '.training.volumetric_rendering.ray_marcher',
    'training.volumetric_rendering.math_utils': '.training.volumetric_rendering.math_utils',
}
_OPS = {
    'torch_utils.ops.bias_act': '.torch_utils.ops.bias_act',
    'torch_utils.ops.upfirdn2d': '.torch_utils.ops.upfirdn2d',
    'torch_utils.ops.filtered_lrelu': '.torch_utils.ops.filtered_lrelu',
}


def install(renderer: bool = True, ops: bool = True, strict: bool = False):
    """Register the drop-in modules. Returns the list of reference module names now served by this package.
    `strict=True` raises if the reference already imported one of them (too late to swap by name)."""
    pkg = __name__.rsplit('.', 1)[0]
    table = {}
    if renderer:
        table.update(_RENDER)
    if ops:
        table.update(_OPS)
    done = []
    for ref_name, ours in table.items():
        try:
            mod = importlib.import_module(ours, pkg)
        except ModuleNotFoundError:
            if strict:
                raise
            continue
        prev = sys.modules.get(ref_name)
        if prev is not None and prev is not mod and strict:
            raise RuntimeError(f'{ref_name} was imported before panic3d_b200.dropin.install(); call install() first')
        sys.modules[ref_name] = mod
        parent_name, leaf = ref_name.rsplit('.', 1)
        parent = sys.modules.get(parent_name)
        if parent is not None:
            setattr(parent, leaf, mod)
        done.append(ref_name)
    return done


def patch_generator(G):
    """For a TriPlaneGenerator that was built BEFORE install() (e.g. unpickled): swap its two hot-path
    sub-modules in place.  Both are parameter-free, so no weights move."""
    from .training.volumetric_rendering.renderer import ImportanceRenderer
    from .training.volumetric_rendering.ray_sampler import RaySampler
    use_triplane = bool(G.rendering_kwargs.get('use_triplane', False))
    G.renderer = ImportanceRenderer(use_triplane=use_triplane)
    G.ray_sampler = RaySampler()
    return G
