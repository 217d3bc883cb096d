"""Generate golden vectors by running the UNMODIFIED reference in the build container.

    PROJECT_DN=/root/reference python tests/golden/make_golden.py

The reference (``/root/reference``) is imported exactly as SURVEY.md section 8c
describes; ``torch.rand_like`` / ``torch.rand`` are patched for the duration of
a call so the stratified / importance jitter the reference draws is the tensor
we inject (the same tensor the oracle and the CUDA kernels receive).  Inputs are
NOT stored: they are regenerated from a numpy-PCG64 seed by
``oracle.renderer_oracle.synth_inputs`` (an input checksum is stored to catch
RNG drift).  Only reference *outputs* are committed, as ``tests/golden/*.npz``.

``/root/reference`` exists only in the build container; nothing in ``tests/``,
``bench.py`` or ``smoke()`` reads it at run time - they read the fixtures.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
os.environ.setdefault('PROJECT_DN', REF)
sys.path[:0] = [ROOT, REF, REF + '/_train/eg3dc/src']
sys.modules.setdefault('kornia', types.ModuleType('kornia'))

import training.triplane as ref_tp                                           # noqa: E402
from training.volumetric_rendering.renderer import ImportanceRenderer        # noqa: E402
from training.volumetric_rendering.ray_sampler import RaySampler             # noqa: E402
import _databacks.lustrous_renders_v1 as ref_dk                              # noqa: E402

from oracle import renderer_oracle as orc                                    # noqa: E402
from tests.golden.cases import RENDER_CASES, POINT_CASES, VOLUME_CASES, build_case_inputs  # noqa: E402


class _InjectRand:
    """Feed pre-drawn uniforms to the reference's torch.rand_like / torch.rand calls, in call order."""

    def __init__(self, queue):
        self.queue = list(queue)

    def __enter__(self):
        self._rl, self._r = torch.rand_like, torch.rand

        def rand_like(t, *a, **k):
            u = self.queue.pop(0)
            assert u.shape == t.shape, (u.shape, t.shape)
            return u.clone()

        def rand(*size, **k):
            u = self.queue.pop(0)
            shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
            assert tuple(u.shape) == shape, (u.shape, shape)
            return u.clone()

        torch.rand_like, torch.rand = rand_like, rand
        return self

    def __exit__(self, *exc):
        torch.rand_like, torch.rand = self._rl, self._r
        assert not self.queue or exc[0] is not None, 'unused injected noise'


def make_ref_decoder(dec):
    d = ref_tp.OSGDecoder(dec['w1'].shape[1], {'decoder_lr_mul': dec['lr_mul'], 'decoder_output_dim': dec['w2'].shape[0] - 1})
    with torch.no_grad():
        d.net[0].weight.copy_(dec['w1']); d.net[0].bias.copy_(dec['b1'])
        d.net[2].weight.copy_(dec['w2']); d.net[2].bias.copy_(dec['b2'])
    d.set_force_sigmoid(bool(dec['force_sigmoid']))
    return d.eval().requires_grad_(False)


def ref_rays(case, c2w, K):
    R = case['R']
    if case.get('ortho'):
        ros, rds = [], []
        for (elev, azim, dist, _fov) in case['cameras']:
            r = ref_dk.get_rays_ortho(elev, azim, dist, case['opts']['box_warp'], R)
            ros.append(r['ray_origins'].float().reshape(1, 3, R * R).permute(0, 2, 1))
            rds.append(r['ray_directions'].float().reshape(1, 3, R * R).permute(0, 2, 1))
        return torch.cat(ros).contiguous(), torch.cat(rds).contiguous()
    return RaySampler()(c2w, K, R)


def run_render_case(name, case):
    planes, dec, c2w, K, u_c, u_f, opts = build_case_inputs(case)
    ro, rd = ref_rays(case, c2w, K)
    renderer = ImportanceRenderer(use_triplane=case.get('use_triplane', True))
    decoder = make_ref_decoder(dec)
    noise = [u_c] + ([u_f] if opts['depth_resolution_importance'] > 0 else [])
    with torch.no_grad(), _InjectRand(noise):
        rgb, depth, wsum, xyz = renderer(planes.contiguous(), decoder, ro, rd, opts,
                                         triplane_crop=case.get('triplane_crop'),
                                         cull_clouds=case.get('cull_clouds'),
                                         binarize_clouds=case.get('binarize_clouds'))
    chk = float(planes.double().sum() + u_c.double().sum() + dec['w1'].double().sum())
    st = int(case.get('store_stride', 1))          # big cases keep every st-th ray of the reference output
    if st > 1:
        rgb, depth, wsum, xyz = (t[:, ::st].contiguous() for t in (rgb, depth, wsum, xyz))
    np.savez_compressed(os.path.join(HERE, f'render_{name}.npz'),
                        rgb=rgb.numpy(), depth=depth.numpy(), wsum=wsum.numpy(), xyz=xyz.numpy(),
                        ro=ro.numpy() if ro.numel() <= 3 * 4096 * 2 else ro[:, ::97].numpy(),
                        rd=rd.numpy() if rd.numel() <= 3 * 4096 * 2 else rd[:, ::97].numpy(),
                        c2w=c2w.numpy(), K=K.numpy(),
                        input_checksum=np.float64(chk), case=json.dumps(case))
    print(f'render_{name}: rgb {tuple(rgb.shape)} mean {rgb.mean():+.5f} wsum mean {wsum.mean():.5f} '
          f'depth [{depth.min():.4f},{depth.max():.4f}]')


def run_point_case(name, case):
    planes, dec, _, _, _, _, opts = build_case_inputs(case)
    rng = np.random.default_rng(case['seed'] + 1000)
    pts = torch.from_numpy(rng.uniform(-0.45, 0.45, size=(planes.shape[0], case['K'], 3)).astype(np.float32))
    renderer = ImportanceRenderer(use_triplane=case.get('use_triplane', True))
    decoder = make_ref_decoder(dec)
    with torch.no_grad():
        out = renderer.run_model(planes.contiguous(), decoder, pts, torch.zeros_like(pts), opts)
    np.savez_compressed(os.path.join(HERE, f'points_{name}.npz'),
                        rgb=out['rgb'].numpy(), sigma=out['sigma'].numpy(), pts=pts.numpy(), case=json.dumps(case))
    print(f'points_{name}: rgb mean {out["rgb"].mean():+.5f} sigma mean {out["sigma"].mean():+.5f}')


def _reference_volume_functions():
    """sigma2density / create_samples / get_eg3d_volume exactly as written in the reference, without importing the module
    (its top-level imports need pyvista, dnnlib, legacy, ...): the three function definitions are sliced out of
    _util/eg3d_metrics3d.py by their AST line ranges and executed in a namespace holding what they reference."""
    import ast
    from training.volumetric_rendering.renderer import triplane_crop_mask, cull_clouds_mask
    path = REF + '/_util/eg3d_metrics3d.py'
    src = open(path).read()
    lines = src.splitlines(keepends=True)
    want = {'sigma2density', 'create_samples', 'get_eg3d_volume'}
    chunks = [''.join(lines[n.lineno - 1:n.end_lineno]) for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in want]
    assert len(chunks) == 3

    class Dict(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__
    ns = dict(torch=torch, np=np, nn=torch.nn, Dict=Dict, device=torch.device('cpu'), triplane_crop_mask=triplane_crop_mask,
              cull_clouds_mask=cull_clouds_mask)
    exec(compile(''.join(chunks), path, 'exec'), ns)
    return ns


def run_volume_case(name, case):
    """get_eg3d_volume (eg3d_metrics3d.py:94-183) on a stand-in G: f() resolves ws, sample_mixed() is the reference
    ImportanceRenderer.run_model on fixed tri-planes (what TriPlaneGenerator.sample_mixed does after its backbone,
    triplane.py:283-298)."""
    ns = _reference_volume_functions()
    planes, dec, _, _, _, _, opts = build_case_inputs(case)
    renderer = ImportanceRenderer(use_triplane=case.get('use_triplane', True))
    decoder = make_ref_decoder(dec)

    class G:
        rendering_kwargs = opts

        @staticmethod
        def f(xin):
            xin['ws'] = torch.zeros(1, 14, 512)

        @staticmethod
        def sample_mixed(coords, dirs, ws, cond, **kw):
            return renderer.run_model(planes.contiguous(), decoder, coords, dirs, opts)
    xin = {'cond': None}
    if 'triplane_crop' in case:
        xin['triplane_crop'] = case['triplane_crop']
    if 'cull_clouds' in case:
        xin['cull_clouds'] = case['cull_clouds']
    vol = ns['get_eg3d_volume'](G, xin, resolution=case['res'], max_batch=777)
    np.savez_compressed(os.path.join(HERE, f'volume_{name}.npz'), coordinates=vol['coordinates'].numpy(), sigmas=vol['sigmas'].numpy(),
                        rgbs=vol['rgbs'].numpy(), densities=vol['densities'].numpy(), case=json.dumps(case))
    d = vol['densities']
    print(f'volume_{name}: sigmas {tuple(vol["sigmas"].shape)} mean {vol["sigmas"].mean():+.4f} masked {(d == -1e3).float().mean():.3f}')


def run_camera_table():
    """camera_params_to_matrix + get_rays_ortho + cam60/spin12 table (lustrous_renders_v1.py:14-104)."""
    cams = [(0.0, a, 1.0, 30.0) for a in range(-180, 180, 30)] + [(10.0, 30.0, 1.0, 30.0), (60.0, -45.0, 1.2, 45.0), (-20.0, 100.0, 0.9, 12.0)]
    c2w, K = [], []
    for e, a, d, f in cams:
        m = ref_dk.camera_params_to_matrix('eg3d_lustrousB', elev=e, azim=a, dist=d, fov=f)
        c2w.append(m['matrix_extrinsic'].numpy()); K.append(m['matrix_intrinsic'].numpy())
    ortho = [(0.0, 0.0, 1.0), (0.0, 90.0, 1.0), (20.0, -135.0, 1.0)]
    oro, ord_ = [], []
    for e, a, d in ortho:
        r = ref_dk.get_rays_ortho(e, a, d, 0.7, 8)
        oro.append(r['ray_origins'].float().numpy()); ord_.append(r['ray_directions'].float().numpy())
    spin = ref_dk.cam60[ref_dk.camsubs['spin12']].numpy()
    np.savez_compressed(os.path.join(HERE, 'cameras.npz'), cams=np.array(cams, np.float64), c2w=np.stack(c2w), K=np.stack(K),
                        ortho=np.array(ortho, np.float64), ortho_ro=np.concatenate(oro), ortho_rd=np.concatenate(ord_),
                        spin12=spin)
    print('cameras: ok', spin[:2])


if __name__ == '__main__':
    only = set(sys.argv[1:])
    torch.set_num_threads(os.cpu_count())
    for name, case in RENDER_CASES.items():
        if not only or name in only:
            run_render_case(name, case)
    for name, case in POINT_CASES.items():
        if not only or name in only:
            run_point_case(name, case)
    for name, case in VOLUME_CASES.items():
        if not only or name in only:
            run_volume_case(name, case)
    if not only or 'cameras' in only:
        run_camera_table()
