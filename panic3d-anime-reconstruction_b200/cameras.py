"""Host-side camera helpers with the reference's conventions (/root/reference/_databacks/
lustrous_renders_v1.py:14-104): the 'eg3d_lustrousB' camera matrices, the cam60/spin12 view table and
orthographic rays.  Pure host math (25 floats per view); rays themselves are generated on the GPU."""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib


def _euler_xyz(ax_deg, ay_deg, az_deg):
    """Extrinsic x, then y, then z rotation = scipy Rotation.from_euler('xyz', ..., degrees=True)."""
    ax, ay, az = (math.radians(float(a)) for a in (ax_deg, ay_deg, az_deg))
    rx = np.array([[1, 0, 0], [0, math.cos(ax), -math.sin(ax)], [0, math.sin(ax), math.cos(ax)]])
    ry = np.array([[math.cos(ay), 0, math.sin(ay)], [0, 1, 0], [-math.sin(ay), 0, math.cos(ay)]])
    rz = np.array([[math.cos(az), -math.sin(az), 0], [math.sin(az), math.cos(az), 0], [0, 0, 1]])
    return rz @ ry @ rx


# lustrous_renders_v1.py:14-30: 5 elevations x 12 azimuths; spin12 = the elev-0 ring starting at azim 0
cam60 = torch.tensor(np.stack(np.meshgrid(np.linspace(60, -20, 5), np.linspace(-180, 150, 12))).T.reshape(60, -1)).float()
camsubs = {'all': list(range(60)), 'front1': [42], 'spin12': [*range(42, 48), *range(36, 42)]}


def camera_params_to_matrix(mode='eg3d_lustrousB', *, elev, azim, dist, fov):
    """-> dict(matrix_intrinsic (3,3), matrix_extrinsic (4,4), camera_label (25,)).  lustrous_renders_v1.py:33-75."""
    assert mode == 'eg3d_lustrousB', 'mode not understood'
    elev, azim, dist, fov = (float(v) for v in (elev, azim, dist, fov))
    focal = 0.5 / np.tan((fov / 2) * np.pi / 180)
    intr = np.asarray([[focal, 0, 0.5], [0, focal, 0.5], [0, 0, 1]], dtype=np.float32)
    R = np.eye(4)
    R[:3, :3] = _euler_xyz(elev, azim, 0).T
    R[[0, 2]] *= -1
    R[2, -1] = -dist
    extr = np.diag([-1.0, 1, -1, 1]) @ np.linalg.inv(R) @ np.diag([1.0, -1, -1, 1])
    intr, extr = torch.tensor(intr).float(), torch.tensor(extr).float()
    return dict(matrix_intrinsic=intr, matrix_extrinsic=extr, camera_label=torch.cat([extr.flatten(), intr.flatten()]))


def get_rays_ortho(elev, azim, dist, boxwarp, resolution, device=None):
    """-> {'ray_origins','ray_directions'} each (1,3,R,R) on `device` (CUDA).  lustrous_renders_v1.py:78-104."""
    if device is None or torch.device(device).type != 'cuda':
        raise RuntimeError('panic3d_b200.get_rays_ortho generates rays on the GPU: pass a CUDA device')
    dev = torch.device(device)
    R = int(resolution)
    rot = torch.tensor(_euler_xyz(-float(elev), float(azim), 0.0), dtype=torch.float32, device=dev).contiguous()
    d = torch.tensor([float(dist)], dtype=torch.float32, device=dev)
    ro = torch.empty((1, R * R, 3), device=dev, dtype=torch.float32)
    rd = torch.empty((1, R * R, 3), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().p3d_raygen_ortho(rot.data_ptr(), d.data_ptr(), 1, R, float(boxwarp), ro.data_ptr(),
                                               rd.data_ptr(), _lib.stream_ptr(dev)))
    to_img = lambda t: t.reshape(1, R, R, 3).permute(0, 3, 1, 2).contiguous()
    return {'ray_origins': to_img(ro), 'ray_directions': to_img(rd)}
