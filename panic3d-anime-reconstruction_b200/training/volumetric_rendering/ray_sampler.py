"""Drop-in for ``training/volumetric_rendering/ray_sampler.py`` (reference ray_sampler.py:18-63)."""
import torch

from ... import _lib


class RaySampler(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.ray_origins_h, self.ray_directions, self.depths, self.image_coords, self.rendering_options = None, None, None, None, None

    def forward(self, cam2world_matrix, intrinsics, resolution):
        """cam2world (N,4,4), intrinsics (N,3,3), resolution int -> ray_origins (N,R*R,3), ray_dirs (N,R*R,3)."""
        if not cam2world_matrix.is_cuda:
            raise RuntimeError('panic3d_b200.RaySampler has no CPU path: tensors must be on a CUDA device')
        dev = cam2world_matrix.device
        N, R = cam2world_matrix.shape[0], int(resolution)
        c2w = cam2world_matrix.detach().float().contiguous()
        K = intrinsics.detach().to(dev).float().contiguous()
        ro = torch.empty((N, R * R, 3), device=dev, dtype=torch.float32)
        rd = torch.empty((N, R * R, 3), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().p3d_raygen_pinhole(c2w.data_ptr(), K.data_ptr(), N, R, ro.data_ptr(), rd.data_ptr(),
                                                     _lib.stream_ptr(dev)))
        return ro, rd
