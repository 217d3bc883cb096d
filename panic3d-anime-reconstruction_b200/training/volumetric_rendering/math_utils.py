"""Host mirror of ``training/volumetric_rendering/math_utils.py`` (reference math_utils.py:25-118): the same five functions.

On the renderer's path ``get_ray_limits_box`` and ``linspace`` (ray limits ``'auto'``, reference renderer.py:165-171,308-314) are
fused into the CUDA renderer (``k_ray_limits`` + ``dev::coarse_depth``); the versions below serve stand-alone callers of the module and
are plain torch ops on whatever device the arguments live on - they are not part of the hot path.  Outputs equal the reference's bit
for bit (``tests/test_math_utils_cpu.py`` against ``tests/golden/math_utils.npz``, written from the imported reference), including its
conventions: ``(-1, -2)`` for rays that miss the box, NaN / inf propagation for axis-parallel rays."""
import torch


def transform_vectors(matrix: torch.Tensor, vectors4: torch.Tensor) -> torch.Tensor:
    """M x M times N x M -> N x M (math_utils.py:25-30)."""
    return torch.matmul(vectors4, matrix.T)


def normalize_vecs(vectors: torch.Tensor) -> torch.Tensor:
    """math_utils.py:33-37."""
    return vectors / torch.norm(vectors, dim=-1, keepdim=True)


def torch_dot(x: torch.Tensor, y: torch.Tensor):
    """math_utils.py:39-43."""
    return (x * y).sum(-1)


def get_ray_limits_box(rays_o: torch.Tensor, rays_d: torch.Tensor, box_side_length):
    """Slab test of rays against the axis-aligned cube of side ``box_side_length`` centred at the origin (math_utils.py:46-98).
    -> (t_near, t_far), each ``rays_o.shape[:-1] + (1,)``; rays that miss get (-1, -2)."""
    shape = rays_o.shape
    o, d = rays_o.detach().reshape(-1, 3), rays_d.detach().reshape(-1, 3)
    half = box_side_length / 2
    lo, hi = torch.full_like(o, -1 * half), torch.full_like(o, 1 * half)
    inv = 1 / d                                                     # +-inf on axis-parallel rays, as in the reference
    backwards = inv < 0
    near = (torch.where(backwards, hi, lo) - o) * inv               # per-axis entry / exit distances
    far = (torch.where(backwards, lo, hi) - o) * inv
    t0, t1 = near[:, 0], far[:, 0]
    hit = torch.ones(o.shape[0], dtype=torch.bool, device=o.device)
    for axis in (1, 2):                                             # x slab against y, then the running interval against z
        hit &= ~torch.logical_or(t0 > far[:, axis], near[:, axis] > t1)
        t0 = torch.maximum(t0, near[:, axis])
        t1 = torch.minimum(t1, far[:, axis])
    t0 = torch.where(hit, t0, torch.full_like(t0, -1))
    t1 = torch.where(hit, t1, torch.full_like(t1, -2))
    return t0.reshape(*shape[:-1], 1), t1.reshape(*shape[:-1], 1)


def linspace(start: torch.Tensor, stop: torch.Tensor, num: int):
    """[num, *start.shape] values evenly spaced from ``start`` to ``stop`` inclusive (math_utils.py:101-118)."""
    steps = (torch.arange(num, dtype=torch.float32, device=start.device) / (num - 1)).reshape([num] + [1] * start.ndim)
    return start[None] + steps * (stop - start)[None]
