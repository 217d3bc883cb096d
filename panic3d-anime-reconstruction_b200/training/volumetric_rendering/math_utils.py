"""Drop-in for ``training/volumetric_rendering/math_utils.py`` (reference math_utils.py:25-118).

Only the two helpers the renderer path uses exist in the reference call graph
(``get_ray_limits_box`` and ``linspace``, both for ray limits 'auto'); they are fused into the CUDA
renderer (k_ray_limits).  The tiny host-side vector helpers are kept for API compatibility."""
import torch


def transform_vectors(matrix: torch.Tensor, vectors4: torch.Tensor) -> torch.Tensor:
    return torch.matmul(vectors4, matrix.T)


def normalize_vecs(vectors: torch.Tensor) -> torch.Tensor:
    return vectors / torch.norm(vectors, dim=-1, keepdim=True)


def torch_dot(x: torch.Tensor, y: torch.Tensor):
    return (x * y).sum(-1)
