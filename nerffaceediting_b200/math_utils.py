"""Drop-in math_utils (reference: training/volumetric_rendering/math_utils.py).  get_ray_limits_box is a
CUDA kernel; the remaining helpers are one-line tensor expressions kept for API parity."""
import torch

from . import ops


def transform_vectors(matrix: torch.Tensor, vectors4: torch.Tensor) -> torch.Tensor:
    """Left-multiplies MxM @ NxM. Returns NxM (math_utils.py:26-31)."""
    return torch.matmul(vectors4, matrix.T)


def normalize_vecs(vectors: torch.Tensor) -> torch.Tensor:
    """vectors / ||vectors|| (math_utils.py:33-37)."""
    return vectors / (torch.norm(vectors, dim=-1, keepdim=True))


def torch_dot(x: torch.Tensor, y: torch.Tensor):
    return (x * y).sum(-1)


def get_ray_limits_box(rays_o: torch.Tensor, rays_d: torch.Tensor, box_side_length):
    """Ray / [-L/2,L/2]^3 slab test -> (tmin, tmax) shaped [..., 1]; rays that miss get (-1, -2)
    (math_utils.py:46-98)."""
    return ops.ray_limits_box(rays_o.detach(), rays_d.detach(), box_side_length)


def linspace(start: torch.Tensor, stop: torch.Tensor, num: int):
    """[num, *start.shape] evenly spaced from start to stop inclusive (math_utils.py:101-118)."""
    steps = torch.arange(num, dtype=torch.float32, device=start.device) / (num - 1)
    for _ in range(start.ndim):
        steps = steps.unsqueeze(-1)
    return start[None] + steps * (stop - start)[None]
