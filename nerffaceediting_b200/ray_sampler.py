"""Drop-in RaySampler (reference: training/volumetric_rendering/ray_sampler.py:24-63): one kernel
instead of ~15 tiny ATen launches.  Expects OpenCV-convention cam2world matrices, as the reference."""
import torch

from . import ops


class RaySampler(torch.nn.Module):
    def __init__(self):
        super().__init__()
        # attributes the reference initialises (ray_sampler.py:21); kept for pickle compatibility
        self.ray_origins_h, self.ray_directions, self.depths, self.image_coords, self.rendering_options = None, None, None, None, None

    def forward(self, cam2world_matrix, intrinsics, resolution):
        """cam2world_matrix [N,4,4], intrinsics [N,3,3] (normalised), resolution int ->
        ray_origins [N,res*res,3], ray_dirs [N,res*res,3]; ray m = row*res + col."""
        return ops.generate_rays(cam2world_matrix, intrinsics, resolution)
