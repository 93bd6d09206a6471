"""Drop-in MipRayMarcher2 / SegMipRayMarcher2 (reference: training/volumetric_rendering/ray_marcher.py).

Same constructor, same forward/run_forward signatures and return shapes; the ~20 elementwise/scan
launches of the reference are one warp-per-ray CUDA kernel (csrc/nfe_march.cu).  Objects carry no
state, so instances unpickled from reference checkpoints (which skip __init__) work unchanged.
"""
import torch

from . import ops


def _check_clamp_mode(rendering_options):
    # ray_marcher.py:32-35,75-78
    assert rendering_options['clamp_mode'] == 'softplus', "MipRayMarcher only supports `clamp_mode`=`softplus`!"


class MipRayMarcher2(torch.nn.Module):
    def __init__(self):
        super().__init__()

    def run_forward(self, colors, densities, depths, rendering_options):
        """colors [N,R,S,C], densities/depths [N,R,S,1] -> (rgb [N,R,C], depth [N,R,1], weights [N,R,S-1,1])
        (ray_marcher.py:25-57)."""
        _check_clamp_mode(rendering_options)
        ops._no_grad_needed(colors, densities, depths)
        rgb, _, depth, weights = ops.composite(colors, densities, depths, None, rendering_options.get('white_back', False))
        return rgb, depth, weights

    def forward(self, colors, densities, depths, rendering_options):
        return self.run_forward(colors, densities, depths, rendering_options)


class SegMipRayMarcher2(torch.nn.Module):
    def __init__(self):
        super().__init__()

    def run_forward(self, colors, segs, densities, depths, rendering_options):
        """As MipRayMarcher2 plus segs [N,R,S,Cs] -> seg [N,R,Cs], composited WITHOUT the [-1,1] rescale
        (ray_marcher.py:68-101)."""
        _check_clamp_mode(rendering_options)
        ops._no_grad_needed(colors, segs, densities, depths)
        rgb, seg, depth, weights = ops.composite(colors, densities, depths, segs, rendering_options.get('white_back', False))
        return rgb, seg, depth, weights

    def forward(self, colors, segs, densities, depths, rendering_options):
        return self.run_forward(colors, segs, densities, depths, rendering_options)
