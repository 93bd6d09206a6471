"""Module-path shadow of the reference's training/volumetric_rendering/ray_sampler.py."""
from nerffaceediting_b200 import ray_sampler as _impl


class RaySampler(_impl.RaySampler):
    pass
