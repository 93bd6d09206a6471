"""Module-path shadow of the reference's training/volumetric_rendering/ray_marcher.py."""
from nerffaceediting_b200 import ray_marcher as _impl


class MipRayMarcher2(_impl.MipRayMarcher2):
    pass


class SegMipRayMarcher2(_impl.SegMipRayMarcher2):
    pass
