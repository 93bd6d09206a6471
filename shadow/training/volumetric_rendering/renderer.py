"""Module-path shadow of the reference's training/volumetric_rendering/renderer.py (see shadow/training/__init__.py).

The classes are thin subclasses so that their `__module__` is this module's path: a generator built or
loaded through the shadow pickles exactly like a reference one and stays loadable by a stock checkout.
"""
from nerffaceediting_b200 import renderer as _impl
from nerffaceediting_b200.renderer import generate_planes, project_onto_planes, sample_from_3dgrid, sample_from_planes  # noqa: F401

from training.volumetric_rendering import math_utils  # noqa: F401
from training.volumetric_rendering.ray_marcher import MipRayMarcher2, SegMipRayMarcher2


class ImportanceRenderer(_impl.ImportanceRenderer):
    def __init__(self):
        super().__init__()
        self.ray_marcher = MipRayMarcher2()


class DisentangledImportanceRenderer(_impl.DisentangledImportanceRenderer):
    def __init__(self):
        super().__init__()
        self.ray_marcher = SegMipRayMarcher2()
