"""Module-path shadow of the reference's training/volumetric_rendering/math_utils.py."""
from nerffaceediting_b200.math_utils import get_ray_limits_box, linspace, normalize_vecs, torch_dot, transform_vectors  # noqa: F401
