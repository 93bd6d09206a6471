"""Shadow of `torch_utils.ops`: bias_act, upfirdn2d and conv2d_resample come from this directory, everything else (fma,
conv2d_gradfix, grid_sample_gradfix, filtered_lrelu, ...) from the reference's `torch_utils/ops/`, appended to the search path."""
import importlib.util
import os
import sys

import torch_utils

if torch_utils.REFERENCE_TORCH_UTILS is not None:
    _ops = os.path.join(torch_utils.REFERENCE_TORCH_UTILS, "ops")
    if os.path.isdir(_ops) and _ops not in __path__:
        __path__.append(_ops)


def load_reference(name):
    """The reference's own module of the same name (for what the shadow delegates: CPU tensors, training, configurations outside the
    generator's inference path), loaded from its file under a private name."""
    key = f"torch_utils.ops._reference_{name}"
    if key in sys.modules:
        return sys.modules[key]
    if torch_utils.REFERENCE_TORCH_UTILS is None:
        raise ImportError(f"shadow torch_utils.ops.{name}: the reference checkout was not found on sys.path / $NFE_REFERENCE")
    path = os.path.join(torch_utils.REFERENCE_TORCH_UTILS, "ops", name + ".py")
    spec = importlib.util.spec_from_file_location(key, path, submodule_search_locations=None)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "torch_utils.ops"          # its relative imports (`from .. import misc`, `from . import conv2d_gradfix`) resolve
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod
