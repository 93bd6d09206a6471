"""Shadow of the reference's `torch_utils` package: module-path drop-in for the backbone / super-resolution plugins.

The reference's persistent classes (SynthesisLayer, ToRGBLayer, ... pickled WITH their source) import their plugins by module path —
`from torch_utils.ops import bias_act, upfirdn2d, conv2d_resample, fma` (training/networks_stylegan2.py:17-21) — so putting this
directory AHEAD of the reference checkout on sys.path makes every generator, freshly constructed or unpickled from a checkpoint, run
its inference convolutions, filters and activations through nerffaceediting_b200 (no JIT build of the reference's CUDA plugins),
while every other `torch_utils.*` module (misc, persistence, custom_ops, training_stats, ops.fma, ops.conv2d_gradfix, ...) still
resolves to the reference, whose `torch_utils/` directory is appended to this package's search path below.
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_repo = os.path.dirname(os.path.dirname(_here))
if _repo not in sys.path:
    sys.path.append(_repo)          # makes `nerffaceediting_b200` importable


def _reference_roots():
    env = os.environ.get("NFE_REFERENCE")
    if env:
        yield env
    for p in list(sys.path):
        root = os.path.abspath(p or ".")
        cand = os.path.join(root, "torch_utils")
        if os.path.abspath(cand) != _here and os.path.isfile(os.path.join(cand, "persistence.py")):
            yield root


REFERENCE_TORCH_UTILS = None
for _root in _reference_roots():
    _t = os.path.join(_root, "torch_utils")
    if os.path.isdir(_t) and _t not in __path__:
        __path__.append(_t)
        REFERENCE_TORCH_UTILS = _t
        break
