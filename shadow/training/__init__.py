"""Shadow of the reference's `training` package: module-path drop-in for the renderer.

The reference pickles its renderer, ray marcher and ray sampler BY MODULE PATH
(`training.volumetric_rendering.{renderer,ray_marcher,ray_sampler}`; they are not persistent classes,
SURVEY.md §7.9/§8b) and `training/triplane.py:14-15` imports them by name.  Putting this directory AHEAD of the
reference checkout on sys.path therefore makes every `TriPlaneGenerator` — freshly constructed or unpickled
from a checkpoint — render through nerffaceediting_b200, while every other `training.*` module
(triplane, networks_stylegan2, superresolution, loss, ...) still resolves to the reference, whose
`training/` directory is appended to this package's search path below.

    PYTHONPATH=/path/to/nfe-b200/shadow:/path/to/NeRFFaceEditing python gen_samples.py ...

The reference checkout is found through $NFE_REFERENCE or by scanning sys.path.
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
        cand = os.path.join(root, "training")
        if os.path.abspath(cand) != _here and os.path.isfile(os.path.join(cand, "triplane.py")):
            yield root


for _root in _reference_roots():
    _t = os.path.join(_root, "training")
    if os.path.isdir(_t) and _t not in __path__:
        __path__.append(_t)
        break
