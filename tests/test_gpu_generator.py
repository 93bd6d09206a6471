"""The whole generator — BASELINE configs[1]: mapping + StyleGAN2 tri-plane backbone + decoders + renderer + super-resolution
(training/triplane.py:18-165) — composed of this package's modules only (`nerffaceediting_b200.triplane.TriPlaneGenerator`), against the
UNMODIFIED reference generator run on CPU with deterministic sampling (tests/golden/make_golden_generator.py -> generator.npz).
Tolerance: 1e-3 of the tensor's largest magnitude through the whole chain in fp32 mode (backbone: 13 split-bf16 convolutions, the render at
1e-4, super-resolution: 5 more), 1e-4 for what comes before the renderer (ws, plane statistics)."""
import os
import sys

import numpy as np
import pytest
import torch

from _util import golden, rel_err

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import conv_cases as cases  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-3


def run(which):
    from nerffaceediting_b200 import ops, triplane
    G = cases.make_generator(triplane.TriPlaneGenerator, which).cuda()
    z, cam, pts = (t.cuda() for t in cases.generator_inputs(which))
    ops.path_counts(reset=True)
    with torch.no_grad():
        ws = G.mapping(z, cam, truncation_psi=0.7, truncation_cutoff=6)
        out = G.synthesis(ws, cam, noise_mode='const')
    return G, ws, cam, pts, out


def close(a, ref, tol, what):
    a = a.float().cpu().numpy()
    assert a.shape == ref.shape and np.isfinite(a).all(), what
    e = rel_err(a, ref)
    assert e < tol, (what, e)


def test_generator_128_matches_the_reference():
    from nerffaceediting_b200 import ops
    g = golden("generator")
    G, ws, cam, pts, out = run('g128')
    assert ops.path_counts().get("render:single-gather", 0) == 1          # statistics + staging + provenance in one pass, one plane set gathered
    close(ws, g["g128.ws"], 1e-4, "ws")
    close(out['plane_mean'], g["g128.plane_mean"], TOL, "plane_mean")
    close(out['plane_var'], g["g128.plane_var"], TOL, "plane_var")
    for k in ('image', 'image_seg', 'image_raw', 'image_depth'):
        close(out[k], g[f"g128.{k}"], TOL, k)
    with torch.no_grad():
        swap = G.synthesis(ws, cam, noise_mode='const', planes_mean=1, planes_var=1)          # appearance swap (triplane.py:98-101)
        sm = G.sample_mixed(pts, torch.zeros_like(pts), ws, noise_mode='const')
    close(swap['image'], g["g128.swap.image"], TOL, "swap image")
    close(swap['image_seg'], g["g128.swap.image_seg"], TOL, "swap seg")
    for k in ('rgb', 'sigma', 'seg'):
        close(sm[k], g[f"g128.sample_mixed.{k}"], TOL, "sample_mixed " + k)
    # forward() = mapping + synthesis
    z = cases.generator_inputs('g128')[0].cuda()
    with torch.no_grad():
        out2 = G(z, cam, truncation_psi=0.7, truncation_cutoff=6, noise_mode='const')
    assert torch.equal(out2['image'], out['image'])


def test_generator_512_matches_the_reference():
    g = golden("generator")
    G, ws, cam, pts, out = run('g512')
    assert out['image'].shape == (1, 3, 512, 512)
    close(out['image'][:, :, ::4, ::4], g["g512.image_s4"], TOL, "image (every 4th pixel)")
    close(out['image'][:, :, 253:259, :], g["g512.image_rows"], TOL, "image rows")
    for k in ('image_seg', 'image_raw', 'image_depth'):
        close(out[k], g[f"g512.{k}"], TOL, k)


def test_generator_state_dict_is_the_reference_s():
    from nerffaceediting_b200 import triplane
    g = golden("generator")
    for which in cases.GENERATORS:
        G = cases.make_generator(triplane.TriPlaneGenerator, which)
        keys = sorted(f"{k}:{'x'.join(map(str, v.shape))}" for k, v in G.state_dict().items())
        assert keys == sorted(g[f"keys.{which}"].tolist())


def test_generator_replays_as_one_cuda_graph():
    """Every kernel of the chain only enqueues work (no allocation inside the library, no synchronisation, no host read-back), so the
    whole generator — ~250 launches — captures as one CUDA graph; the replay is bit-identical and follows in-place input updates."""
    from nerffaceediting_b200 import graphs, triplane
    G = cases.make_generator(triplane.TriPlaneGenerator, 'g128').cuda()
    z, cam, _ = (t.cuda() for t in cases.generator_inputs('g128'))
    with torch.no_grad():
        ref = {k: v.clone() for k, v in G(z, cam, noise_mode='const').items() if k.startswith('image')}
    step = graphs.capture(lambda: G(z, cam, noise_mode='const'))
    assert step.kernels > 50
    out = step()
    for k, v in ref.items():
        assert torch.equal(out[k], v), k
    z2 = torch.roll(z, 1, 0)
    with torch.no_grad():
        ref2 = G(z2, cam, noise_mode='const')['image'].clone()
    z.copy_(z2)                                     # static input updated in place, then replayed
    assert torch.equal(step()['image'], ref2)
