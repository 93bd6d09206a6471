"""Fixture for the WHOLE generator (training/triplane.py:18-165 = BASELINE configs[1]: mapping + StyleGAN2 tri-plane backbone +
decoders + renderer + super-resolution), produced by the UNMODIFIED reference TriPlaneGenerator on CPU (fp32), with the sampling made
deterministic exactly as tests/golden/make_golden.py does (zero stratified jitter, det=True inverse CDF: BASELINE.md §3):

    python tests/golden/make_golden_generator.py        -> tests/golden/generator.npz

    g128  2 poses, SuperresolutionHybrid2X, 128^2 output: image / image_seg / image_raw / image_depth / plane statistics,
          the appearance-swap call (planes_mean = planes_var = 1, triplane.py:98-101) and sample_mixed at 600 points per item
    g512  1 pose, SuperresolutionHybrid8XDC, 512^2 output (stored at every 4th pixel + 6 full rows)
The backbone is narrow (channel_base 4096, channel_max 32: 32 channels up to 128^2, 16 at 256^2) so that the CPU run takes seconds; parameters and inputs are regenerated
from synth_inputs on both sides (tests/golden/conv_cases.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("NFE_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REFERENCE)
sys.path.insert(0, HERE)
import conv_cases as cases  # noqa: E402
from make_golden import deterministic  # noqa: E402
from training.triplane import TriPlaneGenerator  # noqa: E402
from training.volumetric_rendering.renderer import DisentangledImportanceRenderer  # noqa: E402


def N(t):
    return t.detach().cpu().numpy().astype(np.float32)


def main():
    arrays = {}
    with torch.no_grad():
        for which in cases.GENERATORS:
            G = cases.make_generator(TriPlaneGenerator, which)
            G.renderer = deterministic(DisentangledImportanceRenderer)
            G.rendering_kwargs = {k: v for k, v in G.rendering_kwargs.items() if not k.startswith('nfe_')}
            z, cam, pts = cases.generator_inputs(which)
            ws = G.mapping(z, cam, truncation_psi=0.7, truncation_cutoff=6)
            out = G.synthesis(ws, cam, noise_mode='const')
            arrays[f"{which}.ws"] = N(ws)
            if which == 'g512':
                arrays[f"{which}.image_s4"] = N(out['image'][:, :, ::4, ::4])
                arrays[f"{which}.image_rows"] = N(out['image'][:, :, 253:259, :])
            else:
                arrays[f"{which}.image"] = N(out['image'])
                swap = G.synthesis(ws, cam, noise_mode='const', planes_mean=1, planes_var=1)
                arrays[f"{which}.swap.image"] = N(swap['image'])
                arrays[f"{which}.swap.image_seg"] = N(swap['image_seg'])
                sm = G.sample_mixed(pts, torch.zeros_like(pts), ws, noise_mode='const')
                for k in ('rgb', 'sigma', 'seg'):
                    arrays[f"{which}.sample_mixed.{k}"] = N(sm[k])
            for k in ('image_seg', 'image_raw', 'image_depth', 'plane_mean', 'plane_var'):
                arrays[f"{which}.{k}"] = N(out[k])
            arrays[f"keys.{which}"] = np.array([f"{k}:{'x'.join(map(str, v.shape))}" for k, v in G.state_dict().items()])
    np.savez_compressed(os.path.join(HERE, "generator.npz"), **arrays)
    print("generator.npz", len(arrays), "arrays", {k: v.shape for k, v in arrays.items() if not k.startswith('keys')})


if __name__ == "__main__":
    main()
