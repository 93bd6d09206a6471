"""Low-resolution layers of the backbone (512 -> 512 at 4^2 .. 32^2, fp32, batch 8), CUDA-event timed in a stream (warm L2, unlike an
ncu launch list): the layers split K addresses.  usage: [NFE_CONV_KSPLIT=1] python profiles/bench_small_layers.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth_inputs as synth  # noqa: E402
from nerffaceediting_b200 import networks as net  # noqa: E402
from profiles.bench_conv import timed  # noqa: E402

print("NFE_CONV_KSPLIT =", os.environ.get("NFE_CONV_KSPLIT", "(default 8)"))
for res, up in [(4, 1), (8, 2), (8, 1), (16, 2), (16, 1), (32, 2), (32, 1)]:
    layer = synth.fill_module(net.SynthesisLayer(512, 512, w_dim=512, resolution=res, up=up), 11).cuda().eval()
    x = torch.randn(8, 512, res // up, res // up, device="cuda").contiguous(memory_format=torch.channels_last)
    w = torch.randn(8, 512, device="cuda")
    with torch.no_grad():
        ms = timed(lambda: layer(x, w, noise_mode='const'), 5, 30)
    print(f"  512->512 at {res:3d}^2 up={up}: {ms * 1000:7.1f} us per layer call (fold + pack + GEMM [+ filter pass])")
