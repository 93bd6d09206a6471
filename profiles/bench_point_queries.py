"""Micro-benchmark of run_model at shape-extraction scale (gen_samples.py:184-222: 512^3 points in chunks of 1e6):
full query vs density-only query.  usage: python profiles/bench_point_queries.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerffaceediting_b200 import _lib, triplane  # noqa: E402
from nerffaceediting_b200.renderer import DisentangledImportanceRenderer  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
raw = torch.randn(1, 96, 256, 256, device=dev)
dec = triplane.DisentangledOSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32, 'decoder_seg_dim': 15}).to(dev)
ren = DisentangledImportanceRenderer()
m = 4_000_000
coords = (torch.rand(1, m, 3, device=dev) - 0.5)
with torch.no_grad():
    norm = triplane.normalize_plane(raw)[0].view(1, 3, 32, 256, 256)
    planes = raw.view(1, 3, 32, 256, 256)
    for prec in ("fp32", "bf16x3"):
        for only in (False, True):
            opts = {'box_warp': 1, 'nfe_precision': prec, 'nfe_sigma_only': only, 'nfe_cache_planes': True}
            for _ in range(3):
                ren.run_model(norm, planes, dec, coords, None, opts)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ren.run_model(norm, planes, dec, coords, None, opts)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"run_model {prec:7s} sigma_only={only!s:5s}  {ms:7.3f} ms per {m/1e6:.0f}M points  -> {m / ms / 1e3:8.1f} M points/s")
