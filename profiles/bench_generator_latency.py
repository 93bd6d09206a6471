"""Latency of the whole generator (BASELINE configs[1] sizes: cbase 32768, cmax 512, fp32 backbone, fp16 super-resolution, 512^2 output)
at batch 1 — the interactive viewer's case (viz/renderer.py renders one image per frame) — and batch 8, eagerly through the public
classes and replayed as ONE CUDA graph (nerffaceediting_b200.graphs.capture: ~250 launches per image otherwise).
usage: python profiles/bench_generator_latency.py [--json out.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth_inputs as synth  # noqa: E402
from nerffaceediting_b200 import graphs, triplane  # noqa: E402


def timed(fn, warmup=3, iters=20):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    rk = dict(synth.FFHQ_RENDERING_OPTIONS, superresolution_module='training.superresolution.SuperresolutionHybrid8XDC', sr_antialias=True,
              superresolution_noise_mode='none', c_gen_conditioning_zero=False, c_scale=1.0, decoder_lr_mul=1, nfe_deterministic=True)
    out = {"device": torch.cuda.get_device_name(0)}
    with torch.no_grad():
        G = triplane.TriPlaneGenerator(z_dim=512, c_dim=25, w_dim=512, img_resolution=512, img_channels=3, sr_num_fp16_res=4, mapping_kwargs=dict(num_layers=2),
                                       rendering_kwargs=rk, channel_base=32768, channel_max=512, num_fp16_res=0, conv_clamp=None,
                                       sr_kwargs=dict(channel_base=32768, channel_max=512))
        G = synth.fill_module(G, 77).cuda().eval()
        for nb in (1, 8):
            z = torch.randn(nb, 512, device="cuda")
            c2w, k = synth.camera_sweep(nb)
            cam = torch.cat([c2w.reshape(nb, 16), k.reshape(nb, 9)], dim=1).float().cuda()
            eager = timed(lambda: G(z, cam, noise_mode='const'))
            ref = G(z, cam, noise_mode='const')['image'].clone()
            step = graphs.capture(lambda: G(z, cam, noise_mode='const'))
            replay = timed(step)
            same = bool(torch.equal(step()['image'], ref))
            out[f"batch{nb}"] = {"eager_ms": round(eager, 3), "graph_ms": round(replay, 3), "images_per_s_graph": round(nb / replay * 1e3, 1),
                                 "library_kernels_in_graph": step.kernels, "replay_equals_eager": same}
            print(f"batch {nb}: eager {eager:.3f} ms, one CUDA graph {replay:.3f} ms ({nb / replay * 1e3:.0f} images/s), replay == eager: {same}", flush=True)
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
