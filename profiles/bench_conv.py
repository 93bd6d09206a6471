"""Convolution stack on one B200: the default super-resolution head (SuperresolutionHybrid8XDC,
superresolution.py:264-290) and the tri-plane backbone (SynthesisNetwork, 256 x 256 x 96) at BASELINE configs[1]'s batch, through
nerffaceediting_b200.networks, timed with CUDA events; per-layer times of the convolution kernel with the tensor-core rate they imply;
and, beside each 3x3 layer, the reference's own formulation of it — ONE grouped cuDNN convolution over per-sample weights
(networks_stylegan2.py:84-88, conv2d_gradfix -> torch.nn.functional.conv2d) — timed on the same box as the library baseline.

    python profiles/bench_conv.py [--batch 8] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth_inputs as synth  # noqa: E402
from nerffaceediting_b200 import networks as net  # noqa: E402
from nerffaceediting_b200 import _lib  # noqa: E402


def timed(fn, warmup=3, iters=10):
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


def layer_bench(n, i, o, res, up, dtype):
    """One SynthesisLayer at full size: ours (fused noise / bias / lrelu / clamp) against the reference's grouped cuDNN convolution
    (convolution only for up = 1; transposed convolution + the same filter pass for up = 2 is NOT included on the library side, so
    the comparison flatters the library)."""
    layer = synth.fill_module(net.SynthesisLayer(i, o, w_dim=512, resolution=res, up=up, conv_clamp=256), 11).cuda().eval()
    r = res // up
    x = torch.randn(n, i, r, r, device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)
    w = torch.randn(n, 512, device="cuda")
    with torch.no_grad():
        ours = timed(lambda: layer(x, w, noise_mode='const'))
        wg = torch.randn(n * o, i, 3, 3, device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)
        xg = x.reshape(1, n * i, r, r).contiguous(memory_format=torch.channels_last)
        if up == 1:
            lib = timed(lambda: torch.nn.functional.conv2d(xg, wg, padding=1, groups=n))
        else:
            wt = wg.reshape(n, o, i, 3, 3).transpose(1, 2).reshape(n * i, o, 3, 3).contiguous(memory_format=torch.channels_last)
            lib = timed(lambda: torch.nn.functional.conv_transpose2d(xg, wt, stride=2, groups=n))
        # the zero-change route of an unpickled generator: the reference's own modulated_conv2d statements
        # (networks_stylegan2.py:59-66,84-88) around shadow/torch_utils/ops: weights folded by torch ops, the grouped per-sample
        # convolution and the bias_act through this library, NCHW in and out
        from nerffaceediting_b200 import stylegan_ops as sg

        def shadow_route():
            styles = layer.affine(w)
            wm = layer.weight.unsqueeze(0) * styles.reshape(n, 1, -1, 1, 1)
            wm = wm * (wm.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt().reshape(n, -1, 1, 1, 1)
            y = net.conv2d_resample(x_nchw.reshape(1, -1, r, r), wm.reshape(-1, i, 3, 3).to(dtype), f=layer.resample_filter, up=up, padding=1,
                                    groups=n, flip_weight=(up == 1))
            y = y.reshape(n, -1, res, res).add_((layer.noise_const * layer.noise_strength).to(dtype))
            return sg.bias_act(y, layer.bias.to(dtype), act='lrelu', gain=layer.act_gain, clamp=256)
        x_nchw = x.contiguous()
        with torch.no_grad():
            shadow = timed(shadow_route)
    flop = 2.0 * n * r * r * 9 * i * o
    return {"layer": f"{i}->{o} @ {res}^2 up={up} {str(dtype).split('.')[-1]}", "ms": round(ours, 4), "tflops": round(flop / ours / 1e9, 1),
            "shadow_route_ms": round(shadow, 4),
            "cudnn_grouped_ms": round(lib, 4), "cudnn_tflops": round(flop / lib / 1e9, 1), "gflop": round(flop / 1e9, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--json", default=None)
    ap.add_argument("--backbone-only", type=int, default=-1, help="two calls of the backbone with this num_fp16_res and nothing else (launch list under ncu)")
    ap.add_argument("--sr-only", action="store_true", help="two calls of the fp16 SR head and nothing else (for a launch list under ncu)")
    a = ap.parse_args()
    if a.backbone_only >= 0:
        with torch.no_grad():
            bb = synth.fill_module(net.SynthesisNetwork(512, 256, 96, num_fp16_res=a.backbone_only, conv_clamp=256 if a.backbone_only else None), 5).cuda().eval()
            ws = torch.randn(a.batch, bb.num_ws, 512, device="cuda")
            for _ in range(2):
                bb(ws, noise_mode='const')
            torch.cuda.synchronize()
        return
    if a.sr_only:
        with torch.no_grad():
            sr = synth.fill_module(net.SuperresolutionHybrid8XDC(32, 512, 4, True), 3).cuda().eval()
            x = torch.randn(a.batch, 32, 64, 64, device="cuda")
            ws = torch.randn(a.batch, 14, 512, device="cuda")
            for _ in range(2):
                sr(x[:, :3].contiguous(), x, ws, noise_mode='const')
            torch.cuda.synchronize()
        return
    n = a.batch
    out = {"batch": n, "device": torch.cuda.get_device_name(0)}
    layers = []
    for dtype in (torch.float16, torch.float32):
        for (i, o, res, up) in [(256, 256, 256, 1), (128, 128, 512, 1), (256, 128, 512, 2), (32, 256, 256, 2), (512, 512, 64, 1), (512, 512, 32, 1)]:
            if dtype == torch.float32 and res > 256:
                continue
            layers.append(layer_bench(n, i, o, res, up, dtype))
            print(layers[-1], flush=True)
    out["layers"] = layers
    with torch.no_grad():
        for fp16 in (True, False):
            sr = synth.fill_module(net.SuperresolutionHybrid8XDC(32, 512, 4 if fp16 else 0, True), 3).cuda().eval()
            x = torch.randn(n, 32, 64, 64, device="cuda")
            ws = torch.randn(n, 14, 512, device="cuda")
            l0 = _lib.launch_count()
            sr(x[:, :3].contiguous(), x, ws, noise_mode='const')
            launches = _lib.launch_count() - l0
            ms = timed(lambda: sr(x[:, :3].contiguous(), x, ws, noise_mode='const'))
            out["sr8xdc_fp16" if fp16 else "sr8xdc_fp32"] = {"ms": round(ms, 3), "images_per_s": round(n / ms * 1e3, 1), "gflop_conv": round(n * 195.6, 1),
                                                               "tflops": round(n * 195.6 / ms, 1), "library_launches": launches}
            print("sr8xdc", "fp16" if fp16 else "fp32", out["sr8xdc_fp16" if fp16 else "sr8xdc_fp32"], flush=True)
        bb = synth.fill_module(net.SynthesisNetwork(512, 256, 96, num_fp16_res=4), 5).cuda().eval()
        ws = torch.randn(n, bb.num_ws, 512, device="cuda")
        ms = timed(lambda: bb(ws, noise_mode='const'))
        out["backbone_256x96_fp16res4"] = {"ms": round(ms, 3), "planes_per_s": round(n / ms * 1e3, 1)}
        print("backbone", out["backbone_256x96_fp16res4"], flush=True)
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
