import os, sys, torch
sys.path.insert(0, '/root/repo')
import synth_inputs as synth
from nerffaceediting_b200 import networks as net, stylegan_ops as sg
from profiles.bench_conv import timed
n, i, o, res, dtype = 8, 256, 256, 256, torch.float16
layer = synth.fill_module(net.SynthesisLayer(i, o, w_dim=512, resolution=res, conv_clamp=256), 11).cuda().eval()
x = torch.randn(n, i, res, res, device='cuda', dtype=dtype)
w = torch.randn(n, 512, device='cuda')
with torch.no_grad():
    styles = layer.affine(w)
    def fold():
        wm = layer.weight.unsqueeze(0) * styles.reshape(n, 1, -1, 1, 1)
        return (wm * (wm.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt().reshape(n, -1, 1, 1, 1)).reshape(-1, i, 3, 3).to(dtype)
    wm = fold()
    xg = x.reshape(1, -1, res, res)
    print('fold weights      %.3f ms' % timed(fold))
    print('to channels_last  %.3f ms' % timed(lambda: x.contiguous(memory_format=torch.channels_last)))
    xcl = x.contiguous(memory_format=torch.channels_last)
    print('conv (cl in)      %.3f ms' % timed(lambda: net._modconv(xcl, wm.reshape(n, o, i, 3, 3), None, None, 1, None, False, True, None, 'linear', 1.0, None)))
    y = net._modconv(xcl, wm.reshape(n, o, i, 3, 3), None, None, 1, None, False, True, None, 'linear', 1.0, None)
    print('reshape to NCHW   %.3f ms' % timed(lambda: y.reshape(1, n * o, res, res)))
    yn = y.reshape(1, n * o, res, res).reshape(n, o, res, res)
    nz = (layer.noise_const * layer.noise_strength).to(dtype)
    print('add_ noise        %.3f ms' % timed(lambda: yn.add_(nz)))
    print('bias_act          %.3f ms' % timed(lambda: sg.bias_act(yn, layer.bias.to(dtype), act='lrelu', gain=layer.act_gain, clamp=256)))
    print('weight .float()   %.3f ms' % timed(lambda: wm.float()))
