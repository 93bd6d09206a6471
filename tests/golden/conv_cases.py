"""Case tables and input builders shared by make_golden_f3_conv.py (reference side, CPU) and tests/test_gpu_conv_stack.py (this
package, GPU).  `net` / `sr` are the module namespaces to build from: the reference's or nerffaceediting_b200.networks."""
import numpy as np
import torch

import synth_inputs as synth


def T(seed, shape, scale=1.0):
    return torch.from_numpy(synth.hash_normal(seed, tuple(shape)) * np.float32(scale)).float()


MODCONV = {
    "3x3":          dict(seed=100, n=2, i=32, o=48, h=16, w=16, k=3, up=1, demod=True, flip=True, noise='const'),
    "3x3_up2":      dict(seed=110, n=2, i=32, o=32, h=12, w=12, k=3, up=2, demod=True, flip=False, noise='batch'),
    "1x1_nodemod":  dict(seed=120, n=2, i=64, o=3, h=10, w=10, k=1, up=1, demod=False, flip=True, noise=None),
    "wide_ragged":  dict(seed=130, n=1, i=128, o=256, h=20, w=12, k=3, up=1, demod=True, flip=True, noise=None),
    "res4":         dict(seed=140, n=3, i=512, o=512, h=4, w=4, k=3, up=1, demod=True, flip=True, noise='const'),
    "up2_ragged":   dict(seed=150, n=1, i=16, o=64, h=9, w=21, k=3, up=2, demod=True, flip=False, noise=None),
    "3x3_noflip":   dict(seed=160, n=1, i=16, o=16, h=8, w=8, k=3, up=1, demod=True, flip=False, noise=None),
}


def modconv_inputs(c):
    s0 = c['seed']
    x = T(s0, (c['n'], c['i'], c['h'], c['w']))
    w = T(s0 + 1, (c['o'], c['i'], c['k'], c['k']))
    s = 1.0 + 0.3 * T(s0 + 2, (c['n'], c['i']))
    oh, ow = c['h'] * c['up'], c['w'] * c['up']
    noise = None
    if c['noise'] == 'const':
        noise = 0.5 * T(s0 + 3, (oh, ow))
    elif c['noise'] == 'batch':
        noise = 0.5 * T(s0 + 3, (c['n'], 1, oh, ow))
    return x, w, s, noise


LAYERS = {
    "synth_lrelu":      dict(kind='synthesis', seed=200, n=2, i=32, o=64, res=16, up=1, w_dim=48, clamp=1.5, noise_mode='const', gain=1),
    "synth_up2":        dict(kind='synthesis', seed=210, n=2, i=64, o=32, res=32, up=2, w_dim=48, clamp=2.0, noise_mode='const', gain=1),
    "synth_none_gain":  dict(kind='synthesis', seed=220, n=1, i=32, o=32, res=8, up=1, w_dim=32, clamp=None, noise_mode='none', gain=float(np.sqrt(0.5))),
    "torgb3":           dict(kind='torgb', seed=230, n=2, i=64, o=3, res=16, w_dim=48, clamp=0.9),
    "torgb96":          dict(kind='torgb', seed=240, n=1, i=32, o=96, res=24, w_dim=48, clamp=None),
}


def make_layer(net, c):
    if c['kind'] == 'synthesis':
        layer = net.SynthesisLayer(c['i'], c['o'], w_dim=c['w_dim'], resolution=c['res'], up=c['up'], conv_clamp=c['clamp'])
    else:
        layer = net.ToRGBLayer(c['i'], c['o'], w_dim=c['w_dim'], conv_clamp=c['clamp'])
    return synth.fill_module(layer, c['seed'] + 5).eval()


def layer_inputs(c):
    r = c['res'] // c.get('up', 1)
    return T(c['seed'], (c['n'], c['i'], r, r)), T(c['seed'] + 1, (c['n'], c['w_dim']))


BLOCKS = {
    "first":  dict(seed=300, n=2, i=0, o=64, res=4, img_ch=3, is_last=False, w_dim=48, clamp=None),
    "skip":   dict(seed=310, n=2, i=64, o=32, res=32, img_ch=3, is_last=False, w_dim=48, clamp=4.0),
    "last96": dict(seed=320, n=1, i=32, o=32, res=16, img_ch=96, is_last=True, w_dim=48, clamp=None),
}


def make_block(net, c):
    block = net.SynthesisBlock(c['i'], c['o'], w_dim=c['w_dim'], resolution=c['res'], img_channels=c['img_ch'], is_last=c['is_last'],
                               architecture='skip', conv_clamp=c['clamp'], use_fp16=False)
    return synth.fill_module(block, c['seed'] + 5).eval()


def block_inputs(c):
    num_ws = (1 if c['i'] == 0 else 2) + 1
    ws = T(c['seed'] + 1, (c['n'], num_ws, c['w_dim']))
    if c['i'] == 0:
        return None, None, ws
    r = c['res'] // 2
    return T(c['seed'], (c['n'], c['i'], r, r)), T(c['seed'] + 2, (c['n'], c['img_ch'], r, r)), ws


def make_synthesis(net):
    n = net.SynthesisNetwork(w_dim=64, img_resolution=32, img_channels=96, channel_base=2048, channel_max=64, num_fp16_res=0)
    return synth.fill_module(n, 400).eval()


def synthesis_ws(n):
    return T(401, (1, n.num_ws, 64))


def make_mapping(net):
    m = net.MappingNetwork(z_dim=64, c_dim=25, w_dim=64, num_ws=6, num_layers=2)
    return synth.fill_module(m, 500).eval()


def mapping_inputs():
    return T(501, (3, 64)), T(502, (3, 25))


SR = {'2X': dict(cls='SuperresolutionHybrid2X', res=128, in_res=64, feed_res=48, n=2, seed=600),
      '8XDC': dict(cls='SuperresolutionHybrid8XDC', res=512, in_res=128, feed_res=64, n=1, seed=700)}


def make_sr(sr, which, sr_num_fp16_res=0):
    c = SR[which]
    m = getattr(sr, c['cls'])(channels=32, img_resolution=c['res'], sr_num_fp16_res=sr_num_fp16_res, sr_antialias=True)
    return synth.fill_module(m, c['seed'] + 5).eval()


def sr_inputs(which):
    """(rgb, x, ws): the renderer's feature image at `feed_res` (so the bilinear pre-resize runs), its first three channels, and ws."""
    c = SR[which]
    x = T(c['seed'], (c['n'], 32, c['feed_res'], c['feed_res']))
    ws = T(c['seed'] + 1, (c['n'], 14, 512))
    return x[:, :3].contiguous(), x, ws


# ---- the whole generator (training/triplane.py:18-165; BASELINE configs[1]) with a narrow backbone so that the CPU reference is quick
GENERATORS = {
    'g128': dict(seed=800, n=2, img_resolution=128, sr='SuperresolutionHybrid2X'),
    'g512': dict(seed=900, n=1, img_resolution=512, sr='SuperresolutionHybrid8XDC'),
}


def generator_rendering_kwargs(c):
    rk = dict(synth.FFHQ_RENDERING_OPTIONS)
    rk.update(superresolution_module='training.superresolution.' + c['sr'], sr_antialias=True, superresolution_noise_mode='none',
              c_gen_conditioning_zero=False, c_scale=1.0, decoder_lr_mul=1, nfe_deterministic=True)
    return rk


def make_generator(cls, which):
    c = GENERATORS[which]
    g = cls(z_dim=64, c_dim=25, w_dim=512, img_resolution=c['img_resolution'], img_channels=3, sr_num_fp16_res=0,
            mapping_kwargs=dict(num_layers=2), rendering_kwargs=generator_rendering_kwargs(c), channel_base=4096, channel_max=32, num_fp16_res=0)
    return synth.fill_module(g, c['seed'] + 5).eval()


def generator_inputs(which):
    c = GENERATORS[which]
    z = T(c['seed'], (c['n'], 64))
    c2w, k = synth.camera_sweep(c['n'])
    cam = torch.cat([c2w.reshape(c['n'], 16), k.reshape(c['n'], 9)], dim=1).float()
    pts = 0.4 * T(c['seed'] + 1, (c['n'], 600, 3))
    return z, cam, pts
