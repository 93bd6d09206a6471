"""Case tables of the plugin fixtures (plugins_bias_act.npz, plugins_upfirdn2d.npz), shared by make_golden_f3.py (reference side) and
tests/test_gpu_plugins.py (this package, GPU).  Inputs come from synth_inputs.hash_normal with the seeds make_golden_f3.py documents."""
import numpy as np

ACTIVATIONS = ['linear', 'relu', 'lrelu', 'tanh', 'sigmoid', 'elu', 'selu', 'softplus', 'swish']     # bias_act.py:23-33, in order

# (tag, act, shape, dim, has_bias, alpha, gain, clamp)
BIAS_ACT_CASES = [(f"{act}", act, (3, 8, 6, 10), 1, True, None, None, None) for act in ACTIVATIONS] + [
    ("lrelu_clamp", "lrelu", (2, 16, 8, 8), 1, True, None, None, 0.7),          # SynthesisLayer with conv_clamp (networks_stylegan2.py:325-327)
    ("lrelu_alpha_gain", "lrelu", (2, 16, 8, 8), 1, True, 0.1, 3.0, None),
    ("linear_nobias", "linear", (4, 5, 7), 1, False, None, 2.5, None),
    ("linear_clamp", "linear", (2, 3, 16, 16), 1, True, None, None, 0.5),      # ToRGBLayer (networks_stylegan2.py:350)
    ("fc_lrelu", "lrelu", (6, 32), 1, True, None, None, None),                  # FullyConnectedLayer (networks_stylegan2.py:126)
    ("sigmoid_dim0", "sigmoid", (5, 3, 4), 0, True, None, None, None),
    ("softplus_dim2", "softplus", (2, 3, 9), 2, True, None, 1.5, 2.0),
    ("swish_clamp", "swish", (2, 8, 5, 5), 1, True, None, None, 1.0),
    ("tanh_scalar_odd", "tanh", (1, 3, 5, 7), 1, True, None, None, None),       # odd sizes: the scalar (unvectorised) kernel path
]


# (tag, helper, shape, filter taps (1-D list, made 2-D by setup_filter unless separable) or None, kwargs)
UPFIRDN_CASES = [
    ("fir4_after_tconv", "upfirdn2d", (2, 5, 33, 33), [1, 3, 3, 1], dict(padding=[1, 1, 1, 1], gain=4)),      # conv2d_resample.py:128-129 (up=2 path, 16 -> 32)
    ("fir4_wide", "upfirdn2d", (1, 3, 20, 301), [1, 3, 3, 1], dict(padding=[2, 1, 2, 1])),                      # more than two column tiles
    ("upsample2", "upsample2d", (2, 3, 16, 16), [1, 3, 3, 1], dict(up=2)),                                      # SynthesisBlock skip image (networks_stylegan2.py:451)
    ("downsample2", "downsample2d", (2, 4, 16, 18), [1, 3, 3, 1], dict(down=2)),
    ("filter2d", "filter2d", (1, 2, 9, 11), [1, 2, 1], dict()),
    ("up3_down2_asym", "upfirdn2d", (1, 2, 7, 9), [1, 4, 6, 4, 1], dict(up=[3, 1], down=[1, 2], padding=[2, 3, 1, 0], gain=1.5)),
    ("crop_negative_pad", "upfirdn2d", (1, 2, 12, 12), [1, 3, 3, 1], dict(padding=[-1, 2, 3, -2])),
    ("flip_nonsym", "upfirdn2d", (1, 2, 10, 10), "nonsym", dict(padding=1, flip_filter=True)),
    ("noflip_nonsym", "upfirdn2d", (1, 2, 10, 10), "nonsym", dict(padding=1)),
    ("separable8", "upsample2d", (1, 2, 12, 12), [1, 2, 4, 6, 6, 4, 2, 1], dict(up=2)),                         # >= 8 taps: setup_filter keeps it 1-D
    ("identity_none", "upfirdn2d", (1, 2, 6, 6), None, dict(up=2, padding=[0, 1, 0, 1])),
]
