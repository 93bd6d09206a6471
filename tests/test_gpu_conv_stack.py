"""Convolution stack of the backbone / super-resolution module (csrc/nfe_modconv.cu through
nfe_modulated_conv2d) against fixtures the UNMODIFIED reference produced on CPU in fp32 (tests/golden/make_golden_f3_conv.py ->
conv_stack.npz).  Networks are rebuilt on both sides from synth_inputs.fill_module with the same seeds.

Tolerances (max |a-b| / max |b|, `rel_err`, and the element-wise `elem_err` with a 5e-2 floor at ten times the figure):
  fp32 activations (bf16 hi/lo split operands, three MMAs per product, fp32 accumulate): 1e-4 for single layers, 3e-4 through a
  whole network (the 8XDC head is 5 convolutions over up to 2304-term sums);
  fp16 activations (what the reference's own fp16 blocks compute with): 1e-2 against the fp32 fixture."""
import os
import sys

import numpy as np
import pytest
import torch

from _util import elem_err, golden, rel_err

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import conv_cases as cases  # noqa: E402

pytestmark = pytest.mark.gpu
TOL32, TOL32_NET, TOL16 = 1e-4, 3e-4, 1e-2


def cuda(t):
    return None if t is None else t.cuda()


def check(y, ref, tol, what):
    y = y.float().cpu().numpy()
    assert y.shape == ref.shape, (what, y.shape, ref.shape)
    assert np.isfinite(y).all(), what
    e, ee = rel_err(y, ref), elem_err(y, ref, floor=5e-2)
    assert e < tol and ee < 10 * tol, (what, e, ee)


@pytest.mark.parametrize("tag", list(cases.MODCONV))
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_modulated_conv2d(tag, dtype):
    from nerffaceediting_b200 import networks as net
    from nerffaceediting_b200 import stylegan_ops as sg
    c = cases.MODCONV[tag]
    x, w, s, noise = cases.modconv_inputs(c)
    f = sg.setup_filter([1, 3, 3, 1]).cuda() if c['up'] == 2 else None
    with torch.no_grad():
        y = net.modulated_conv2d(cuda(x).to(dtype), cuda(w), cuda(s), noise=cuda(noise), up=c['up'], padding=c['k'] // 2, resample_filter=f,
                                 demodulate=c['demod'], flip_weight=c['flip'])
    assert y.dtype == dtype and y.is_contiguous(memory_format=torch.channels_last)
    check(y, golden("conv_stack")[f"modconv.{tag}"], TOL32 if dtype == torch.float32 else TOL16, tag)


def test_modulated_conv2d_accepts_any_input_layout_and_rejects_what_it_cannot_do():
    from nerffaceediting_b200 import networks as net
    c = cases.MODCONV["3x3"]
    x, w, s, noise = cases.modconv_inputs(c)
    ref = golden("conv_stack")["modconv.3x3"]
    with torch.no_grad():
        y_cl = net.modulated_conv2d(cuda(x).contiguous(memory_format=torch.channels_last), cuda(w), cuda(s), noise=cuda(noise), padding=1)
        y_nchw = net.modulated_conv2d(cuda(x), cuda(w).contiguous(memory_format=torch.channels_last), cuda(s), noise=cuda(noise), padding=1)
    assert torch.equal(y_cl, y_nchw)
    check(y_cl, ref, TOL32, "layouts")
    with pytest.raises(RuntimeError):
        net.modulated_conv2d(x, w, s, padding=1)                                            # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        net.modulated_conv2d(cuda(x).requires_grad_(True), cuda(w), cuda(s), padding=1)     # forward-only
    with pytest.raises(AssertionError):
        net.modulated_conv2d(cuda(x), cuda(w), cuda(s), padding=1, down=2)
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            net.modulated_conv2d(cuda(x)[:, :24], cuda(w)[:, :24], cuda(s)[:, :24], padding=1)   # in_channels not a multiple of 16


@pytest.mark.parametrize("tag", list(cases.LAYERS))
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_layers(tag, dtype):
    from nerffaceediting_b200 import networks as net
    c = cases.LAYERS[tag]
    layer = cases.make_layer(net, c).cuda()
    x, w = cases.layer_inputs(c)
    with torch.no_grad():
        if c['kind'] == 'synthesis':
            y = layer(cuda(x).to(dtype), cuda(w), noise_mode=c['noise_mode'], gain=c['gain'])
        else:
            y = layer(cuda(x).to(dtype), cuda(w))
    check(y, golden("conv_stack")[f"layer.{tag}"], TOL32 if dtype == torch.float32 else TOL16, tag)


@pytest.mark.parametrize("tag", list(cases.BLOCKS))
def test_blocks(tag):
    from nerffaceediting_b200 import networks as net
    c = cases.BLOCKS[tag]
    block = cases.make_block(net, c).cuda()
    x, img, ws = cases.block_inputs(c)
    with torch.no_grad():
        x2, img2 = block(cuda(x), cuda(img), cuda(ws), noise_mode='const')
    g = golden("conv_stack")
    check(x2, g[f"block.{tag}.x"], TOL32_NET, tag + ".x")
    check(img2, g[f"block.{tag}.img"], TOL32_NET, tag + ".img")
    assert img2.dtype == torch.float32 and img2.is_contiguous()


def test_synthesis_network_and_mapping():
    from nerffaceediting_b200 import networks as net
    g = golden("conv_stack")
    n = cases.make_synthesis(net).cuda()
    with torch.no_grad():
        img = n(cuda(cases.synthesis_ws(n)), noise_mode='const')
    check(img, g["synthesis.img"], TOL32_NET, "synthesis")
    m = cases.make_mapping(net).cuda()
    z, cnd = cases.mapping_inputs()
    with torch.no_grad():
        check(m(cuda(z), cuda(cnd), truncation_psi=0.7, truncation_cutoff=4), g["mapping.ws"], 1e-5, "mapping")
        check(m(cuda(z), cuda(cnd)), g["mapping.ws_plain"], 1e-5, "mapping plain")


def test_superresolution_2x():
    from nerffaceediting_b200 import networks as net
    m = cases.make_sr(net, '2X').cuda()
    rgb, x, ws = cases.sr_inputs('2X')
    with torch.no_grad():
        out = m(cuda(rgb), cuda(x), cuda(ws), noise_mode='const')
    check(out, golden("conv_stack")["sr2x.rgb"], TOL32_NET, "sr2x")


@pytest.mark.parametrize("fp16", [False, True])
def test_superresolution_8xdc_full_size(fp16):
    """The default 512 x 512 head at full size (superresolution.py:264-290): 64 -> 128 (bilinear) -> 256 (256 ch) -> 512 (128 ch)."""
    from nerffaceediting_b200 import networks as net
    g = golden("conv_stack")
    m = cases.make_sr(net, '8XDC', sr_num_fp16_res=4 if fp16 else 0).cuda()
    rgb, x, ws = cases.sr_inputs('8XDC')
    with torch.no_grad():
        out = m(cuda(rgb), cuda(x), cuda(ws), noise_mode='const')
    assert out.shape == (1, 3, 512, 512) and out.dtype == torch.float32
    tol = TOL16 if fp16 else TOL32_NET
    if fp16:
        # with these random weights the fp16 blocks saturate at conv_clamp = 256 in places, as the reference's would: compare the bulk
        assert rel_err(out[:, :, ::4, ::4].cpu().numpy(), g["sr8xdc.rgb_s4"]) < 5e-2
        return
    check(out[:, :, ::4, ::4], g["sr8xdc.rgb_s4"], tol, "sr8xdc subsampled")
    check(out[:, :, 253:259, :], g["sr8xdc.rgb_rows"], tol, "sr8xdc rows")
    mom = g["sr8xdc.moments"]
    assert abs(out.double().mean().item() - mom[0]) < tol * mom[2] and abs(out.double().std().item() - mom[1]) < tol * mom[2]


def _torch_modconv(x, w, s, noise, up, demodulate, flip_weight, f):
    """fp64 restatement of networks_stylegan2.py:34-91 + conv2d_resample.py:117-139 with torch's own convolutions (test oracle for the
    shapes no fixture covers; the fixtures above pin the restatement's cases to the reference itself)."""
    x, w, s = x.double(), w.double(), s.double()
    n, i, h, wd = x.shape
    o, _, k, _ = w.shape
    wm = w.unsqueeze(0) * s.reshape(n, 1, i, 1, 1)
    if demodulate:
        wm = wm * (wm.square().sum(dim=[2, 3, 4], keepdim=True) + 1e-8).rsqrt()
    outs = []
    for b in range(n):
        wb = wm[b] if flip_weight else wm[b].flip([2, 3])
        if up == 1:
            outs.append(torch.nn.functional.conv2d(x[b:b + 1], wb, padding=k // 2))
        else:
            t = torch.nn.functional.conv_transpose2d(x[b:b + 1], wm[b].transpose(0, 1) if not flip_weight else wm[b].flip([2, 3]).transpose(0, 1), stride=2)
            ff = (f.double() * 4).flip([0, 1])[None, None].repeat(o, 1, 1, 1)
            fw = f.shape[-1]
            p0, p1 = k // 2 + (fw + 1) // 2 - (k - 1), k // 2 + (fw - 2) // 2 - (k - 2)        # conv2d_resample.py:97-101,124-127
            outs.append(torch.nn.functional.conv2d(torch.nn.functional.pad(t, [p0, p1, p0, p1]), ff, groups=o))
    y = torch.cat(outs)
    return y + noise.double() if noise is not None else y


@pytest.mark.parametrize("shape", [
    # n, in, out, h, w, k, up        (kc = 16 / 32 / 64 chunks, ragged windows, several N tiles, twin / pair / single-window CTAs)
    (3, 48, 24, 5, 7, 3, 1), (1, 96, 40, 33, 9, 3, 1), (2, 16, 8, 4, 4, 1, 1), (2, 64, 384, 8, 24, 3, 1), (1, 32, 128, 40, 40, 3, 1),
    (5, 48, 16, 7, 5, 3, 2), (1, 160, 96, 16, 8, 3, 2), (2, 256, 256, 32, 32, 3, 1), (9, 128, 128, 16, 16, 3, 1), (1, 32, 3, 50, 50, 1, 1),
    # odd numbers of 16-column groups (a last staged segment of 1 or 3 groups), a channel count that rules out 16-byte stores
    (2, 32, 48, 9, 9, 3, 1), (1, 64, 112, 12, 20, 3, 1), (1, 32, 112, 8, 8, 3, 2), (1, 32, 20, 6, 6, 3, 1), (2, 16, 16, 9, 5, 1, 1),
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_modulated_conv2d_shapes_vs_torch(shape, dtype):
    from nerffaceediting_b200 import networks as net
    from nerffaceediting_b200 import stylegan_ops as sg
    n, i, o, h, w, k, up = shape
    g = torch.Generator(device="cpu").manual_seed(sum(shape))
    x = torch.randn(n, i, h, w, generator=g).cuda()
    wt = torch.randn(o, i, k, k, generator=g).cuda()
    s = (1.0 + 0.3 * torch.randn(n, i, generator=g)).cuda()
    noise = (0.3 * torch.randn(n, 1, h * up, w * up, generator=g)).cuda()
    f = sg.setup_filter([1, 3, 3, 1]).cuda()
    demod, flip = (k == 3), (up == 1)
    xin = x.to(dtype)
    ref = _torch_modconv(xin.float(), wt, s, noise, up, demod, flip, f)
    with torch.no_grad():
        y = net.modulated_conv2d(xin, wt, s, noise=noise, up=up, padding=k // 2, resample_filter=f if up == 2 else None, demodulate=demod, flip_weight=flip)
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    e = rel_err(y.float().cpu().numpy(), ref.float().cpu().numpy())
    assert e < tol, (shape, dtype, e)


@pytest.mark.parametrize("case", [
    # in, out, res, up, batch   (enough windows for the persistent CTAs: several N tiles = the bias table follows the window's tile;
    #                            N <= 128 = two accumulator sets; up = 2 = one set of four phase accumulators)
    (32, 512, 128, 1, 4), (64, 128, 256, 1, 2), (32, 64, 256, 2, 4), (32, 256, 128, 1, 8),
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_synthesis_layer_large_grids_vs_torch(case, dtype):
    """SynthesisLayer (networks_stylegan2.py:276-330) on grids of several windows per SM against an fp64 composition of the same
    formulas: modulated convolution, noise, bias, lrelu * sqrt(2), clamp."""
    from nerffaceediting_b200 import networks as net
    import synth_inputs as synth
    i, o, res, up, n = case
    layer = synth.fill_module(net.SynthesisLayer(i, o, w_dim=64, resolution=res, up=up, conv_clamp=256), sum(case)).cuda().eval()
    g = torch.Generator(device="cpu").manual_seed(7 + sum(case))
    x = torch.randn(n, i, res // up, res // up, generator=g).cuda().to(dtype)
    w = torch.randn(n, 64, generator=g).cuda()
    with torch.no_grad():
        y = layer(x, w, noise_mode='const')
        styles = layer.affine(w)
        noise = (layer.noise_const * layer.noise_strength)[None, None].expand(n, 1, res, res)
        ref = _torch_modconv(x.float(), layer.weight, styles, noise, up, True, up == 1, layer.resample_filter)
        ref = ref + layer.bias.double().reshape(1, -1, 1, 1)
        ref = (torch.nn.functional.leaky_relu(ref, 0.2) * (2.0 ** 0.5)).clamp(-256, 256)
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    e = rel_err(y.float().cpu().numpy(), ref.float().cpu().numpy())
    assert e < tol, (case, dtype, e)


@pytest.mark.parametrize("tag", ["3x3", "3x3_up2", "1x1_nodemod", "wide_ragged"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_shadow_conv2d_resample_serves_the_fused_modulated_conv2d(tag, dtype):
    """shadow/torch_utils/ops/conv2d_resample.py is what an UNPICKLED generator's modulated_conv2d calls (networks_stylegan2.py:84-88:
    x [1, N*I, H, W], per-sample weights [N*O, I, k, k], groups = N): the grouped form through the per-item-weight path of the kernel,
    against the reference fixture of the same layer."""
    shadow = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "shadow")
    if shadow not in sys.path:
        sys.path.insert(0, shadow)
    for name in [m for m in sys.modules if m == "torch_utils" or m.startswith("torch_utils.")]:
        del sys.modules[name]
    from torch_utils.ops import bias_act, conv2d_resample, upfirdn2d
    assert "shadow" in conv2d_resample.__file__ and "shadow" in bias_act.__file__ and "shadow" in upfirdn2d.__file__
    c = cases.MODCONV[tag]
    x, w, s, noise = (cuda(t) for t in cases.modconv_inputs(c))
    n, o, i, k = c['n'], c['o'], c['i'], c['k']
    f = upfirdn2d.setup_filter([1, 3, 3, 1]).cuda() if c['up'] == 2 else None
    with torch.no_grad():
        wm = w.unsqueeze(0) * s.reshape(n, 1, -1, 1, 1)                      # networks_stylegan2.py:59-66
        if c['demod']:
            wm = wm * (wm.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt().reshape(n, -1, 1, 1, 1)
        xg = x.to(dtype).reshape(1, -1, *x.shape[2:])
        y = conv2d_resample.conv2d_resample(x=xg, w=wm.reshape(-1, i, k, k).to(dtype), f=f, up=c['up'], padding=k // 2, groups=n, flip_weight=c['flip'])
        y = y.reshape(n, -1, *y.shape[2:])
        if noise is not None:
            y = y.add_(noise.to(y.dtype))
    assert y.shape[1] == o
    check(y, golden("conv_stack")[f"modconv.{tag}"], TOL32 if dtype == torch.float32 else TOL16, "shadow " + tag)


@pytest.mark.parametrize("taps", ["1331", "121", "random4x4", "12"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_up2_resample_filters(taps, dtype):
    """up = 2 with other resample filters than the default: a 3-tap one (different paddings), a non-separable 4 x 4 one (the filter
    pass's direct path) and a 2-tap one, against the fp64 torch restatement."""
    from nerffaceediting_b200 import networks as net
    from nerffaceediting_b200 import stylegan_ops as sg
    g = torch.Generator(device="cpu").manual_seed(11)
    f = {"1331": sg.setup_filter([1, 3, 3, 1]), "121": sg.setup_filter([1, 2, 1]), "12": sg.setup_filter([1, 2]),
         "random4x4": torch.rand(4, 4, generator=g) / 8}[taps].cuda()
    n, i, o, h, w = 2, 32, 48, 9, 13
    x = torch.randn(n, i, h, w, generator=g).cuda().to(dtype)
    wt = torch.randn(o, i, 3, 3, generator=g).cuda()
    s = (1.0 + 0.3 * torch.randn(n, i, generator=g)).cuda()
    ref = _torch_modconv(x.float(), wt, s, None, 2, True, False, f)
    with torch.no_grad():
        y = net.modulated_conv2d(x, wt, s, up=2, padding=1, resample_filter=f, demodulate=True, flip_weight=False)
    assert y.shape == ref.shape
    e = rel_err(y.float().cpu().numpy(), ref.float().cpu().numpy())
    assert e < (1e-4 if dtype == torch.float32 else 1e-2), (taps, dtype, e)


def test_modulated_conv2d_c_abi_rejects_bad_arguments():
    import ctypes
    from nerffaceediting_b200 import _lib
    lib = _lib.load()
    x = torch.zeros(1, 8, 8, 16, device="cuda")
    w = torch.zeros(16, 16, 3, 3, device="cuda")
    st = torch.ones(1, 16, device="cuda")
    y = torch.zeros(1, 8, 8, 16, device="cuda")
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")

    def call(**kw):
        a = dict(x=x.data_ptr(), weight=w.data_ptr(), styles=st.data_ptr(), y=y.data_ptr(), batch=1, in_ch=16, out_ch=16, in_h=8, in_w=8, ksize=3,
                 up=1, demodulate=1, flip_weight=1, act=1, gain=1.0, clamp=-1.0, dtype=0)
        a.update(kw)
        return lib.nfe_modulated_conv2d(_lib.NfeModconvArgs(**a), ws.data_ptr(), ws.numel(), None)
    assert call() == 0
    assert call(act=7) != 0 and b"act" in lib.nfe_last_error()                     # only linear / relu / lrelu are fused
    assert call(up=2) != 0 and b"filter" in lib.nfe_last_error()                   # up = 2 without a resample filter
    assert call(x=None) != 0
    assert lib.nfe_modulated_conv2d(_lib.NfeModconvArgs(x=x.data_ptr(), weight=w.data_ptr(), styles=st.data_ptr(), y=y.data_ptr(), batch=1, in_ch=16,
                                                        out_ch=16, in_h=8, in_w=8, ksize=3, up=1, demodulate=1, flip_weight=1, act=1, gain=1.0,
                                                        clamp=-1.0, dtype=0), ws.data_ptr(), 16, None) != 0            # workspace too small
    torch.cuda.synchronize()
