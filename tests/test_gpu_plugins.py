"""bias_act / upfirdn2d kernels (first step) through the C ABI, against fixtures the UNMODIFIED reference's
`impl='ref'` paths produced (tests/golden/make_golden_f3.py -> plugins_bias_act.npz, plugins_upfirdn2d.npz): values, first-order
gradients and, for bias_act, the second-order terms of the reference's nested autograd Functions (bias_act.py:126-209,
upfirdn2d.py:218-273).  Tolerances: 1e-5 relative to the tensor's largest magnitude for fp32 (the kernels and the reference both
compute in fp32; only the accumulation order differs), 4e-3 for fp16 / bf16 storage.

One deliberate difference from the `ref` fixtures: for act='linear' with a clamp (ToRGBLayer) the reference's CUDA plugin does NOT mask the
gradient where the output was clamped — `linear` saves no output (bias_act.py:23,153-156: ref=''), so bias_act.cu:144-145 sees yref = 0
everywhere — while its pure-PyTorch `ref` path differentiates through torch.clamp and does.  This library replaces the plugin, i.e. what
GPU training runs, so that case is checked against the plugin's behaviour (dx = dy * gain)."""
import os
import sys

import numpy as np
import pytest
import torch

from _util import golden, rel_err

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import plugin_cases as pc  # noqa: E402
import synth_inputs as synth  # noqa: E402

pytestmark = pytest.mark.gpu
TOL32, TOL16 = 1e-5, 4e-3


def T(a, **kw):
    return torch.from_numpy(np.ascontiguousarray(a)).float().cuda().requires_grad_(kw.get("grad", False))


@pytest.mark.parametrize("k", range(len(pc.BIAS_ACT_CASES)), ids=[c[0] for c in pc.BIAS_ACT_CASES])
@pytest.mark.parametrize("channels_last", [False, True])
def test_bias_act_values_and_gradients(k, channels_last):
    from nerffaceediting_b200 import stylegan_ops as sg
    tag, act, shape, dim, has_b, alpha, gain, clamp = pc.BIAS_ACT_CASES[k]
    if channels_last and len(shape) != 4:
        pytest.skip("channels-last needs a 4-D tensor")
    g = golden("plugins_bias_act")
    x = T(synth.hash_normal(1000 + k, shape) * 1.5, grad=True)
    xin = x.contiguous(memory_format=torch.channels_last) if channels_last else x
    b = T(synth.hash_normal(1100 + k, (shape[dim],)) * 0.5, grad=True) if has_b else None
    w, w2 = T(synth.hash_normal(1200 + k, shape)), T(synth.hash_normal(1300 + k, shape))
    y = sg.bias_act(xin, b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp)
    assert rel_err(y.detach().cpu().numpy(), g[f"{tag}.y"]) < TOL32
    ins = [x] + ([b] if has_b else [])
    grads = torch.autograd.grad((y * w).sum(), ins, create_graph=True)
    ref_dx, ref_db = g[f"{tag}.dx"], g[f"{tag}.db"] if has_b else None
    if act == 'linear' and clamp is not None:            # the plugin's unmasked gradient (see the module docstring)
        ref_dx = w.cpu().numpy() * (gain if gain is not None else 1.0)
        ref_db = ref_dx.sum(axis=tuple(i for i in range(len(shape)) if i != dim))
    assert rel_err(grads[0].detach().cpu().numpy(), ref_dx) < TOL32
    if has_b:
        assert rel_err(grads[1].detach().cpu().numpy(), ref_db) < 2e-5
    if grads[0].requires_grad:
        g2 = torch.autograd.grad((grads[0] * w2).sum(), ins, allow_unused=True)
        ddx = g2[0].cpu().numpy() if g2[0] is not None else np.zeros(shape, np.float32)
        ref = g[f"{tag}.ddx"]
        assert np.max(np.abs(ddx - ref)) <= TOL32 * max(np.max(np.abs(ref)), 1.0)
        if has_b:
            ddb = g2[1].cpu().numpy() if g2[1] is not None else np.zeros((shape[dim],), np.float32)
            refb = g[f"{tag}.ddb"]
            assert np.max(np.abs(ddb - refb)) <= 4e-5 * max(np.max(np.abs(refb)), 1.0)
    else:
        assert not g[f"{tag}.ddx"].any()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_bias_act_half_storage(dtype):
    from nerffaceediting_b200 import stylegan_ops as sg
    g = golden("plugins_bias_act")
    k = [c[0] for c in pc.BIAS_ACT_CASES].index("lrelu_clamp")
    tag, act, shape, dim, has_b, alpha, gain, clamp = pc.BIAS_ACT_CASES[k]
    x = T(synth.hash_normal(1000 + k, shape) * 1.5).to(dtype)
    b = T(synth.hash_normal(1100 + k, (shape[dim],)) * 0.5).to(dtype)
    y = sg.bias_act(x, b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp)
    assert y.dtype == dtype
    assert rel_err(y.float().cpu().numpy(), g[f"{tag}.y"]) < (TOL16 if dtype == torch.float16 else 2e-2)


def test_bias_act_argument_errors():
    from nerffaceediting_b200 import stylegan_ops as sg
    x = torch.zeros(2, 4, 3, 3, device="cuda")
    with pytest.raises(RuntimeError):
        sg.bias_act(x, torch.zeros(5, device="cuda"))                    # bias length != channels (bias_act.cpp:48)
    with pytest.raises(RuntimeError):
        sg.bias_act(torch.zeros(2, 4), torch.zeros(4))                   # CPU tensors: no fallback
    with pytest.raises(AssertionError):
        sg.bias_act(x, act="gelu")
    assert sg.bias_act(torch.zeros(0, 4, device="cuda"), torch.zeros(4, device="cuda")).shape == (0, 4)


def _filter(taps):
    from nerffaceediting_b200 import stylegan_ops as sg
    if taps is None:
        return None
    if taps == "nonsym":
        return torch.from_numpy(np.array([[1, 2, 0], [0, 3, 5], [7, 0, 1]], np.float32) / 19.0)
    return sg.setup_filter(taps)


@pytest.mark.parametrize("k", range(len(pc.UPFIRDN_CASES)), ids=[c[0] for c in pc.UPFIRDN_CASES])
@pytest.mark.parametrize("channels_last", [False, True])
def test_upfirdn2d_values_and_gradients(k, channels_last):
    from nerffaceediting_b200 import stylegan_ops as sg
    tag, helper, shape, taps, kw = pc.UPFIRDN_CASES[k]
    g = golden("plugins_upfirdn2d")
    x = T(synth.hash_normal(2000 + k, shape), grad=True)
    xin = x.contiguous(memory_format=torch.channels_last) if channels_last else x
    f = _filter(taps)
    f = f.cuda() if f is not None else None
    y = getattr(sg, helper)(xin, f, **kw)
    ref = g[f"{tag}.y"]
    assert tuple(y.shape) == ref.shape
    assert rel_err(y.detach().cpu().numpy(), ref) < TOL32
    w = T(synth.hash_normal(2100 + k, ref.shape))
    dx, = torch.autograd.grad((y * w).sum(), [x])
    assert rel_err(dx.cpu().numpy(), g[f"{tag}.dx"]) < TOL32


def test_upfirdn2d_half_storage_and_errors():
    from nerffaceediting_b200 import stylegan_ops as sg
    g = golden("plugins_upfirdn2d")
    k = [c[0] for c in pc.UPFIRDN_CASES].index("fir4_after_tconv")
    tag, helper, shape, taps, kw = pc.UPFIRDN_CASES[k]
    x = T(synth.hash_normal(2000 + k, shape)).half()
    y = sg.upfirdn2d(x, _filter(taps).cuda(), **kw)
    assert y.dtype == torch.float16 and rel_err(y.float().cpu().numpy(), g[f"{tag}.y"]) < TOL16
    with pytest.raises(RuntimeError):
        sg.upfirdn2d(torch.zeros(1, 1, 2, 2, device="cuda"), sg.setup_filter([1, 3, 3, 1]).cuda())    # image smaller than the filter
    with pytest.raises(RuntimeError):
        sg.upfirdn2d(torch.zeros(1, 1, 8, 8), None)                                                     # CPU tensor


def test_setup_filter_matches_the_reference_forms():
    from nerffaceediting_b200 import stylegan_ops as sg
    f = sg.setup_filter([1, 3, 3, 1])
    assert f.shape == (4, 4) and abs(float(f.sum()) - 1.0) < 1e-6                     # upfirdn2d.py:95-111: outer product, normalised
    assert sg.setup_filter([1, 2, 4, 6, 6, 4, 2, 1]).ndim == 1                        # >= 8 taps stay separable
    assert sg.setup_filter(None).shape == (1, 1)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 96, 33, 17), (1, 3, 64, 64), (3, 200, 5, 7), (1, 64, 128, 128)])
def test_layout_convert_is_bit_exact(dtype, shape):
    """nfe_layout_convert (contiguous NCHW <-> channels-last through a shared-memory tile) against torch's own copies."""
    from nerffaceediting_b200 import networks as net
    x = torch.randn(shape, device="cuda").to(dtype)
    cl = net._channels_last(x)
    assert cl.is_contiguous(memory_format=torch.channels_last) and torch.equal(cl, x)
    assert torch.equal(cl.permute(0, 2, 3, 1).contiguous(), x.permute(0, 2, 3, 1).contiguous())
    back = net._nchw(cl)
    assert back.is_contiguous() and torch.equal(back, x)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("shape", [(2, 96, 33, 17), (8, 3, 64, 64), (3, 200, 5, 7), (1, 4, 9, 9)])
def test_image_accumulate_is_the_reference_s_two_steps(dtype, shape):
    """nfe_image_accumulate == `y.to(float32, contiguous_format); img.add_(y)` (networks_stylegan2.py:456-457), bit for bit."""
    from nerffaceediting_b200 import networks as net
    y = torch.randn(shape, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    img = torch.randn(shape, device="cuda")
    want = img.clone().add_(y.to(dtype=torch.float32, memory_format=torch.contiguous_format))
    got = net._accumulate_image(img, y)
    assert got is img and got.is_contiguous() and torch.equal(got, want)
    first = net._accumulate_image(None, y)                      # first block: the image IS y
    assert first.dtype == torch.float32 and first.is_contiguous() and torch.equal(first, y.float())
