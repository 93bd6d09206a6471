"""Backward parity (BASELINE config 4, SURVEY.md §8 row "backward"): gradients of a weighted sum of every
renderer output w.r.t. both plane tensors and the decoder parameters, against the gradients the unmodified
reference's autograd graph produced (tests/golden/backward.npz, made by tests/golden/make_golden.py).

Tolerance: 1e-4 relative (max-abs error over max-abs reference) on fp32, the forward's own bar; the scatter-add
is atomic, so two runs may differ at the ulp level, like the reference's grid_sampler_2d_backward.
"""
import numpy as np
import pytest
import torch

import synth_inputs as synth
from _util import golden, rel_err
from test_gpu_parity import N, T, torch_decoder

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _loss(g, tag, out, dev):
    if len(out) == 3:
        rgb, depth, wsum = out
        return (rgb * T(g[f"{tag}.wr"], dev)).sum() + (depth * T(g[f"{tag}.wd"], dev)).sum() + (wsum * T(g[f"{tag}.ww"], dev)).sum()
    rgb, seg, depth, wsum = out
    return ((rgb * T(g[f"{tag}.wr"], dev)).sum() + (seg * T(g[f"{tag}.ws"], dev)).sum() + (depth * T(g[f"{tag}.wd"], dev)).sum()
            + (wsum * T(g[f"{tag}.ww"], dev)).sum())


def _run(g, tag, kind, opts, dev, precision="fp32"):
    from nerffaceediting_b200.ray_sampler import RaySampler
    from nerffaceediting_b200.renderer import DisentangledImportanceRenderer, ImportanceRenderer
    n, hw, res = 2, 16, 8
    raw = T(synth.hash_normal(300, (n, 96, hw, hw)) * np.float32(1.5) - np.float32(0.3), dev)
    with torch.no_grad():
        o, d = RaySampler()(T(g["cam2world"], dev), T(g["intrinsics"], dev), res)
    dec = torch_decoder(g, f"{tag}.dec", kind, 1.0, dev)
    opts = dict(opts, nfe_deterministic=True, nfe_precision=precision)
    planes = raw.view(n, 3, 32, hw, hw).clone().requires_grad_(True)
    if kind == "osg":
        out = ImportanceRenderer()(planes, dec, o, d, opts)
        norm = None
    else:
        norm = T(g[f"{tag}.norm"], dev).requires_grad_(True)
        out = DisentangledImportanceRenderer()(norm, planes, dec, o, d, opts)
    loss = _loss(g, tag, out, dev)
    loss.backward()
    return loss, norm, planes, dec


BASE = dict(synth.FFHQ_RENDERING_OPTIONS, depth_resolution=12, depth_resolution_importance=12)
CASES = {"dis": ("dis", BASE), "osg": ("osg", BASE), "dis_wb": ("dis", dict(BASE, white_back=True))}


@pytest.mark.parametrize("tag", list(CASES))
def test_render_backward_vs_reference_autograd(dev, tag):
    g = golden("backward")
    kind, opts = CASES[tag]
    loss, norm, planes, dec = _run(g, tag, kind, opts, dev)
    assert abs(float(loss.detach()) - float(g[f"{tag}.loss"])) <= TOL * max(1.0, abs(float(g[f"{tag}.loss"])))
    assert planes.grad is not None and planes.grad.shape == g[f"{tag}.g_planes"].shape
    assert rel_err(N(planes.grad), g[f"{tag}.g_planes"]) < TOL
    if norm is not None:
        assert rel_err(N(norm.grad), g[f"{tag}.g_norm"]) < TOL
    for name, p in dec.named_parameters():
        ref = g[f"{tag}.g_dec.{name}"]
        assert p.grad is not None and p.grad.shape == ref.shape, name
        assert rel_err(N(p.grad), ref) < TOL, name


def test_render_backward_tensor_core_forward(dev):
    """Training with the tensor-core forward (bf16x3): gradients still within the fp32 bar."""
    g = golden("backward")
    loss, norm, planes, dec = _run(g, "dis", "dis", BASE, dev, precision="bf16x3")
    assert rel_err(N(planes.grad), g["dis.g_planes"]) < 2 * TOL
    assert rel_err(N(norm.grad), g["dis.g_norm"]) < 2 * TOL


def test_normalize_plane_backward(dev):
    from nerffaceediting_b200 import triplane
    g = golden("backward")
    x = T(synth.hash_normal(301, (2, 96, 8, 8)) * np.float32(1.5) - np.float32(0.3), dev).requires_grad_(True)
    norm, mean, std = triplane.normalize_plane(x)
    ((norm * T(g["norm.gw"], dev)).sum() + (mean * T(g["norm.gm"], dev)).sum() + (std * T(g["norm.gs"], dev)).sum()).backward()
    assert rel_err(N(x.grad), g["norm.g_x"]) < 1e-5


def test_training_step_through_normalize_and_render(dev):
    """The generator's own chain: raw planes -> normalize_plane -> renderer -> loss; gradient reaches the raw planes
    through both the normalised and the raw branch, and agrees with the sum of the two golden branch gradients
    pushed through the analytic normalisation backward."""
    from nerffaceediting_b200 import triplane
    from nerffaceediting_b200.ray_sampler import RaySampler
    from nerffaceediting_b200.renderer import DisentangledImportanceRenderer
    g = golden("backward")
    n, hw, res = 2, 16, 8
    raw = T(synth.hash_normal(300, (n, 96, hw, hw)) * np.float32(1.5) - np.float32(0.3), dev).requires_grad_(True)
    with torch.no_grad():
        o, d = RaySampler()(T(g["cam2world"], dev), T(g["intrinsics"], dev), res)
    dec = torch_decoder(g, "dis.dec", "dis", 1.0, dev)
    norm, _, _ = triplane.normalize_plane(raw)
    out = DisentangledImportanceRenderer()(norm.view(n, 3, 32, hw, hw), raw.view(n, 3, 32, hw, hw), dec, o, d,
                                           dict(BASE, nfe_deterministic=True, nfe_precision="fp32"))
    _loss(g, "dis", out, dev).backward()
    # expected: d loss / d raw = g_planes + normalize_backward(g_norm), the latter by torch autograd on the formula
    x = raw.detach().cpu().double().requires_grad_(True)
    mean = x.mean(dim=(-1, -2), keepdim=True)
    std = x.var(dim=(-1, -2), keepdim=True).sqrt()
    ((x - mean) / (std + 1e-8)).backward(torch.from_numpy(g["dis.g_norm"]).double().view(n, 96, hw, hw))
    want = x.grad.float().numpy() + g["dis.g_planes"].reshape(n, 96, hw, hw)
    assert rel_err(N(raw.grad), want) < TOL


def test_backward_is_skipped_under_no_grad_and_for_frozen_inputs(dev):
    from nerffaceediting_b200.ray_sampler import RaySampler
    from nerffaceediting_b200.renderer import ImportanceRenderer
    g = golden("backward")
    with torch.no_grad():
        o, d = RaySampler()(T(g["cam2world"], dev), T(g["intrinsics"], dev), 8)
    dec = torch_decoder(g, "osg.dec", "osg", 1.0, dev)
    planes = torch.randn(2, 3, 32, 16, 16, device=dev)
    opts = dict(BASE, nfe_deterministic=True)
    with torch.no_grad():
        a = ImportanceRenderer()(planes, dec, o, d, opts)
    assert not a[0].requires_grad
    for p in dec.parameters():
        p.requires_grad_(False)
    b = ImportanceRenderer()(planes, dec, o, d, opts)    # nothing to differentiate: inference path
    assert not b[0].requires_grad
    c = ImportanceRenderer()(planes.clone().requires_grad_(True), dec, o, d, opts)
    assert c[0].requires_grad
    # frozen decoder: only the planes receive a gradient
    c[0].sum().backward()
    assert all(p.grad is None for p in dec.parameters())


def test_denormalize_plane_backward(dev):
    """out = planes * std + mean (triplane.py:66-68); statistics of one identity broadcast over the batch
    (triplane.py:98-103) get the batch-summed gradient.  Checked against autograd on the formula."""
    from nerffaceediting_b200 import triplane
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(2, 96, 8, 8, generator=gen)
    for stat_batch in (2, 1):
        m, s = torch.randn(stat_batch, 96, 1, 1, generator=gen), torch.rand(stat_batch, 96, 1, 1, generator=gen) + 0.5
        w = torch.randn(2, 96, 8, 8, generator=gen)
        ref = [t.clone().double().requires_grad_(True) for t in (x, m, s)]
        ((ref[0] * ref[2] + ref[1]) * w.double()).sum().backward()
        ours = [t.clone().to(dev).requires_grad_(True) for t in (x, m, s)]
        (triplane.denormalize_plane(*ours) * w.to(dev)).sum().backward()
        for a, b in zip(ours, ref):
            assert a.grad.shape == b.grad.shape
            assert rel_err(N(a.grad), b.grad.float().numpy()) < 1e-5


def test_fused_decoder_backward_matches_library_gemm_path(dev, monkeypatch):
    """The fused tcgen05 decoder backward (nfe_field_bwd) against the same backward run on library GEMMs
    (NFE_BWD_LIBRARY_GEMM=1, itself pinned to the reference's autograd above), at a size where every CTA walks several
    tiles, so the TMEM-resident parameter-gradient accumulators are exercised across tiles."""
    from nerffaceediting_b200 import triplane
    from nerffaceediting_b200.ray_sampler import RaySampler
    from nerffaceediting_b200.renderer import DisentangledImportanceRenderer
    g = golden("backward")
    n, hw, res = 2, 32, 40
    c2w, k = synth.camera_sweep(n)
    with torch.no_grad():
        o, d = RaySampler()(T(c2w.numpy(), dev), T(k.numpy(), dev), res)
    opts = dict(synth.FFHQ_RENDERING_OPTIONS, depth_resolution=24, depth_resolution_importance=24, nfe_deterministic=True, nfe_precision="fp32")
    gen = torch.Generator().manual_seed(11)
    proj = [torch.randn(n, res * res, c, generator=gen).to(dev) for c in (32, 15, 1, 1)]
    grads = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("NFE_BWD_LIBRARY_GEMM", mode)
        raw = T(synth.hash_normal(77, (n, 96, hw, hw)) * np.float32(1.5) - np.float32(0.3), dev).requires_grad_(True)
        dec = torch_decoder(g, "dis.dec", "dis", 1.0, dev)
        norm, _, _ = triplane.normalize_plane(raw)
        out = DisentangledImportanceRenderer()(norm.view(n, 3, 32, hw, hw), raw.view(n, 3, 32, hw, hw), dec, o, d, opts)
        sum((a * b).sum() for a, b in zip(out, proj)).backward()
        grads[mode] = [raw.grad] + [p.grad for p in dec.parameters()]
    names = ["raw planes"] + [nm for nm, _ in dec.named_parameters()]
    for nm, a, b in zip(names, grads["0"], grads["1"]):
        assert rel_err(N(a), N(b)) < TOL, nm


def test_fused_decoder_backward_lr_multiplier_and_frozen_parameters(dev):
    """FullyConnectedLayer gains (decoder_lr_mul != 1) go through the chain rule; frozen parameters get no gradient."""
    from nerffaceediting_b200.ray_sampler import RaySampler
    from nerffaceediting_b200.renderer import DisentangledImportanceRenderer
    from nerffaceediting_b200.triplane import DisentangledOSGDecoder
    g = golden("backward")
    with torch.no_grad():
        o, d = RaySampler()(T(g["cam2world"], dev), T(g["intrinsics"], dev), 8)
    torch.manual_seed(5)
    dec = DisentangledOSGDecoder(32, {'decoder_lr_mul': 0.5, 'decoder_output_dim': 32, 'decoder_seg_dim': 15}).to(dev)
    with torch.no_grad():
        for p in dec.parameters():
            p.add_(0.3 * torch.randn_like(p))
    dec.geo_net[0].bias.requires_grad_(False)
    planes = torch.randn(2, 3, 32, 16, 16, device=dev)
    norm = torch.randn(2, 3, 32, 16, 16, device=dev)
    opts = dict(BASE, nfe_deterministic=True, nfe_precision="fp32")
    res = {}
    import os
    for mode in ("1", "0"):
        os.environ["NFE_BWD_LIBRARY_GEMM"] = mode
        try:
            for p in dec.parameters():
                p.grad = None
            a, b = norm.clone().requires_grad_(True), planes.clone().requires_grad_(True)
            out = DisentangledImportanceRenderer()(a, b, dec, o, d, opts)
            (out[0].sum() + (out[1] ** 2).sum() + out[2].sum()).backward()
            res[mode] = [a.grad, b.grad] + [p.grad for p in dec.parameters()]
        finally:
            os.environ.pop("NFE_BWD_LIBRARY_GEMM", None)
    assert res["0"][2 + 1] is None and res["1"][2 + 1] is None            # geo_net[0].bias is frozen
    for x, y in zip(res["0"], res["1"]):
        if x is not None:
            assert rel_err(N(x), N(y)) < TOL


def _chain(dev, g, single_gather, swap, n=2, hw=32, res=23, s=16, seed=21):     # 529*16 samples per item: one tile straddles the two items
    """raw -> normalize_plane [-> denormalize_plane with other statistics] -> disentangled renderer (bf16x3) -> loss;
    returns the gradients w.r.t. the raw planes, the swapped statistics and the decoder parameters."""
    from nerffaceediting_b200 import triplane
    from nerffaceediting_b200.ray_sampler import RaySampler
    from nerffaceediting_b200.renderer import DisentangledImportanceRenderer
    c2w, k = synth.camera_sweep(n)
    with torch.no_grad():
        o, d = RaySampler()(T(c2w.numpy(), dev), T(k.numpy(), dev), res)
    opts = dict(synth.FFHQ_RENDERING_OPTIONS, depth_resolution=s, depth_resolution_importance=s, nfe_deterministic=True,
                nfe_precision="bf16x3", nfe_single_gather=single_gather)
    gen = torch.Generator().manual_seed(seed)
    proj = [torch.randn(n, res * res, c, generator=gen).to(dev) for c in (32, 15, 1, 1)]
    raw = T(synth.hash_normal(78, (n, 96, hw, hw)) * np.float32(1.5) - np.float32(0.3), dev).requires_grad_(True)
    dec = torch_decoder(g, "dis.dec", "dis", 1.0, dev)
    norm, mean, std = triplane.normalize_plane(raw)
    stats = []
    if swap:
        k_items = 1 if swap == "one" else n
        stats = [(torch.randn(k_items, 96, 1, 1, generator=gen) * 0.5).to(dev).requires_grad_(True),
                 (torch.rand(k_items, 96, 1, 1, generator=gen) + 0.5).to(dev).requires_grad_(True)]
        planes = triplane.denormalize_plane(norm, stats[0], stats[1])
    else:
        planes = raw
    out = DisentangledImportanceRenderer()(norm.view(n, 3, 32, hw, hw), planes.view(n, 3, 32, hw, hw), dec, o, d, opts)
    sum((a * b).sum() for a, b in zip(out, proj)).backward()
    return [raw.grad] + [t.grad for t in stats] + [p.grad for p in dec.parameters()], [float(x.detach().abs().max()) for x in out]


@pytest.mark.parametrize("swap", [None, "all", "one"])
def test_single_gather_backward_matches_two_gather_backward(dev, swap):
    """The AFFINE backward (planes = norm*scale + shift never read; gradient through the statistics) against the
    two-gather backward on the same chain: raw-plane, swapped-statistics and decoder gradients agree."""
    g = golden("backward")
    one, _ = _chain(dev, g, True, swap)
    two, _ = _chain(dev, g, False, swap)
    names = ["raw planes"] + (["mean'", "std'"] if swap else []) + ["geo.w1", "geo.b1", "geo.w2", "geo.b2", "app.w1", "app.b1", "app.w2", "app.b2"]
    for nm, a, b in zip(names, one, two):
        assert a is not None and b is not None and a.shape == b.shape, nm
        assert rel_err(N(a), N(b)) < 2 * TOL, nm


def test_single_gather_training_chain_vs_reference_autograd(dev):
    """normalize_plane -> renderer chain with the tensor-core forward and the single-gather backward, against the golden
    gradients of the reference's autograd pushed through the analytic normalisation backward."""
    from nerffaceediting_b200 import triplane
    from nerffaceediting_b200.ray_sampler import RaySampler
    from nerffaceediting_b200.renderer import DisentangledImportanceRenderer
    g = golden("backward")
    n, hw, res = 2, 16, 8
    raw = T(synth.hash_normal(300, (n, 96, hw, hw)) * np.float32(1.5) - np.float32(0.3), dev).requires_grad_(True)
    with torch.no_grad():
        o, d = RaySampler()(T(g["cam2world"], dev), T(g["intrinsics"], dev), res)
    dec = torch_decoder(g, "dis.dec", "dis", 1.0, dev)
    norm, _, _ = triplane.normalize_plane(raw)
    out = DisentangledImportanceRenderer()(norm.view(n, 3, 32, hw, hw), raw.view(n, 3, 32, hw, hw), dec, o, d,
                                           dict(BASE, nfe_deterministic=True, nfe_precision="bf16x3"))
    _loss(g, "dis", out, dev).backward()
    x = raw.detach().cpu().double().requires_grad_(True)
    mean = x.mean(dim=(-1, -2), keepdim=True)
    std = x.var(dim=(-1, -2), keepdim=True).sqrt()
    ((x - mean) / (std + 1e-8)).backward(torch.from_numpy(g["dis.g_norm"]).double().view(n, 96, hw, hw))
    want = x.grad.float().numpy() + g["dis.g_planes"].reshape(n, 96, hw, hw)
    assert rel_err(N(raw.grad), want) < 2 * TOL
    for name, p in dec.named_parameters():
        assert rel_err(N(p.grad), g[f"dis.g_dec.{name}"]) < 2 * TOL, name
