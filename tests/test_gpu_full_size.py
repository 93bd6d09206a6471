"""Parity at BASELINE.json's FULL sizes (VERDICT r01 weak #1b-d, #3), against fixtures the unmodified reference produced
(tests/golden/make_golden_r02.py):

* configs[2]: batch 32, 256^2 planes, 128^2 rays, appearance statistics rolled by one item, 48+48 — every 1021st ray vs golden;
* configs[3]: 256^2 planes, batch 4, 64^2 rays, 48+48, forward + backward — loss, maps, decoder gradients, and the two 100 MB
  plane gradients through sums / absolute sums per (item, channel), random projections and their largest entries
  (every CTA of field_bwd_kernel walks ~80 tiles here, so the TMEM-resident weight-gradient accumulators are exercised);
* the true drop-in flow: a generator that normalises with its OWN torch ops (training/triplane.py:61-68, what an unpickled
  TriPlaneGenerator does) feeding the module-path shadow renderer, with the path the call took asserted;
* the reference's own stochastic mode, statistically over 64 seeds per side.

Tolerance: 1e-4 relative in the max-norm (`rel_err`), the bar of BASELINE.json, AND element-wise with an absolute floor of
1 % of the tensor's magnitude (`elem_err`), the strict reading; both are asserted where maps are compared.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import synth_inputs as synth
from _util import elem_err, golden, rel_err
from test_gpu_parity import N, T, torch_decoder

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-4
ELEM_TOL = 5e-4      # element-wise, elements >= 1 % of the tensor's max (smaller ones are held to 5e-6 of the max, absolute)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _planes(seed, n, dev):
    """[n,96,256,256] regenerated item by item (the splitmix hash takes ~40 B of host memory per element while it runs)."""
    out = torch.empty((n, 96, 256, 256), device=dev)
    per = 96 * 256 * 256
    for i in range(n):
        out[i] = T(synth.hash_normal(seed, (96, 256, 256), offset=i * per) * np.float32(1.5) - np.float32(0.3), dev)
    return out


def _check_maps(got, want, tag):
    assert rel_err(got, want) < TOL, (tag, rel_err(got, want))
    assert elem_err(got, want) < ELEM_TOL, (tag, elem_err(got, want))


# ---------------------------------------------------------------------------------------------- configs[2]
@pytest.mark.parametrize("flow", ["package", "foreign"])
def test_config3_batch32_statistics_swap_vs_reference(dev, flow):
    """flow 'package': normalize_plane / denormalize_plane of this package (single-gather identity after the swap);
    flow 'foreign': the same planes made by plain torch ops (no provenance: both plane sets staged and gathered)."""
    from nerffaceediting_b200 import ops, triplane
    from nerffaceediting_b200.ray_sampler import RaySampler
    from nerffaceediting_b200.renderer import DisentangledImportanceRenderer
    g = golden("render_c3")
    n, hw, res, stride = 32, 256, int(g["res"]), int(g["ray_stride"])
    raw = _planes(int(g["planes_seed"]), n, dev)
    dec = torch_decoder(g, "dec", "dis", 1.0, dev)
    opts = dict(synth.FFHQ_RENDERING_OPTIONS, nfe_deterministic=True)
    ops.path_counts(reset=True)
    with torch.no_grad():
        o, d = RaySampler()(T(g["cam2world"], dev), T(g["intrinsics"], dev), res)
        if flow == "package":
            norm, mean, std = triplane.normalize_plane(raw)
            planes = triplane.denormalize_plane(norm, mean.roll(1, 0), std.roll(1, 0))
        else:
            mean, std = raw.mean(dim=(-1, -2), keepdim=True), raw.var(dim=(-1, -2), keepdim=True).sqrt()
            norm = (raw - mean) / (std + 1e-8)
            planes = norm * std.roll(1, 0) + mean.roll(1, 0)
        assert rel_err(N(mean).reshape(n, 96), g["mean"]) < 1e-5 and rel_err(N(std).reshape(n, 96), g["std"]) < 1e-5
        del raw
        rgb, seg, depth, wsum = DisentangledImportanceRenderer()(norm.view(n, 3, 32, hw, hw), planes.view(n, 3, 32, hw, hw), dec, o, d, opts)
    assert rgb.shape == (n, res * res, 32) and seg.shape == (n, res * res, 15)
    paths = ops.path_counts()
    assert paths.get("render:single-gather" if flow == "package" else "render:two-gather", 0) == 1, paths
    for name, t in (("rgb", rgb), ("seg", seg), ("depth", depth), ("wsum", wsum)):
        _check_maps(N(t[:, ::stride]), g[name], f"{flow}.{name}")


# ---------------------------------------------------------------------------------------------- configs[3]
def _grad_summary_check(prefix, grad, g, n):
    """grad [n,96,256,256] on the device against the fixture's summaries of the reference gradient."""
    flat = grad.reshape(n, 96, -1)
    scale = float(g[f"{prefix}.max_abs"])
    # largest entries: positions and values
    idx = torch.from_numpy(g[f"{prefix}.top_idx"]).to(grad.device)
    top = flat.reshape(-1)[idx]
    assert rel_err(N(top), g[f"{prefix}.top_val"]) < 2 * TOL, prefix
    assert abs(float(flat.abs().max()) - scale) <= 2 * TOL * scale
    # per-(item, channel) absolute sums: no cancellation, so relative per entry
    abs_sum = N(flat.abs().double().sum(-1))
    assert np.max(np.abs(abs_sum - g[f"{prefix}.abs_sum"]) / g[f"{prefix}.abs_sum"]) < 2 * TOL, prefix
    # signed sums and random projections cancel: errors are measured against the absolute mass they are sums of
    mass = g[f"{prefix}.abs_sum"]
    sums = N(flat.double().sum(-1))
    assert np.max(np.abs(sums - g[f"{prefix}.sum"]) / mass) < 2 * TOL, prefix
    host = flat.double().cpu()
    for k, seed in enumerate(range(900, 908)):
        p = torch.from_numpy(synth.hash_normal(seed, tuple(flat.shape))).double()
        got = float((host * p).sum())
        assert abs(got - float(g[f"{prefix}.proj"][k])) <= 2 * TOL * float(mass.sum()) / np.sqrt(mass.size), (prefix, k)
    # support: texels no sample touched carry exactly zero
    nnz = int(torch.count_nonzero(flat))
    assert abs(nnz - int(g[f"{prefix}.nnz"])) <= 0.001 * int(g[f"{prefix}.nnz"]), (prefix, nnz, int(g[f"{prefix}.nnz"]))


@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
def test_config4_training_step_full_size_vs_reference_autograd(dev, precision):
    from nerffaceediting_b200.ray_sampler import RaySampler
    from nerffaceediting_b200.renderer import DisentangledImportanceRenderer
    g = golden("backward_full")
    n, hw, res = 4, 256, 64
    raw = _planes(int(g["planes_seed"]), n, dev)
    mean, std = raw.mean(dim=(-1, -2), keepdim=True), raw.var(dim=(-1, -2), keepdim=True).sqrt()
    norm = ((raw - mean) / (std + 1e-8)).view(n, 3, 32, hw, hw).clone().requires_grad_(True)
    planes = raw.view(n, 3, 32, hw, hw).clone().requires_grad_(True)
    dec = torch_decoder(g, "dec", "dis", 1.0, dev)
    with torch.no_grad():
        o, d = RaySampler()(T(g["cam2world"], dev), T(g["intrinsics"], dev), res)
    gen = torch.Generator().manual_seed(int(g["proj_seed"]))
    ws = [torch.randn(n, res * res, c, generator=gen).to(dev) for c in (32, 15, 1, 1)]
    out = DisentangledImportanceRenderer()(norm, planes, dec, o, d, dict(synth.FFHQ_RENDERING_OPTIONS, nfe_deterministic=True, nfe_precision=precision))
    for name, t in zip(("rgb", "seg", "depth", "wsum"), out):
        _check_maps(N(t[:, ::64]), g[f"out.{name}"], name)
    loss = sum((a * b).sum() for a, b in zip(out, ws))
    assert abs(float(loss.detach()) - float(g["loss"])) <= TOL * max(1.0, abs(float(g["loss"])))
    loss.backward()
    for name, p in dec.named_parameters():
        assert rel_err(N(p.grad), g[f"g_dec.{name}"]) < 2 * TOL, name
    _grad_summary_check("g_norm", norm.grad.reshape(n, 96, hw, hw), g, n)
    _grad_summary_check("g_planes", planes.grad.reshape(n, 96, hw, hw), g, n)


# ---------------------------------------------------------------------------------------------- drop-in flow with foreign normalise
_DROPIN_SCRIPT = r"""
import json, sys, numpy as np, torch
root = sys.argv[1]
sys.path.insert(0, root + "/shadow")           # the module-path shadow AHEAD of (here: instead of) the reference checkout
sys.path.insert(1, root)
import synth_inputs as synth
from training.volumetric_rendering.renderer import DisentangledImportanceRenderer      # resolves to shadow/ -> nerffaceediting_b200
from training.volumetric_rendering.ray_sampler import RaySampler
from nerffaceediting_b200 import ops
from nerffaceediting_b200.triplane import DisentangledOSGDecoder

class Generator(torch.nn.Module):
    # the hot-path part of TriPlaneGenerator (training/triplane.py:56-68,84-125) with ITS OWN plane statistics in torch ops:
    # what an unpickled generator runs, because the class travels with its source inside the pickle
    def __init__(self):
        super().__init__()
        self.renderer = DisentangledImportanceRenderer()
        self.ray_sampler = RaySampler()
        self.decoder = DisentangledOSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32, 'decoder_seg_dim': 15})
        self.rendering_kwargs = dict(synth.FFHQ_RENDERING_OPTIONS, nfe_deterministic=True)
    def normalize_plane(self, planes):
        mean = torch.mean(planes, dim=(-1, -2), keepdim=True)
        var = torch.sqrt(torch.var(planes, dim=(-1, -2), keepdim=True))
        return (planes - mean) / (var + 1e-8), mean, var
    def denormalize_plane(self, planes, mean, var):
        return planes * var + mean
    def synthesis(self, planes, c, res, planes_mean=None, planes_var=None):
        cam2world, intrinsics = c[:, :16].view(-1, 4, 4), c[:, 16:25].view(-1, 3, 3)
        o, d = self.ray_sampler(cam2world, intrinsics, res)
        norm_planes, mean, var = self.normalize_plane(planes)
        if planes_mean is not None:
            planes = self.denormalize_plane(norm_planes, planes_mean, planes_var)
        norm_planes = norm_planes.view(len(norm_planes), 3, 32, norm_planes.shape[-2], norm_planes.shape[-1])
        planes = planes.view(len(planes), 3, 32, planes.shape[-2], planes.shape[-1])
        f, s, dpt, w = self.renderer(norm_planes, planes, self.decoder, o, d, self.rendering_kwargs)
        n = f.shape[0]
        return {'feature': f.permute(0, 2, 1).reshape(n, 32, res, res), 'seg': s.permute(0, 2, 1).reshape(n, 15, res, res),
                'depth': dpt.permute(0, 2, 1).reshape(n, 1, res, res), 'plane_mean': mean, 'plane_var': var}

dev = torch.device("cuda:0")
torch.manual_seed(3)
G = Generator().to(dev)
n, hw, res = 2, 64, 32
raw = torch.from_numpy(synth.hash_normal(41, (n, 96, hw, hw)) * np.float32(1.5) - np.float32(0.3)).to(dev)
c2w, k = synth.camera_sweep(n)
c = torch.cat([c2w.reshape(n, 16), k.reshape(n, 9)], 1).to(dev)
ops.path_counts(reset=True)
with torch.no_grad():
    a = G.synthesis(raw, c, res)
    b = G.synthesis(raw, c, res, planes_mean=a['plane_mean'].roll(1, 0), planes_var=a['plane_var'].roll(1, 0))
paths = ops.path_counts()
np.savez(sys.argv[2], fa=a['feature'].cpu().numpy(), sa=a['seg'].cpu().numpy(), da=a['depth'].cpu().numpy(),
         fb=b['feature'].cpu().numpy(), sb=b['seg'].cpu().numpy(), db=b['depth'].cpu().numpy(),
         **{"dec." + k_: v.cpu().numpy() for k_, v in G.decoder.state_dict().items()})
print(json.dumps({"paths": paths, "renderer_module": type(G.renderer).__module__}))
"""


def test_drop_in_flow_with_foreign_normalise_through_the_shadow(tmp_path, dev):
    """An unpatched generator: its own torch normalize / denormalize, the renderer and ray sampler resolved through shadow/.
    The renderer has no provenance for these planes, so it must take the two-gather path — and still match the oracle."""
    import json
    from oracle import nfe_oracle as orc
    from _util import oracle_decoder
    out = str(tmp_path / "dropin.npz")
    r = subprocess.run([sys.executable, "-c", _DROPIN_SCRIPT, ROOT, out], capture_output=True, text=True, timeout=600,
                       env={k: v for k, v in os.environ.items() if k != "PYTHONPATH"})
    assert r.returncode == 0, r.stderr[-2000:]
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["renderer_module"] == "training.volumetric_rendering.renderer"
    assert info["paths"].get("render:two-gather") == 2 and "render:single-gather" not in info["paths"], info
    z = np.load(out)
    n, hw, res = 2, 64, 32
    raw = synth.hash_normal(41, (n, 96, hw, hw)) * np.float32(1.5) - np.float32(0.3)
    kind, a, b, cd, sd = oracle_decoder({k: z[k] for k in z.files}, "dec", "dis")
    on, om, osd = orc.normalize_plane(raw)
    c2w, k = synth.camera_sweep(n)
    oo, od = orc.generate_rays(c2w.numpy(), k.numpy(), res)
    opts = synth.FFHQ_RENDERING_OPTIONS
    table = torch.linspace(opts['ray_start'], opts['ray_end'], 48).numpy()
    dc = orc.sample_stratified(n, res * res, 48, table=table, ray_start=opts['ray_start'], ray_end=opts['ray_end'])
    u = torch.linspace(0, 1, 48).numpy()
    for tag, planes in (("a", raw), ("b", orc.denormalize_plane(on, np.roll(om, 1, 0), np.roll(osd, 1, 0)))):
        rgb, seg, depth, _ = orc.render(kind, a, b, on.reshape(n, 3, 32, hw, hw), planes.reshape(n, 3, 32, hw, hw), oo, od, dc, u, 48, cd, sd)
        _check_maps(z["f" + tag].reshape(n, 32, -1).transpose(0, 2, 1), rgb, tag + ".feature")
        _check_maps(z["s" + tag].reshape(n, 15, -1).transpose(0, 2, 1), seg, tag + ".seg")
        _check_maps(z["d" + tag].reshape(n, 1, -1).transpose(0, 2, 1), depth, tag + ".depth")


# ---------------------------------------------------------------------------------------------- the reference's stochastic mode
def test_stochastic_mode_statistics_vs_reference(dev):
    """The reference only has stochastic sampling (renderer.py:180-190,210-211).  64 Philox-seeded renders here against 64
    torch.rand-seeded renders of the unmodified reference (fixture): per-ray means agree within the standard error of the
    difference, and the spreads agree."""
    from nerffaceediting_b200 import triplane
    from nerffaceediting_b200.ray_sampler import RaySampler
    from nerffaceediting_b200.renderer import DisentangledImportanceRenderer
    g = golden("stochastic")
    n, hw, res, seeds = 1, 64, 16, int(g["seeds"])
    raw = T(synth.hash_normal(int(g["planes_seed"]), (n, 96, hw, hw)) * np.float32(1.5) - np.float32(0.3), dev)
    dec = torch_decoder(g, "dec", "dis", 1.0, dev)
    opts = dict(synth.FFHQ_RENDERING_OPTIONS, depth_resolution=24, depth_resolution_importance=24)      # stochastic: the default
    acc = None
    with torch.no_grad():
        o, d = RaySampler()(T(g["cam2world"], dev), T(g["intrinsics"], dev), res)
        norm, _, _ = triplane.normalize_plane(raw)
        first = None
        for s in range(seeds):
            torch.manual_seed(5000 + s)
            out = DisentangledImportanceRenderer()(norm.view(n, 3, 32, hw, hw), raw.view(n, 3, 32, hw, hw), dec, o, d, opts)
            st = torch.cat(out, dim=-1).double()
            if first is None:
                first = st
            elif s == 1:
                assert not torch.equal(first, st)                      # another seed, another jitter
            acc = [st, st * st] if acc is None else [acc[0] + st, acc[1] + st * st]
        torch.manual_seed(5000)
        again = torch.cat(DisentangledImportanceRenderer()(norm.view(n, 3, 32, hw, hw), raw.view(n, 3, 32, hw, hw), dec, o, d, opts), dim=-1).double()
        assert torch.equal(again, first)                               # torch.manual_seed makes a stochastic render reproducible
    mean = (acc[0] / seeds).cpu().numpy()
    std = (acc[1] / seeds - (acc[0] / seeds) ** 2).clamp_min(0).sqrt().cpu().numpy()
    ref_mean, ref_std = g["mean"].astype(np.float64), g["std"].astype(np.float64)
    se = np.sqrt((std ** 2 + ref_std ** 2) / seeds) + 1e-6 * np.abs(ref_mean).max()
    zscore = np.abs(mean - ref_mean) / se
    # 1 x 256 x 49 comparisons: a handful beyond 4 sigma would already be suspicious, none beyond 6
    assert zscore.max() < 6.0, zscore.max()
    assert (zscore > 4.0).mean() < 1e-3
    assert abs(np.mean(zscore ** 2) - 1.0) < 0.25                       # differences are noise-sized, not systematically small or large
    ratio = (std.mean(axis=(0, 1)) + 1e-9) / (ref_std.mean(axis=(0, 1)) + 1e-9)
    assert np.all(np.abs(ratio - 1) < 0.1), ratio
