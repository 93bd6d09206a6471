"""Regression tests for the round-1 advisor findings and the registry hardening."""
import numpy as np
import pytest
import torch

import synth_inputs as synth
from _util import golden, rel_err
from test_gpu_parity import N, T, torch_decoder

pytestmark = pytest.mark.gpu
BASE = dict(synth.FFHQ_RENDERING_OPTIONS, depth_resolution=12, depth_resolution_importance=12, nfe_deterministic=True)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _scene(dev, n=2, hw=32, res=12):
    from nerffaceediting_b200.ray_sampler import RaySampler
    g = golden("backward")
    raw = T(synth.hash_normal(78, (n, 96, hw, hw)) * np.float32(1.5) - np.float32(0.3), dev)
    c2w, k = synth.camera_sweep(n)
    with torch.no_grad():
        o, d = RaySampler()(c2w.to(dev), k.to(dev), res)
    return g, raw, o, d


@pytest.mark.parametrize("detach", ["norm", "planes", "stats"])
def test_single_gather_training_honours_a_detached_branch(dev, detach):
    """renderer(norm.detach(), planes) / renderer(norm, planes.detach()) on a pair that has provenance: the identity must not be
    used (it would re-route or drop gradient); the result must equal the two-gather backward with the same detachments."""
    from nerffaceediting_b200 import ops, triplane
    from nerffaceediting_b200.renderer import DisentangledImportanceRenderer
    g, raw0, o, d = _scene(dev)
    n, hw = raw0.shape[0], raw0.shape[-1]
    gen = torch.Generator().manual_seed(3)
    proj = [torch.randn(n, o.shape[1], c, generator=gen).to(dev) for c in (32, 15, 1, 1)]
    grads, paths = {}, {}
    for single in (True, False):
        raw = raw0.clone().requires_grad_(True)
        dec = torch_decoder(g, "dis.dec", "dis", 1.0, dev)
        norm, mean, std = triplane.normalize_plane(raw)
        planes = raw
        if detach == "norm":
            norm = norm.detach()
        elif detach == "planes":
            planes = raw.detach()
        else:           # a swap with constant statistics: the de-normalised planes depend on the raw ones through norm only
            planes = triplane.denormalize_plane(norm, mean.detach().roll(1, 0), std.detach().roll(1, 0))
        ops.path_counts(reset=True)
        out = DisentangledImportanceRenderer()(norm.view(n, 3, 32, hw, hw), planes.view(n, 3, 32, hw, hw), dec, o, d,
                                               dict(BASE, nfe_precision="bf16x3", nfe_single_gather=single))
        paths[single] = ops.path_counts()
        sum((a * b).sum() for a, b in zip(out, proj)).backward()
        grads[single] = [raw.grad] + [p.grad for p in dec.parameters()]
    if detach in ("norm", "planes"):
        assert "render:training-single-gather" not in paths[True], paths[True]          # the identity was refused
    else:
        assert paths[True].get("render:training-single-gather") == 1, paths[True]       # constant statistics are fine
    for a, b in zip(grads[True], grads[False]):
        assert (a is None) == (b is None)
        if a is not None:
            assert rel_err(N(a), N(b)) < 2e-4


def test_cuda_graph_of_the_renderer_alone_follows_in_place_plane_updates(dev):
    """capturing ONLY the renderer call on planes that already have staging / provenance records used to bake the
    capture-time staging buffer into the graph; a replay after raw.copy_(new) then rendered the old planes."""
    from nerffaceediting_b200 import graphs, triplane
    from nerffaceediting_b200.renderer import DisentangledImportanceRenderer
    g, raw, o, d = _scene(dev)
    n, hw = raw.shape[0], raw.shape[-1]
    dec = torch_decoder(g, "dis.dec", "dis", 1.0, dev)
    ren = DisentangledImportanceRenderer()
    opts = dict(BASE, nfe_cache_planes=True)
    with torch.no_grad():
        norm, _, _ = triplane.normalize_plane(raw)          # registers staging + provenance BEFORE the capture
        ren(norm.view(n, 3, 32, hw, hw), raw.view(n, 3, 32, hw, hw), dec, o, d, opts)
    call = graphs.capture(lambda: ren(norm.view(n, 3, 32, hw, hw), raw.view(n, 3, 32, hw, hw), dec, o, d, opts))
    with torch.no_grad():
        new_raw = T(synth.hash_normal(79, (n, 96, hw, hw)) * np.float32(1.5) - np.float32(0.3), dev)
        new_norm = triplane.normalize_plane(new_raw)[0]
        raw.copy_(new_raw)
        norm.copy_(new_norm)
        replayed = [t.clone() for t in call()]
        eager = ren(new_norm.view(n, 3, 32, hw, hw), new_raw.view(n, 3, 32, hw, hw), dec, o, d, BASE)
    for a, b in zip(replayed, eager):
        assert rel_err(N(a), N(b)) < 1e-5          # (two-gather inside the graph vs single-gather eager: 2e-7 apart)


def test_staged_path_does_not_silently_drop_decoder_gradients(dev):
    """A decoder that is not one of the reference's three goes through the stage kernels, which have no backward: if its
    parameters require grad the call must raise, not return detached maps."""
    from nerffaceediting_b200.renderer import ImportanceRenderer

    class Custom(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(32, 33)

        def forward(self, feats, dirs):
            x = self.lin(feats.mean(1))
            return {"rgb": torch.sigmoid(x[..., 1:]), "sigma": x[..., :1]}
    g, raw, o, d = _scene(dev)
    dec = Custom().to(dev)
    planes = raw.view(2, 3, 32, 32, 32)
    with torch.no_grad():
        out = ImportanceRenderer()(planes, dec, o, d, BASE)
    assert out[0].shape == (2, o.shape[1], 32)
    with pytest.raises(RuntimeError, match="backward"):
        ImportanceRenderer()(planes, dec, o, d, BASE)
    with pytest.raises(RuntimeError, match="ray_origins"):
        ImportanceRenderer()(planes, dec, o.clone().requires_grad_(True), d, BASE)


def test_workspace_chunks_draw_independent_noise(dev, monkeypatch):
    """a render split into workspace chunks reused the same Philox (seed, offset, local index) in every chunk.
    Two identical batch items rendered as two chunks with density noise must not come out identical."""
    from nerffaceediting_b200.renderer import DisentangledImportanceRenderer
    g, raw, o, d = _scene(dev, n=1, hw=32, res=16)
    dec = torch_decoder(g, "dis.dec", "dis", 1.0, dev)
    planes = raw.view(1, 3, 32, 32, 32).expand(2, -1, -1, -1, -1).contiguous()
    o2, d2 = o.expand(2, -1, -1).contiguous(), d.expand(2, -1, -1).contiguous()
    opts = dict(BASE, density_noise=0.5)
    ren = DisentangledImportanceRenderer()
    with torch.no_grad():
        torch.manual_seed(0)
        whole = ren(planes, planes, dec, o2, d2, opts)
        monkeypatch.setenv("NFE_WORKSPACE_MB", "1")            # one item needs ~1.2 MB: every item (or ray block) is its own chunk
        torch.manual_seed(0)
        chunked = ren(planes, planes, dec, o2, d2, opts)
    assert not torch.equal(whole[0][0], whole[0][1])          # one launch: the items draw different noise
    assert not torch.equal(chunked[0][0], chunked[0][1])      # chunked: still different
    assert all(torch.isfinite(t).all() for t in chunked)
