"""SURVEY.md §8f row f4 — the training-side consumers of the rendered maps (training/loss.py:28-157,276-293) on the GPU, against
the unmodified reference's outputs and autograd gradients (tests/golden/losses.npz from make_golden.py, losses_bwd.npz from
make_golden_r02.py) and against the CPU restatement (oracle/nfe_losses_oracle.py).  Tolerances: loss values 1e-5, histogram
cells 5e-5 (a cell is an fp32 sum over up to 1024 pixels whose order differs from the reference's bmm), gradients 1e-4 relative
(max-norm)."""
import numpy as np
import pytest
import torch

import synth_inputs as synth
from _util import golden, rel_err
from oracle import nfe_losses_oracle as lo
from test_gpu_parity import N, T

pytestmark = pytest.mark.gpu
B, RES = 3, 32


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _inputs(dev):
    img = np.tanh(synth.hash_normal(701, (B, 3, RES, RES)))
    seg = synth.hash_normal(702, (B, 15, RES, RES)) * np.float32(2.0)
    seg[:, 13] += np.float32(0.8) * np.linspace(-1, 1, RES, dtype=np.float32)[None, :, None]
    seg[:, 1] += np.float32(1.5)
    seg[1, 2] -= np.float32(100.0)          # label 2 is empty in item 1
    seg[0, 8] -= np.float32(100.0)          # label 8 is empty in the target item
    g = golden("losses")
    return T(img.astype(np.float32), dev), T(seg.astype(np.float32), dev), T(g["labels19"].astype(np.int64), dev), g


def test_remap_seg_and_cross_entropy(dev):
    from nerffaceediting_b200 import losses
    img, seg, labels19, g = _inputs(dev)
    gb = golden("losses_bwd")
    remapped = losses.remap_seg(labels19.clone())
    assert np.array_equal(N(remapped), g["remapped"].astype(np.int64))
    assert np.array_equal(N(remapped), lo.remap_seg(g["labels19"].astype(np.int64)))
    odd = torch.tensor([0, 18, 19, 25, -1, 7], device=dev)
    assert N(losses.remap_seg(odd.clone())).tolist() == [0, 14, 19, 25, -1, 5]              # out-of-range labels pass through
    logits = seg.clone().requires_grad_(True)
    ce = losses.seg_cross_entropy(logits, remapped.squeeze(1))
    assert abs(float(ce) - float(g["cross_entropy"])) < 1e-5 * float(g["cross_entropy"])
    assert abs(float(ce) - lo.seg_cross_entropy(N(seg), N(remapped.squeeze(1)))) < 1e-5 * float(ce)
    (ce * 1.7).backward()
    assert rel_err(N(logits.grad), gb["g_ce_logits"]) < 1e-4


def test_rgb_uv_histograms_vs_reference(dev):
    from nerffaceediting_b200 import losses
    img, seg, _, g = _inputs(dev)
    hist = losses.RGBuvHistBlock()(img.reshape(B, 3, -1))
    assert hist.shape == (B, 3, 64, 64)
    assert rel_err(N(hist)[:, :, ::2, ::2], g["hist_whole"]) < 5e-5
    assert rel_err(N(hist.sum(dim=(2, 3))), g["hist_whole_sums"]) < 1e-5
    assert abs(float(hist.sum()) - B) < 1e-4                                                # each item's histogram is normalised
    assert rel_err(N(hist), lo.rgb_uv_hist(N(img).reshape(B, 3, -1))) < 5e-5
    with pytest.raises(NotImplementedError):
        losses.RGBuvHistBlock(method='RBF')


def test_histogram_distances_forward_and_backward(dev):
    from nerffaceediting_b200 import losses
    img, seg, _, g = _inputs(dev)
    gb = golden("losses_bwd")
    ext = losses.RGBuvHistBlock()
    x = img.clone().requires_grad_(True)
    d = losses.compute_seg_hist_dist(ext, x, seg)
    assert abs(float(d) - float(g["seg_hist_dist"])) < 1e-5 * float(g["seg_hist_dist"])
    assert abs(float(d) - lo.seg_hist_dist(N(img), N(seg))) < 1e-5 * float(d)
    (d * 0.9).backward()
    assert float(x.grad[0].abs().max()) == 0.0                                             # item 0 is the detached target
    assert rel_err(N(x.grad), gb["g_seg_hist_img"]) < 1e-4
    x = img.clone().requires_grad_(True)
    d = losses.compute_whole_hist_dist(ext, x)
    assert abs(float(d) - float(g["whole_hist_dist"])) < 1e-5 * float(g["whole_hist_dist"])
    (d * 1.3).backward()
    assert rel_err(N(x.grad), gb["g_whole_hist_img"]) < 1e-4
    # colours beyond (-1, 1) are clamped: no gradient there
    x = (img * 1.6).clone().requires_grad_(True)
    d = losses.compute_whole_hist_dist(ext, x)
    assert abs(float(d) - float(gb["whole_hist_dist_clamped"])) < 1e-5 * float(gb["whole_hist_dist_clamped"])
    d.backward()
    assert rel_err(N(x.grad), gb["g_whole_hist_img_clamped"]) < 1e-4
    assert float(x.grad[x.detach().abs() > 1].abs().max()) == 0.0
    # the torch-level distance on given histograms agrees with the fused one
    h = ext(img.reshape(B, 3, -1))
    assert abs(float(losses.compute_hist_dist(h[:1], h[1:])) - float(g["whole_hist_dist"])) < 1e-5


def test_histogram_distance_at_image_size_is_deterministic_and_weighted(dev):
    """512^2-pixel images (the super-resolved output the reference also feeds, loss.py:286): two runs agree bit for bit (ordered
    compaction, no atomics in the forward), and the per-label result is the weighted sum of single-label distances."""
    from nerffaceediting_b200 import losses
    torch.manual_seed(3)
    img = torch.tanh(torch.randn(2, 3, 512, 512, device=dev))
    seg = torch.randn(2, 15, 512, 512, device=dev)
    ext = losses.RGBuvHistBlock()
    a = losses.compute_seg_hist_dist(ext, img, seg)
    b = losses.compute_seg_hist_dist(ext, img, seg)
    assert torch.equal(a, b) and torch.isfinite(a)
    lab = seg.argmax(1)
    total = 0.0
    for label, w in losses.SEG2WEIGHT.items():
        hs = torch.stack([ext(img[j][:, lab[j] == label].reshape(1, 3, -1))[0] for j in range(2)])
        total += w * float(losses.compute_hist_dist(hs[:1], hs[1:]))
    assert abs(float(a) - total) < 1e-5 * total
