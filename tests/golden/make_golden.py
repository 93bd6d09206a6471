"""Generates tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors for this path, so these fixtures — the
reference's own outputs on seeded inputs — are what pins the oracle.  Deterministic sampling does not
exist in the reference API (renderer.py:180-190,210-211); it is imposed from outside exactly as
BASELINE.md §3 describes: `torch.rand_like` returns zeros while `sample_stratified` runs, and
`sample_pdf` is called with det=True.  Large planes are regenerated from `synth.hash_normal(seed)`
instead of being stored.
"""
import contextlib
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("NFE_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import synth_inputs as synth  # noqa: E402
from training.triplane import DisentangledOSGDecoder, OSGDecoder, SegmentationOSGDecoder  # noqa: E402
from training.volumetric_rendering import math_utils as ref_math  # noqa: E402
from training.volumetric_rendering.ray_marcher import MipRayMarcher2, SegMipRayMarcher2  # noqa: E402
from training.volumetric_rendering.ray_sampler import RaySampler  # noqa: E402
from training.volumetric_rendering.renderer import (DisentangledImportanceRenderer, ImportanceRenderer,  # noqa: E402
                                                    generate_planes, sample_from_planes)

torch.manual_seed(0)
torch.set_num_threads(8)
T = torch.from_numpy


@contextlib.contextmanager
def zero_jitter():
    orig = torch.rand_like
    torch.rand_like = lambda x, *a, **k: torch.zeros_like(x)
    try:
        yield
    finally:
        torch.rand_like = orig


class Stages:
    """Captures what the reference never returns: coarse weights, fine depths, searchsorted indices."""

    def __init__(self):
        self.depths_fine = None
        self.weights_coarse = None
        self.inds = None


def deterministic(cls, stages=None):
    class Det(cls):
        def sample_stratified(self, *a, **k):
            with zero_jitter():
                return super().sample_stratified(*a, **k)

        def sample_importance(self, z_vals, weights, n_importance):
            out = super().sample_importance(z_vals, weights, n_importance)
            if stages is not None:
                stages.weights_coarse = weights.detach().clone()
                stages.depths_fine = out.detach().clone()
            return out

        def sample_pdf(self, bins, weights, n_importance, det=False, eps=1e-5):
            orig = torch.searchsorted

            def spy(*a, **k):
                r = orig(*a, **k)
                if stages is not None:
                    stages.inds = r.clone()
                return r
            torch.searchsorted = spy
            try:
                return super().sample_pdf(bins, weights, n_importance, det=True, eps=eps)
            finally:
                torch.searchsorted = orig
    return Det()


def randomize_biases(dec, scale=0.5):
    """Biases initialise to zero (networks_stylegan2.py:110); make them count."""
    with torch.no_grad():
        for name, p in dec.named_parameters():
            if name.endswith("bias"):
                p.copy_(torch.randn_like(p) * scale)
    return dec


def decoder_arrays(prefix, dec):
    out = {}
    for k, v in dec.state_dict().items():
        out[f"{prefix}.{k}"] = v.numpy().copy()
    return out


def make_decoder(kind, lr_mul=1):
    opt = {'decoder_lr_mul': lr_mul, 'decoder_output_dim': 32, 'decoder_seg_dim': 15}
    cls = {'osg': OSGDecoder, 'dis': DisentangledOSGDecoder, 'seg': SegmentationOSGDecoder}[kind]
    return randomize_biases(cls(32, opt))


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()})
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.0f} KB")


# ------------------------------------------------------------------ 1. plane statistics (triplane.py:56-68,93-107)
def gold_stats():
    planes = T(synth.hash_normal(11, (3, 96, 8, 8))) * 1.5 - 0.3
    mean = torch.mean(planes, dim=(-1, -2), keepdim=True)
    var = torch.sqrt(torch.var(planes, dim=(-1, -2), keepdim=True))
    norm = (planes - mean) / (var + 1e-8)
    swapped = norm * torch.roll(var, 1, 0) + torch.roll(mean, 1, 0)          # statistics exchange
    item0 = norm * var[0][None, ...] + mean[0][None, ...]                    # planes_mean=0, planes_var=0
    save("stats", seed=11, mean=mean, std=var, norm=norm, denorm_swapped=swapped, denorm_item0=item0)


# ------------------------------------------------------------------ 2. rays + box limits (ray_sampler.py, math_utils.py)
def gold_rays():
    c2w = synth.look_at_cam2world([math.pi / 2 - 0.4, math.pi / 2, math.pi / 2 + 0.3],
                                  [math.pi / 2 - 0.25, math.pi / 2, math.pi / 2 + 0.2])
    k = synth.fov_to_intrinsics().unsqueeze(0).repeat(3, 1, 1).clone()
    k[1, 0, 1] = 0.05          # skew
    k[2, 0, 2] = 0.45          # off-centre principal point
    k[2, 1, 1] = 3.9
    o, d = RaySampler()(c2w, k, 6)
    tmin, tmax = ref_math.get_ray_limits_box(o, d, box_side_length=1)
    # a wide-FOV camera so that some rays miss the box
    kw = synth.fov_to_intrinsics(60.0).unsqueeze(0)
    ow, dw = RaySampler()(c2w[:1], kw, 6)
    tminw, tmaxw = ref_math.get_ray_limits_box(ow, dw, box_side_length=1.6)
    save("rays", cam2world=c2w, intrinsics=k, resolution=6, origins=o, dirs=d, tmin=tmin, tmax=tmax,
         intrinsics_wide=kw, origins_wide=ow, dirs_wide=dw, tmin_wide=tminw, tmax_wide=tmaxw)


# ------------------------------------------------------------------ 3. gather (renderer.py:23-65)
def gold_gather():
    planes = T(synth.hash_normal(21, (2, 3, 32, 12, 10)))
    coords = T(synth.hash_normal(22, (2, 120, 3))) * 0.35
    coords[0, :6] = torch.tensor([[0.5, 0.5, 0.5], [-0.5, -0.5, -0.5], [0.498, -0.499, 0.0],
                                  [0.7, 0.1, 0.1], [0.1, -0.9, 0.2], [0.0, 0.0, 0.0]])
    axes = generate_planes()
    out1 = sample_from_planes(axes, planes, coords, padding_mode='zeros', box_warp=1)
    out16 = sample_from_planes(axes, planes, coords, padding_mode='zeros', box_warp=1.6)
    # known-answer plane: value = 1000*row + col, texel centres
    h, w = 12, 10
    ka = (1000 * torch.arange(h).float()[:, None] + torch.arange(w).float()[None, :]).expand(1, 3, 32, h, w).contiguous()
    save("gather", seed_planes=21, seed_coords=22, planes_shape=(2, 3, 32, 12, 10), coords=coords, out_bw1=out1, out_bw16=out16,
         ka_out=sample_from_planes(axes, ka, coords[:1], padding_mode='zeros', box_warp=1)[:, :, :, :2])


# ------------------------------------------------------------------ 4. decoders (triplane.py:167-270)
def gold_decoders():
    arrays = {}
    fn = T(synth.hash_normal(31, (2, 3, 40, 32)))
    fd = T(synth.hash_normal(32, (2, 3, 40, 32))) * 1.5 - 0.3
    fd[0, :, 0] = 30.0     # drives a hidden unit past the softplus threshold
    dirs = torch.zeros(2, 40, 3)
    for kind, lr in (('osg', 1), ('dis', 1), ('seg', 1), ('dis', 0.5)):
        dec = make_decoder(kind, lr)
        tag = f"{kind}_lr{lr}"
        arrays.update(decoder_arrays(tag, dec))
        with torch.no_grad():
            out = dec(fd, dirs) if kind == 'osg' else dec(fn, fd, dirs)
        for k, v in out.items():
            arrays[f"{tag}.out.{k}"] = v
    save("decoders", feat_norm_seed=31, feat_denorm_seed=32, feat_denorm=fd, feat_norm=fn, **arrays)


# ------------------------------------------------------------------ 5. ray marchers (ray_marcher.py)
def gold_march():
    n, r, s = 2, 6, 12
    colors = torch.rand(n, r, s, 32)
    segs = torch.randn(n, r, s, 15) * 3
    sigma = torch.randn(n, r, s, 1) * 4 + 1
    sigma[0, 0] = -60.0                                  # zero-weight ray -> NaN depth -> global max
    sigma[0, 1] = 80.0                                   # opaque at the first interval
    depths = torch.sort(torch.rand(n, r, s, 1) * 1.05 + 2.25, dim=2)[0]
    depths[1, 2, 5] = depths[1, 2, 4]                    # zero-width interval
    opts = {'clamp_mode': 'softplus', 'white_back': False}
    optw = {'clamp_mode': 'softplus', 'white_back': True}
    a = MipRayMarcher2()(colors, sigma, depths, opts)
    aw = MipRayMarcher2()(colors, sigma, depths, optw)
    b = SegMipRayMarcher2()(colors, segs, sigma, depths, opts)
    bw = SegMipRayMarcher2()(colors, segs, sigma, depths, optw)
    save("march", colors=colors, segs=segs, sigma=sigma, depths=depths,
         mip_rgb=a[0], mip_depth=a[1], mip_weights=a[2], mip_rgb_wb=aw[0],
         seg_rgb=b[0], seg_seg=b[1], seg_depth=b[2], seg_weights=b[3], seg_rgb_wb=bw[0], seg_seg_wb=bw[1])


# ------------------------------------------------------------------ 6. importance resampling (renderer.py:194-253)
def gold_resample():
    arrays = {}
    for tag, (rays, s, s_f) in {"a": (40, 48, 48), "b": (24, 24, 40), "c": (8, 96, 96)}.items():
        z = torch.linspace(2.25, 3.3, s).reshape(1, 1, s, 1).repeat(1, rays, 1, 1).contiguous()
        z = z + torch.rand(1, rays, s, 1) * (1.05 / (s - 1)) * (torch.arange(rays).reshape(1, rays, 1, 1) % 2)
        w = torch.rand(1, rays, s - 1, 1) ** 4
        w[0, 0] = 0.0                                     # flat pdf
        w[0, 1] = 0.0
        w[0, 1, s // 2] = 0.9                             # single spike
        w[0, 2, :3] = 0.3                                 # mass in the unreachable first bins
        st = Stages()
        ren = deterministic(ImportanceRenderer, st)
        out = ren.sample_importance(z, w, s_f)
        arrays.update({f"{tag}.z": z, f"{tag}.w": w, f"{tag}.s_f": s_f, f"{tag}.out": out,
                       f"{tag}.inds": st.inds, f"{tag}.u": torch.linspace(0, 1, s_f)})
    save("resample", **arrays)


# ------------------------------------------------------------------ 7. merge (renderer.py:150-167,288-300)
def gold_unify():
    n, r, s1, s2 = 1, 5, 10, 7
    d1 = torch.sort(torch.rand(n, r, s1, 1), dim=2)[0]
    d2 = torch.rand(n, r, s2, 1)                          # unsorted, as in stochastic mode
    c1, c2 = torch.rand(n, r, s1, 32), torch.rand(n, r, s2, 32)
    g1, g2 = torch.rand(n, r, s1, 15), torch.rand(n, r, s2, 15)
    s_1, s_2 = torch.randn(n, r, s1, 1), torch.randn(n, r, s2, 1)
    a = ImportanceRenderer().unify_samples(d1, c1, s_1, d2, c2, s_2)
    b = DisentangledImportanceRenderer().unify_samples(d1, c1, g1, s_1, d2, c2, g2, s_2)
    save("unify", d1=d1, d2=d2, c1=c1, c2=c2, g1=g1, g2=g2, s1=s_1, s2=s_2,
         a_depths=a[0], a_colors=a[1], a_sigma=a[2], b_depths=b[0], b_colors=b[1], b_segs=b[2], b_sigma=b[3])


# ------------------------------------------------------------------ 8. full forward (renderer.py:88-148,301-363)
def run_forward(kind, planes_seed, plane_shape, cams, res, opts, ray_stride=1, lr=1, poses=None):
    n = plane_shape[0]
    raw = T(synth.hash_normal(planes_seed, plane_shape)) * 1.5 - 0.3        # [N,96,H,W] "backbone output"
    if poses is not None:
        # one identity under several poses (utils.py:78-80): the reference needs the plane batch to match the ray batch
        assert n == 1
        n = poses
        raw = raw.expand(n, -1, -1, -1).contiguous()
    c2w, k = cams
    o, d = RaySampler()(c2w, k, res)
    o, d = o[:, ::ray_stride].contiguous(), d[:, ::ray_stride].contiguous()
    dec = make_decoder(kind, lr)
    st = Stages()
    with torch.no_grad():
        if kind == 'osg':
            planes = raw.view(n, 3, 32, plane_shape[-2], plane_shape[-1])
            rgb, depth, wsum = deterministic(ImportanceRenderer, st)(planes, dec, o, d, opts)
            seg = None
        else:
            mean = torch.mean(raw, dim=(-1, -2), keepdim=True)
            var = torch.sqrt(torch.var(raw, dim=(-1, -2), keepdim=True))
            norm = ((raw - mean) / (var + 1e-8)).view(n, 3, 32, plane_shape[-2], plane_shape[-1])
            planes = raw.view(n, 3, 32, plane_shape[-2], plane_shape[-1])
            rgb, seg, depth, wsum = deterministic(DisentangledImportanceRenderer, st)(norm, planes, dec, o, d, opts)
    out = {"origins": o, "dirs": d, "rgb": rgb, "depth": depth, "wsum": wsum}
    if seg is not None:
        out["seg"] = seg
    if st.depths_fine is not None:
        out["depths_fine"] = st.depths_fine
        out["weights_coarse"] = st.weights_coarse
    out.update(decoder_arrays("dec", dec))
    return out


def gold_render():
    base = dict(synth.FFHQ_RENDERING_OPTIONS)
    cams2 = synth.camera_sweep(2)
    cases = {
        "osg_12_12": ('osg', dict(base, depth_resolution=12, depth_resolution_importance=12)),
        "dis_12_12": ('dis', dict(base, depth_resolution=12, depth_resolution_importance=12)),
        "seg_12_12": ('seg', dict(base, depth_resolution=12, depth_resolution_importance=12)),
        "dis_48_48": ('dis', dict(base)),
        "dis_white_bw16": ('dis', dict(base, depth_resolution=16, depth_resolution_importance=16, white_back=True, box_warp=1.6)),
        "dis_disparity": ('dis', dict(base, depth_resolution=16, depth_resolution_importance=8, disparity_space_sampling=True)),
        "dis_single_pass": ('dis', dict(base, depth_resolution=24, depth_resolution_importance=0)),
        "osg_auto": ('osg', dict(base, depth_resolution=12, depth_resolution_importance=12, ray_start='auto', ray_end='auto')),
        "dis_lr05": ('dis', dict(base, depth_resolution=12, depth_resolution_importance=12)),
    }
    arrays = {}
    for i, (tag, (kind, opts)) in enumerate(cases.items()):
        out = run_forward(kind, 100 + i, (2, 96, 16, 16), cams2, 8, opts, lr=0.5 if tag == "dis_lr05" else 1)
        arrays.update({f"{tag}.{k}": v for k, v in out.items()})
        arrays[f"{tag}.seed"] = 100 + i
    save("render_small", cam2world=cams2[0], intrinsics=cams2[1], **arrays)

    # config 1 / 2 shapes: 3x32x256x256 planes, 64^2 image, 48+48; every 16th ray kept (256 rays)
    cams1 = synth.camera_sweep(1)
    arrays = {}
    for tag, kind, seed in (("c1_osg", 'osg', 201), ("c2_dis", 'dis', 202)):
        out = run_forward(kind, seed, (1, 96, 256, 256), cams1, 64, dict(base), ray_stride=16)
        arrays.update({f"{tag}.{k}": v for k, v in out.items()})
        arrays[f"{tag}.seed"] = seed
    save("render_full", cam2world=cams1[0], intrinsics=cams1[1], **arrays)


def gold_video_sweep():
    """BASELINE configs[4] shape: ONE identity (3x32x256x256 planes) under 3 poses, 256^2 rays of which every 1021st is kept
    (65 per pose), 96+96 samples.  The fixture stores the rays and the reference outputs; planes come back from the seed."""
    cams = synth.camera_sweep(3)
    opts = dict(synth.FFHQ_RENDERING_OPTIONS, depth_resolution=96, depth_resolution_importance=96)
    out = run_forward('dis', 203, (1, 96, 256, 256), cams, 256, opts, ray_stride=1021, poses=3)
    arrays = {f"c5_dis.{k}": v for k, v in out.items()}
    arrays["c5_dis.seed"] = 203
    save("render_video_sweep", cam2world=cams[0], intrinsics=cams[1], **arrays)


# ------------------------------------------------------------------ 9. backward (config 4): reference autograd
def gold_backward():
    """Gradients of a weighted sum of ALL outputs w.r.t. both plane tensors and the decoder parameters, taken by
    the reference's own autograd graph (deterministic sampling; depths_fine is detached in the reference too,
    renderer.py:198,211)."""
    base = dict(synth.FFHQ_RENDERING_OPTIONS, depth_resolution=12, depth_resolution_importance=12)
    cams = synth.camera_sweep(2)
    arrays = {}
    for tag, kind, opts in (("dis", 'dis', base), ("osg", 'osg', base), ("dis_wb", 'dis', dict(base, white_back=True))):
        n, hw, res = 2, 16, 8
        raw = T(synth.hash_normal(300, (n, 96, hw, hw))) * 1.5 - 0.3
        o, d = RaySampler()(cams[0], cams[1], res)
        dec = make_decoder(kind, 1)
        g = torch.Generator().manual_seed(7)
        r = o.shape[1]
        wr, ws_, wd, ww = torch.randn(n, r, 32, generator=g), torch.randn(n, r, 15, generator=g), torch.randn(n, r, 1, generator=g), torch.randn(n, r, 1, generator=g)
        if kind == 'osg':
            planes = raw.view(n, 3, 32, hw, hw).clone().requires_grad_(True)
            rgb, depth, wsum = deterministic(ImportanceRenderer)(planes, dec, o, d, opts)
            loss = (rgb * wr).sum() + (depth * wd).sum() + (wsum * ww).sum()
            loss.backward()
            arrays[f"{tag}.g_planes"] = planes.grad
        else:
            mean = torch.mean(raw, dim=(-1, -2), keepdim=True)
            var = torch.sqrt(torch.var(raw, dim=(-1, -2), keepdim=True))
            norm = ((raw - mean) / (var + 1e-8)).view(n, 3, 32, hw, hw).clone().requires_grad_(True)
            planes = raw.view(n, 3, 32, hw, hw).clone().requires_grad_(True)
            rgb, seg, depth, wsum = deterministic(DisentangledImportanceRenderer)(norm, planes, dec, o, d, opts)
            loss = (rgb * wr).sum() + (seg * ws_).sum() + (depth * wd).sum() + (wsum * ww).sum()
            loss.backward()
            arrays[f"{tag}.g_norm"] = norm.grad
            arrays[f"{tag}.g_planes"] = planes.grad
            arrays[f"{tag}.norm"] = norm.detach()
        arrays[f"{tag}.loss"] = loss.detach()
        for k_, p_ in dec.named_parameters():
            arrays[f"{tag}.g_dec.{k_}"] = p_.grad
        arrays.update(decoder_arrays(f"{tag}.dec", dec))
        arrays.update({f"{tag}.wr": wr, f"{tag}.ws": ws_, f"{tag}.wd": wd, f"{tag}.ww": ww})
    # normalize_plane backward (triplane.py:56-65) through autograd
    x = (T(synth.hash_normal(301, (2, 96, 8, 8))) * 1.5 - 0.3).requires_grad_(True)
    mean = torch.mean(x, dim=(-1, -2), keepdim=True)
    var = torch.sqrt(torch.var(x, dim=(-1, -2), keepdim=True))
    nrm = (x - mean) / (var + 1e-8)
    gw = torch.randn(2, 96, 8, 8, generator=torch.Generator().manual_seed(8))
    gm, gs = torch.randn(2, 96, 1, 1, generator=torch.Generator().manual_seed(9)), torch.randn(2, 96, 1, 1, generator=torch.Generator().manual_seed(10))
    ((nrm * gw).sum() + (mean * gm).sum() + (var * gs).sum()).backward()
    arrays.update({"norm.gw": gw, "norm.gm": gm, "norm.gs": gs, "norm.g_x": x.grad})
    save("backward", cam2world=cams[0], intrinsics=cams[1], **arrays)


# ------------------------------------------------------------------ 10. SR pre-resize (superresolution.py:282-286)
def gold_resize():
    """The super-resolution module's own call on the rendered feature image: F.interpolate(..., mode='bilinear',
    align_corners=False, antialias=sr_antialias).  Cases: the shipped 64^2 -> 128^2 (both antialias settings), a 256^2
    render taken DOWN to 128^2 (config 5), and odd sizes."""
    import torch.nn.functional as F
    arrays = {}
    for tag, (n, c, h, w, oh, ow, aa) in {"up64_aa": (2, 5, 64, 64, 128, 128, True), "up64": (2, 5, 64, 64, 128, 128, False),
                                         "down256_aa": (1, 3, 256, 256, 128, 128, True), "down256": (1, 3, 256, 256, 128, 128, False),
                                         "odd_aa": (1, 2, 45, 37, 128, 128, True), "odd_down_aa": (1, 2, 301, 173, 128, 96, True),
                                         "odd_down": (1, 2, 301, 173, 128, 96, False)}.items():
        x = T(synth.hash_normal(500 + h + oh + int(aa), (n, c, h, w)))
        y = F.interpolate(x, size=(oh, ow), mode='bilinear', align_corners=False, antialias=aa)
        arrays[f"{tag}.cfg"] = np.array([n, c, h, w, oh, ow, int(aa), 500 + h + oh + int(aa)])
        arrays[f"{tag}.out"] = y[:, :, ::3, ::5].contiguous()          # a strided subset keeps the fixture small
    save("resize", **arrays)


def gold_losses():
    """SURVEY.md §8f row f4 — the training-side consumers of the rendered maps, from the reference's own training/loss.py:
    remap_seg, the segmentation cross-entropy (loss.py:276-277), RGBuvHistBlock and the per-label / whole-image histogram
    distances (loss.py:57-157).  Inputs come back from the seeds; outputs are stored (the histograms as a strided subset)."""
    from training.loss import RGBuvHistBlock, compute_seg_hist_dist, compute_whole_hist_dist, remap_seg
    b, res = 3, 32
    img = torch.tanh(T(synth.hash_normal(701, (b, 3, res, res))))                  # image_raw in (-1, 1)
    seg = T(synth.hash_normal(702, (b, 15, res, res))) * 2.0                        # image_seg logits
    seg[:, 13] += 0.8 * torch.linspace(-1, 1, res)[None, :, None]                   # hair at the top, skin in the middle: uneven masks,
    seg[:, 1] += 1.5
    seg[1, 2] -= 100.0                                                              # label 2 is empty in item 1 (zero histogram)
    seg[0, 8] -= 100.0                                                              # label 8 is empty in the target item
    labels19 = (T(synth.hash_normal(703, (b, 1, res, res))).abs() * 6.5).long().clamp(0, 18)
    hist = RGBuvHistBlock()
    with torch.no_grad():
        h_whole = hist(img.reshape(b, 3, -1))
        ce = torch.nn.CrossEntropyLoss()(seg, remap_seg(labels19.clone()).squeeze(1))
        save("losses", cfg=np.array([b, res, 701, 702, 703]), labels19=labels19.numpy().astype(np.int8),
             remapped=remap_seg(labels19.clone()).numpy().astype(np.int8), cross_entropy=ce,
             hist_whole=h_whole[:, :, ::2, ::2].contiguous(), hist_whole_sums=h_whole.sum(dim=(2, 3)),
             whole_hist_dist=compute_whole_hist_dist(hist, img), seg_hist_dist=compute_seg_hist_dist(hist, img, seg),
             argmax_counts=torch.stack([(seg.argmax(1) == i).sum(dim=(1, 2)) for i in range(15)], dim=1))


if __name__ == "__main__":
    if "--only-losses" in sys.argv:
        gold_losses()
        sys.exit(0)
    if "--only-resize" in sys.argv:
        gold_resize()
        sys.exit(0)
    if "--only-video-sweep" in sys.argv:
        gold_video_sweep()
        sys.exit(0)
    if "--only-backward" in sys.argv:
        gold_backward()
        sys.exit(0)
    gold_stats()
    gold_rays()
    gold_gather()
    gold_decoders()
    gold_march()
    gold_resample()
    gold_unify()
    gold_render()
    gold_video_sweep()
    gold_backward()
    gold_resize()
    gold_losses()
    import platform
    with open(os.path.join(HERE, "PROVENANCE.txt"), "w") as f:
        f.write(f"generated by tests/golden/make_golden.py from the reference at {REF}\n"
                f"torch {torch.__version__}, numpy {np.__version__}, {platform.processor() or platform.machine()}, "
                f"threads {torch.get_num_threads()}, seed 0\n")
