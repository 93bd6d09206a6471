"""Training-side consumers of the rendered maps: drop-in forms of the helpers of the reference's
training/loss.py that sit right behind the renderer in the training step.

    remap_seg(seg)                                   loss.py:28-53
    seg_cross_entropy(image_seg, labels)             torch.nn.CrossEntropyLoss()(image_seg, labels)            loss.py:276-277
    RGBuvHistBlock()(x)                              normalised RGB-uv histograms [L,3,64,64] of x [L,3,N]      loss.py:57-121
    compute_hist_dist(target_hist, input_hist)       Hellinger distance / batch                                loss.py:123-126
    compute_seg_hist_dist(extractor, img, seg)       per-label histograms under the argmax mask, label weights loss.py:128-153
    compute_whole_hist_dist(extractor, img)                                                                    loss.py:155-157

The two distances and the cross-entropy are differentiable (autograd Functions over the CUDA kernels of csrc/nfe_losses.cu);
the reference runs the per-label distance as 12 labels x batch Python iterations (~40 launches each), here it is four kernel
launches forward and two backward.  CUDA tensors only: there is no CPU path.
"""
import ctypes

import torch

from . import _lib
from .ops import _cuda_f32, _Guard, _ptr, _stream

# training/loss.py:128-141
SEG2WEIGHT = {0: 1 / 15, 1: 3 / 15, 2: 1 / 75, 4: 1 / 75, 5: 1 / 75, 7: 1 / 15, 8: 1 / 75, 9: 1 / 15, 10: 1 / 15, 12: 1 / 15, 13: 5 / 15, 14: 1 / 15}
_TABLES = {}


def _tables(device):
    key = str(device)
    if key not in _TABLES:
        _TABLES[key] = (torch.linspace(-3, 3, steps=64).to(device),                                       # RGBuvHistBlock's bin centres
                        torch.tensor(list(SEG2WEIGHT.keys()), dtype=torch.int32, device=device),
                        torch.tensor(list(SEG2WEIGHT.values()), dtype=torch.float32, device=device),
                        torch.zeros(1, dtype=torch.int32, device=device), torch.ones(1, dtype=torch.float32, device=device))
    return _TABLES[key]


def remap_seg(seg):
    """BiSeNet's 19 labels -> the generator's 15 (loss.py:50-53; the reference rewrites `seg` in place and returns it)."""
    if not seg.is_cuda:
        raise RuntimeError("remap_seg: expected a CUDA tensor (this path has no CPU fallback)")
    src = seg.contiguous().long()
    out = torch.empty_like(src)
    with _Guard(src):
        _lib.check(_lib.load().nfe_remap_seg(_ptr(src), src.numel(), _ptr(out), _stream(src)), "nfe_remap_seg")
    if seg.dtype == torch.int64 and seg.is_contiguous():
        seg.copy_(out)
        return seg
    return out.to(seg.dtype)


class _SegCrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        x = _cuda_f32(logits, "image_seg")
        n, c = x.shape[0], x.shape[1]
        hw = x.numel() // max(n * c, 1)
        lab = labels.contiguous().long()
        if lab.numel() != n * hw:
            raise RuntimeError(f"seg_cross_entropy: labels {tuple(labels.shape)} do not match logits {tuple(logits.shape)}")
        loss = torch.empty((), device=x.device, dtype=torch.float32)
        acc = torch.empty(1, device=x.device, dtype=torch.float64)
        with _Guard(x):
            _lib.check(_lib.load().nfe_seg_cross_entropy_fwd(_ptr(x), _ptr(lab), n, c, hw, _ptr(loss), _ptr(acc), _stream(x)), "nfe_seg_cross_entropy_fwd")
        ctx.save_for_backward(x, lab)
        ctx.dims = (n, c, hw, logits.shape)
        return loss

    @staticmethod
    def backward(ctx, g):
        x, lab = ctx.saved_tensors
        n, c, hw, shape = ctx.dims
        g_x = torch.empty_like(x)
        gl = g.contiguous().float().reshape(1)
        with _Guard(x):
            _lib.check(_lib.load().nfe_seg_cross_entropy_bwd(_ptr(x), _ptr(lab), n, c, hw, _ptr(gl), _ptr(g_x), _stream(x)), "nfe_seg_cross_entropy_bwd")
        return g_x.reshape(shape), None


def seg_cross_entropy(image_seg, labels):
    """torch.nn.CrossEntropyLoss()(image_seg [N,C,H,W], labels [N,H,W]) (loss.py:276-277), differentiable w.r.t. image_seg."""
    return _SegCrossEntropy.apply(image_seg, labels)


class _HistDist(torch.autograd.Function):
    """sum_l w_l * Hellinger(hist_l(item 0), hist_l(items 1..)) / (B-1), per label (seg given) or for the whole image."""

    @staticmethod
    def forward(ctx, img, seg, sigma):
        x = _cuda_f32(img, "gen_img")
        b = x.shape[0]
        if x.dim() < 3 or x.shape[1] != 3:
            raise RuntimeError(f"histogram distance: expected an image batch [B,3,...], got {tuple(img.shape)}")
        p = x.numel() // (b * 3)
        lin, ids, weights, id0, w1 = _tables(x.device)
        if seg is not None:
            sg = _cuda_f32(seg.detach(), "gen_seg")
            c_seg = sg.shape[1]
            if sg.shape[0] != b or sg.numel() // (b * c_seg) != p:
                raise RuntimeError(f"histogram distance: seg {tuple(seg.shape)} does not match the image {tuple(img.shape)}")
            n_labels = ids.numel()
        else:
            sg, c_seg, n_labels, ids, weights = None, 0, 1, id0, w1
        dev = x.device
        raw = torch.empty((n_labels, b, 3, 64, 64), device=dev)
        norm = torch.empty_like(raw)
        totals = torch.empty((n_labels, b), device=dev)
        s_ws = torch.empty(n_labels, device=dev)
        dist = torch.empty(n_labels, device=dev)
        loss = torch.empty((), device=dev)
        with _Guard(x):
            _lib.check(_lib.load().nfe_hist_dist_fwd(_ptr(x), _ptr(sg), _ptr(ids), _ptr(lin), b, c_seg, n_labels, p, ctypes.c_float(sigma), _ptr(weights),
                                                     _ptr(raw), _ptr(norm), _ptr(totals), _ptr(s_ws), _ptr(dist), _ptr(loss), _stream(x)), "nfe_hist_dist_fwd")
        ctx.saved = (x, sg, ids, lin, weights, raw, norm, totals, s_ws)
        ctx.dims = (b, c_seg, n_labels, p, float(sigma), img.shape)
        ctx.hists = norm
        return loss

    @staticmethod
    def backward(ctx, g):
        x, sg, ids, lin, weights, raw, norm, totals, s_ws = ctx.saved
        b, c_seg, n_labels, p, sigma, shape = ctx.dims
        g_img = torch.zeros_like(x)
        g_raw = torch.empty_like(raw)
        gl = g.contiguous().float().reshape(1)
        with _Guard(x):
            _lib.check(_lib.load().nfe_hist_dist_bwd(_ptr(x), _ptr(sg), _ptr(ids), _ptr(lin), b, c_seg, n_labels, p, ctypes.c_float(sigma), _ptr(raw),
                                                     _ptr(norm), _ptr(totals), _ptr(s_ws), _ptr(weights), _ptr(gl), _ptr(g_raw), _ptr(g_img), _stream(x)),
                       "nfe_hist_dist_bwd")
        return g_img.reshape(shape), None, None


class RGBuvHistBlock(torch.nn.Module):
    """RGB-uv histogram feature (loss.py:57-121); only the configuration the reference trains with is built:
    h=64, method='inverse-quadratic', intensity_scale=True."""

    def __init__(self, h=64, method='inverse-quadratic', sigma=0.02, intensity_scale=True):
        super().__init__()
        if h != 64 or method != 'inverse-quadratic' or not intensity_scale:
            raise NotImplementedError("RGBuvHistBlock: only h=64, method='inverse-quadratic', intensity_scale=True is built")
        self.EPS = 1e-6
        self.h = h
        self.method = method
        self.intensity_scale = intensity_scale
        self.sigma = sigma

    def forward(self, x):
        """x [L,3,N] in (-1,1) -> normalised histograms [L,3,64,64] (inference form: no autograd through the histograms
        themselves; the differentiable quantities are the distances below)."""
        xx = _cuda_f32(x.detach(), "x")
        length, _, n = xx.shape
        lin, _, _, id0, w1 = _tables(xx.device)
        dev = xx.device
        raw = torch.empty((1, length, 3, 64, 64), device=dev)
        norm = torch.empty_like(raw)
        scratch = [torch.empty((1, length), device=dev), torch.empty(1, device=dev), torch.empty(1, device=dev), torch.empty((), device=dev)]
        if n == 0:
            return torch.zeros((length, 3, 64, 64), device=dev)
        with _Guard(xx):
            _lib.check(_lib.load().nfe_hist_dist_fwd(_ptr(xx), None, _ptr(id0), _ptr(lin), length, 0, 1, n, ctypes.c_float(self.sigma), _ptr(w1), _ptr(raw),
                                                     _ptr(norm), _ptr(scratch[0]), _ptr(scratch[1]), _ptr(scratch[2]), _ptr(scratch[3]), _stream(xx)),
                       "nfe_hist_dist_fwd")
        return norm[0]


def compute_hist_dist(target_hist, input_hist):
    """(1/sqrt 2) * sqrt(sum (sqrt(target) - sqrt(input))^2) / input.shape[0] (loss.py:123-126) on given histograms: a handful of
    elementwise ops on [B,3,64,64] tensors, kept in torch for callers that hold histograms already."""
    return (1 / 2 ** 0.5) * torch.sqrt(torch.sum(torch.square(torch.sqrt(target_hist) - torch.sqrt(input_hist)))) / input_hist.shape[0]


def compute_seg_hist_dist(HistExtractor, gen_img, gen_seg):
    """Per-label histogram distance (loss.py:142-153): item 0 of the batch is the (detached) target; differentiable w.r.t. gen_img."""
    return _HistDist.apply(gen_img, gen_seg, getattr(HistExtractor, "sigma", 0.02))


def compute_whole_hist_dist(HistExtractor, gen_img):
    """Whole-image histogram distance (loss.py:155-157)."""
    return _HistDist.apply(gen_img, None, getattr(HistExtractor, "sigma", 0.02))
