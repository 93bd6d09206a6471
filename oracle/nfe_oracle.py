"""ctypes/numpy front-end of the CPU oracle (oracle/nfe_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Every function takes and returns numpy float32 arrays shaped like the reference's tensors
(reference file:line citations live next to the C functions).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnfe_oracle.so")

DEC_OSG, DEC_DISENTANGLED, DEC_SEGMENTATION = 0, 1, 2

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """Compile the C restatement with the committed Makefile (gcc, seconds)."""
    src = os.path.join(_HERE, "nfe_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


class _Mlp(ctypes.Structure):
    _fields_ = [("w1", _f32p), ("b1", _f32p), ("w2", _f32p), ("b2", _f32p),
                ("in_dim", ctypes.c_int), ("hidden", ctypes.c_int), ("out_dim", ctypes.c_int),
                ("wgain1", ctypes.c_float), ("bgain1", ctypes.c_float),
                ("wgain2", ctypes.c_float), ("bgain2", ctypes.c_float)]


class _Cfg(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("C", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int),
                ("s_c", ctypes.c_int), ("s_f", ctypes.c_int),
                ("color_dim", ctypes.c_int), ("seg_dim", ctypes.c_int), ("white_back", ctypes.c_int),
                ("box_warp", ctypes.c_float), ("density_noise", ctypes.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.nfo_render.restype = ctypes.c_int
        _lib.nfo_run_model.restype = ctypes.c_int
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(_f32p)


class Mlp:
    """FC(in->hidden) · Softplus · FC(hidden->out) with FullyConnectedLayer gains
    (training/networks_stylegan2.py:96-127).  Keeps the numpy arrays alive."""

    def __init__(self, w1, b1, w2, b2, wgain1=None, bgain1=1.0, wgain2=None, bgain2=1.0):
        self.w1, self.b1, self.w2, self.b2 = _f(w1), _f(b1), _f(w2), _f(b2)
        hidden, in_dim = self.w1.shape
        out_dim = self.w2.shape[0]
        wgain1 = 1.0 / np.sqrt(in_dim) if wgain1 is None else wgain1
        wgain2 = 1.0 / np.sqrt(hidden) if wgain2 is None else wgain2
        self.c = _Mlp(_p(self.w1), _p(self.b1), _p(self.w2), _p(self.b2), in_dim, hidden, out_dim,
                      float(wgain1), float(bgain1), float(wgain2), float(bgain2))
        self.out_dim = out_dim

    @classmethod
    def from_torch(cls, seq):
        """seq = Sequential(FullyConnectedLayer, Softplus, FullyConnectedLayer) (duck-typed)."""
        a, b = seq[0], seq[2]
        return cls(a.weight.detach().cpu().numpy(), a.bias.detach().cpu().numpy(),
                   b.weight.detach().cpu().numpy(), b.bias.detach().cpu().numpy(),
                   float(a.weight_gain), float(a.bias_gain), float(b.weight_gain), float(b.bias_gain))

    def ref(self):
        return ctypes.byref(self.c)


def decoder_nets(decoder):
    """(kind, net_a, net_b, color_dim, seg_dim) from a reference-style decoder module."""
    if hasattr(decoder, "geo_net") and hasattr(decoder, "app_net"):
        a, b = Mlp.from_torch(decoder.geo_net), Mlp.from_torch(decoder.app_net)
        return DEC_DISENTANGLED, a, b, b.out_dim, a.out_dim - 1
    if hasattr(decoder, "net") and hasattr(decoder, "seg_net"):
        a, b = Mlp.from_torch(decoder.net), Mlp.from_torch(decoder.seg_net)
        return DEC_SEGMENTATION, a, b, a.out_dim - 1, b.out_dim
    a = Mlp.from_torch(decoder.net)
    return DEC_OSG, a, None, a.out_dim - 1, 0


# --------------------------------------------------------------------------- plane statistics
def plane_stats(planes):
    """planes [..., H, W] -> (mean, std) with shape [..., 1, 1]   (triplane.py:56-60)."""
    x = _f(planes)
    hw = x.shape[-1] * x.shape[-2]
    slabs = x.size // hw
    mean = np.empty(slabs, np.float32)
    std = np.empty(slabs, np.float32)
    lib().nfo_plane_stats(_p(x), ctypes.c_int64(slabs), ctypes.c_int64(hw), _p(mean), _p(std))
    shp = x.shape[:-2] + (1, 1)
    return mean.reshape(shp), std.reshape(shp)


def normalize_plane(planes):
    """(norm, mean, std)   (triplane.py:61-65)."""
    x = _f(planes)
    mean, std = plane_stats(x)
    hw = x.shape[-1] * x.shape[-2]
    out = np.empty_like(x)
    lib().nfo_normalize(_p(x), _p(mean), _p(std), ctypes.c_int64(x.size // hw), ctypes.c_int64(hw), _p(out))
    return out, mean, std


def denormalize_plane(norm, mean, std):
    """norm * std + mean with the reference's broadcasting (triplane.py:66-68,98-103)."""
    x = _f(norm)
    mean = _f(np.broadcast_to(mean, x.shape[:-2] + (1, 1)))
    std = _f(np.broadcast_to(std, x.shape[:-2] + (1, 1)))
    hw = x.shape[-1] * x.shape[-2]
    slabs = x.size // hw
    out = np.empty_like(x)
    lib().nfo_denormalize(_p(x), _p(mean), _p(std), ctypes.c_int64(slabs), ctypes.c_int64(slabs),
                          ctypes.c_int64(hw), _p(out))
    return out


def set_num_threads(n):
    """OpenMP team size of the oracle (libgomp reads OMP_NUM_THREADS only once, and torchrun exports it as 1)."""
    return int(lib().nfo_set_num_threads(int(n)))


def resize_bilinear(x, size, antialias=True):
    """F.interpolate(x, size=(size, size), mode='bilinear', align_corners=False, antialias=...) on [N,C,H,W]
    (superresolution.py:282-286)."""
    x = _f(x)
    n, c, h, w = x.shape
    oh, ow = (size, size) if isinstance(size, int) else size
    out = np.empty((n, c, oh, ow), np.float32)
    lib().nfo_resize_bilinear(_p(x), ctypes.c_int64(n * c), int(h), int(w), int(oh), int(ow), int(bool(antialias)), _p(out))
    return out


# --------------------------------------------------------------------------- rays
def generate_rays(cam2world, intrinsics, resolution):
    c, k = _f(cam2world).reshape(-1, 16), _f(intrinsics).reshape(-1, 9)
    n = c.shape[0]
    o = np.empty((n, resolution * resolution, 3), np.float32)
    d = np.empty_like(o)
    lib().nfo_generate_rays(_p(c), _p(k), n, int(resolution), _p(o), _p(d))
    return o, d


def ray_limits_box(origins, dirs, box_side_length):
    o, d = _f(origins), _f(dirs)
    n = o.size // 3
    tmin = np.empty(o.shape[:-1] + (1,), np.float32)
    tmax = np.empty_like(tmin)
    lib().nfo_ray_limits_box(_p(o), _p(d), ctypes.c_int64(n), ctypes.c_float(box_side_length), _p(tmin), _p(tmax))
    return tmin, tmax


def sample_stratified(n, r, s_c, table=None, ray_start=0.0, ray_end=0.0, start_per_ray=None,
                      end_per_ray=None, disparity=False, jitter=None):
    """depths_coarse [N,R,S,1]; `table` is torch.linspace(...) evaluated on the host."""
    mode = 2 if disparity else (1 if start_per_ray is not None else 0)
    out = np.empty((n, r, s_c, 1), np.float32)
    tb = _f(table) if table is not None else None
    sp = _f(start_per_ray) if start_per_ray is not None else None
    ep = _f(end_per_ray) if end_per_ray is not None else None
    jt = _f(jitter) if jitter is not None else None
    lib().nfo_sample_stratified(ctypes.c_int64(n * r), int(s_c), mode, _p(tb), ctypes.c_double(ray_start),
                                ctypes.c_double(ray_end), _p(sp), _p(ep), _p(jt), _p(out))
    return out


# --------------------------------------------------------------------------- gather / decoders
def sample_from_planes(planes, coords, box_warp):
    """planes [N,3,C,H,W], coords [N,M,3] -> [N,3,M,C]   (renderer.py:55-65)."""
    pl, co = _f(planes), _f(coords)
    n, _, c, h, w = pl.shape
    m = co.shape[1]
    out = np.empty((n, 3, m, c), np.float32)
    lib().nfo_sample_planes(_p(pl), _p(co), n, ctypes.c_int64(m), c, h, w, ctypes.c_float(box_warp), _p(out))
    return out


def fc(x, weight, bias, wgain, bgain):
    x, w, b = _f(x), _f(weight), _f(bias)
    y = np.empty((x.shape[0], w.shape[0]), np.float32)
    lib().nfo_fc(_p(x), ctypes.c_int64(x.shape[0]), w.shape[1], w.shape[0], _p(w), _p(b),
                 ctypes.c_float(wgain), ctypes.c_float(bgain), _p(y))
    return y


def decode(kind, net_a, net_b, feat_norm, feat_denorm, color_dim, seg_dim):
    """decoder(...) stand-alone on [N,3,M,C] features -> dict(rgb, sigma[, seg])."""
    fd = _f(feat_denorm)
    fn = _f(feat_norm) if feat_norm is not None else None
    n, _, m, c = fd.shape
    rgb = np.empty((n, m, color_dim), np.float32)
    sigma = np.empty((n, m, 1), np.float32)
    seg = np.empty((n, m, seg_dim), np.float32) if seg_dim else None
    lib().nfo_decoder(kind, net_a.ref(), net_b.ref() if net_b else None, _p(fn), _p(fd), n, ctypes.c_int64(m), c,
                      color_dim, seg_dim, _p(rgb), _p(sigma), _p(seg))
    out = {"rgb": rgb, "sigma": sigma}
    if seg is not None:
        out["seg"] = seg
    return out


# --------------------------------------------------------------------------- marching / resampling
def ray_march(colors, sigma, depths, segs=None, white_back=False):
    """[N,R,S,*] inputs -> (rgb[N,R,cc], seg[N,R,cs]|None, depth[N,R,1], weights[N,R,S-1,1])."""
    col, sg, dp = _f(colors), _f(sigma), _f(depths)
    n, r, s, cc = col.shape
    sgs = _f(segs) if segs is not None else None
    cs = sgs.shape[-1] if sgs is not None else 0
    rgb = np.empty((n, r, cc), np.float32)
    seg = np.empty((n, r, cs), np.float32) if cs else None
    depth = np.empty((n, r, 1), np.float32)
    weights = np.empty((n, r, s - 1, 1), np.float32)
    lib().nfo_ray_march(_p(col), _p(sgs), _p(sg), _p(dp), ctypes.c_int64(n * r), s, cc, cs, int(bool(white_back)),
                        _p(rgb), _p(seg), _p(depth), _p(weights))
    return rgb, seg, depth, weights


def importance_resample(z_vals, weights, s_f, u, return_indices=False):
    """z_vals [Rn,S], weights [Rn,S-1], u [S_f] or [Rn,S_f] -> samples [Rn,S_f] (+ below, above)."""
    z, w, uu = _f(z_vals), _f(weights), _f(u)
    rn, s = z.shape
    out = np.empty((rn, s_f), np.float32)
    below = np.empty((rn, s_f), np.int32)
    above = np.empty((rn, s_f), np.int32)
    lib().nfo_importance_resample(_p(z), _p(w), ctypes.c_int64(rn), s, int(s_f), _p(uu), int(uu.ndim == 2),
                                  _p(out), below.ctypes.data_as(_i32p), above.ctypes.data_as(_i32p))
    return (out, below, above) if return_indices else out


def unify_order(depths_cat):
    """depths_cat [Rn,S] (coarse then fine) -> stable ascending order [Rn,S] int32."""
    d = _f(depths_cat)
    order = np.empty(d.shape, np.int32)
    lib().nfo_unify_order(_p(d), ctypes.c_int64(d.shape[0]), d.shape[1], order.ctypes.data_as(_i32p))
    return order


# --------------------------------------------------------------------------- full forward
def _cfg(kind, planes, s_c, s_f, color_dim, seg_dim, white_back, box_warp):
    _, _, c, h, w = planes.shape
    return _Cfg(kind, c, h, w, int(s_c), int(s_f), color_dim, seg_dim, int(bool(white_back)), float(box_warp), 0.0)


def render(kind, net_a, net_b, planes_norm, planes_denorm, origins, dirs, depths_coarse, u_fine,
           s_f, color_dim, seg_dim, box_warp=1.0, white_back=False, return_stages=False):
    """Full two-pass forward.  depths_coarse [N,R,S_c]; u_fine [S_f] or [N*R,S_f] (ignored if s_f==0).
    Returns (rgb[N,R,cc], seg|None, depth[N,R,1], wsum[N,R,1]) (+ dict of stage outputs)."""
    pd = _f(planes_denorm)
    pn = _f(planes_norm) if planes_norm is not None else None
    o, d, dc = _f(origins), _f(dirs), _f(depths_coarse)
    n, r, _ = o.shape
    dc = dc.reshape(n, r, -1)
    s_c = dc.shape[-1]
    cfg = _cfg(kind, pd, s_c, s_f, color_dim, seg_dim, white_back, box_warp)
    uf = _f(u_fine) if s_f > 0 else None
    rgb = np.empty((n, r, color_dim), np.float32)
    seg = np.empty((n, r, seg_dim), np.float32) if seg_dim else None
    depth = np.empty((n, r, 1), np.float32)
    wsum = np.empty((n, r, 1), np.float32)
    dfine = np.empty((n, r, max(s_f, 1)), np.float32) if return_stages else None
    wc = np.empty((n, r, s_c - 1), np.float32) if return_stages else None
    rc = lib().nfo_render(ctypes.byref(cfg), net_a.ref(), net_b.ref() if net_b else None, _p(pn), _p(pd), pd.shape[0],
                          _p(o), _p(d), n, ctypes.c_int64(r), _p(dc), _p(uf), int(uf is not None and uf.ndim == 2),
                          _p(rgb), _p(seg), _p(depth), _p(wsum), _p(dfine), _p(wc))
    if rc != 0:
        raise ValueError("nfo_render: unsupported sizes")
    if return_stages:
        return rgb, seg, depth, wsum, {"depths_fine": dfine, "weights_coarse": wc}
    return rgb, seg, depth, wsum


def run_model(kind, net_a, net_b, planes_norm, planes_denorm, coords, color_dim, seg_dim, box_warp=1.0):
    pd = _f(planes_denorm)
    pn = _f(planes_norm) if planes_norm is not None else None
    co = _f(coords)
    n, m, _ = co.shape
    cfg = _cfg(kind, pd, 2, 0, color_dim, seg_dim, False, box_warp)
    rgb = np.empty((n, m, color_dim), np.float32)
    sigma = np.empty((n, m, 1), np.float32)
    seg = np.empty((n, m, seg_dim), np.float32) if seg_dim else None
    rc = lib().nfo_run_model(ctypes.byref(cfg), net_a.ref(), net_b.ref() if net_b else None, _p(pn), _p(pd), pd.shape[0],
                             _p(co), n, ctypes.c_int64(m), _p(rgb), _p(sigma), _p(seg))
    if rc != 0:
        raise ValueError("nfo_run_model: unsupported sizes")
    out = {"rgb": rgb, "sigma": sigma}
    if seg is not None:
        out["seg"] = seg
    return out
