"""Autograd for the render path (BASELINE config 4: training-step forward + backward).

Forward is the same fused CUDA path as inference; its workspace (per-sample densities and {sigma, seg, rgb}
records of both passes) is kept for the backward.  Backward, per SURVEY.md §3.3 / §7.10:

  1. nfe_composite_bwd          ray gradients -> per-sample gradients (warp per ray, CUDA)
  2. nfe_feature_mean_fwd       decoder inputs recomputed from the planes (no [N,3,M,32] tensors were saved)
  3. decoder MLP backward       on library GEMMs (torch.addmm / autograd over [M,32] features): plain GEMMs,
                                the one place this package uses cuBLAS
  4. nfe_feature_mean_bwd       scatter-add of the feature gradients into channel-last plane gradients (CUDA,
                                red.global.add.v4.f32 — atomic, hence ulp-level run-to-run noise like the
                                reference's grid_sampler_2d_backward)
  5. nfe_planes_from_channel_last   back to the reference's [N,3,32,H,W]

Gradients reach the two plane tensors and the decoder parameters; sample positions carry none (camera labels
are data and depths_fine is detached in the reference, renderer.py:198,211).
"""
import ctypes
import os

import torch
import torch.nn.functional as F

from . import _lib, ops

MAX_ROWS = 1 << 22     # decoder-backward chunk (rows of [M,32] features) to bound the GEMM intermediates


def _mlp(x, seq):
    """FC . Softplus . FC with the FullyConnectedLayer gains (networks_stylegan2.py:114-123)."""
    a, b = seq[0], seq[2]
    h = F.softplus(torch.addmm((a.bias * a.bias_gain).unsqueeze(0), x, (a.weight * a.weight_gain).t()))
    return torch.addmm((b.bias * b.bias_gain).unsqueeze(0), h, (b.weight * b.weight_gain).t())


def _decode(kind, seq_a, seq_b, fn, fd):
    """(sigma [M,1], seg [M,15]|None, rgb [M,32]) exactly as the three decoders compute them (triplane.py:178-270)."""
    if kind == ops.DEC_DISENTANGLED:
        g = _mlp(fn, seq_a)
        return g[:, :1], g[:, 1:], torch.sigmoid(_mlp(fd, seq_b)) * (1 + 2 * 0.001) - 0.001
    x = _mlp(fd, seq_a)
    rgb = torch.sigmoid(x[:, 1:]) * (1 + 2 * 0.001) - 0.001
    return x[:, :1], (_mlp(fd, seq_b) if kind == ops.DEC_SEGMENTATION else None), rgb


class RenderFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state, norm_planes, planes, scale_src, shift_src, *params):
        """state: dict(kind, seq_a, seq_b, cfg kwargs, rays, depths_coarse, u_fine[, affine_eps]).  params are the decoder
        parameters, passed only so that autograd tracks them; the kernels read them from the modules.
        scale_src / shift_src (or None): statistics with planes == norm_planes*(scale_src + affine_eps) + shift_src per
        (item, channel) — the single-gather identity: `planes` is then never read and its gradient flows through them."""
        kind = state["kind"]
        affine = None
        if scale_src is not None:
            k = scale_src.numel() // 96
            affine = ((scale_src.detach().reshape(k, 96).float() + state["affine_eps"]).contiguous(), shift_src.detach().reshape(k, 96).float().contiguous())
        denorm_cl = ops.planes_channel_last(planes) if affine is None else None
        norm_cl = ops.planes_channel_last(norm_planes) if kind == ops.DEC_DISENTANGLED else None
        cfg = ops.make_cfg(kind, norm_cl if denorm_cl is None else denorm_cl, affine=affine, **state["cfg"])
        o, d = state["rays"]
        rgb, seg, depth, wsum, minmax, st = ops.render_fwd(cfg, state["seq_a"], state["seq_b"], norm_cl, denorm_cl, o, d, state["depths_coarse"],
                                                          state["u_fine"], return_stages=True, keep_workspace=True)
        ctx.state, ctx.cfg = state, cfg
        ctx.affine = affine
        ctx.stat_shapes = None if scale_src is None else (scale_src.shape, shift_src.shape)
        ctx.saved = (norm_cl, denorm_cl, st, minmax)
        ctx.plane_shapes = (None if norm_planes is None else norm_planes.shape, planes.shape)
        ctx.mark_non_differentiable(minmax)
        if seg is None:
            return rgb, depth, wsum, minmax
        return rgb, seg, depth, wsum, minmax

    @staticmethod
    def backward(ctx, *grads):
        state, cfg = ctx.state, ctx.cfg
        kind, seq_a, seq_b = state["kind"], state["seq_a"], state["seq_b"]
        norm_cl, denorm_cl, st, minmax = ctx.saved
        has_seg = kind != ops.DEC_OSG
        if has_seg:
            g_rgb, g_seg, g_depth, g_wsum = grads[0], grads[1], grads[2], grads[3]
        else:
            g_rgb, g_seg, g_depth, g_wsum = grads[0], None, grads[1], grads[2]
        o, d = state["rays"]
        n, r, _ = o.shape
        dev = o.device
        s_c, s_f = cfg.s_c, cfg.s_f
        lib = _lib.load()
        stream = torch.cuda.current_stream(dev).cuda_stream

        def f32(t):
            return None if t is None else t.contiguous().float()
        g_rgb = f32(g_rgb) if g_rgb is not None else torch.zeros((n, r, 32), device=dev)
        g_seg, g_depth, g_wsum = f32(g_seg), f32(g_depth), f32(g_wsum)
        dc = state["depths_coarse"].reshape(n, r, s_c).contiguous()
        df = st["depths_fine"].reshape(n, r, s_f).contiguous() if s_f else None
        g_rec_c = torch.empty((n * r, s_c, 48), device=dev)
        g_rec_f = torch.empty((n * r, s_f, 48), device=dev) if s_f else None
        P = ops._ptr
        with torch.cuda.device(dev):
            # 1. compositing backward
            _lib.check(lib.nfe_composite_bwd(P(dc), P(st["sigma_c"]), P(st["rec_c"]), s_c, P(df), P(st.get("sigma_f")), P(st.get("rec_f")), s_f,
                                             n * r, cfg.seg_dim, cfg.white_back, P(g_rgb), P(g_seg), P(g_depth), P(g_wsum), P(minmax),
                                             P(g_rec_c), P(g_rec_f), stream), "nfe_composite_bwd")
            params = [p for seq in (seq_a, seq_b) if seq is not None for p in (seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias)]
            live = [i for i, p in enumerate(params) if ctx.needs_input_grad[5 + i]]
            g_params = [torch.zeros_like(p) if i in live else None for i, p in enumerate(params)]
            affine = ctx.affine
            any_cl = denorm_cl if denorm_cl is not None else norm_cl
            pb, _, h, w, _ = any_cl.shape
            g_denorm_cl = torch.zeros_like(denorm_cl) if denorm_cl is not None else None
            g_norm_cl = torch.zeros_like(norm_cl) if norm_cl is not None else None
            g_scale = g_shift = None
            if affine is not None:
                g_scale, g_shift = torch.zeros_like(affine[0]), torch.zeros_like(affine[1])
            fused = kind == ops.DEC_DISENTANGLED and (affine is not None or os.environ.get("NFE_BWD_LIBRARY_GEMM", "0") != "1")
            if fused:
                # 2-4 fused: one tcgen05 kernel per pass recomputes features and activations, back-propagates both
                # decoder nets, scatter-adds into the plane gradients and accumulates the parameter gradients
                mlp_a, mlp_b = ops.MlpRef(seq_a, dev), ops.MlpRef(seq_b, dev)
                full = [g if g is not None else torch.zeros_like(p) for g, p in zip(g_params, params)]
                for depths, s, rec, g_rec in ((dc, s_c, st["rec_c"], g_rec_c), (df, s_f, st.get("rec_f"), g_rec_f)):
                    if not s:
                        continue
                    _lib.check(lib.nfe_field_bwd(kind, P(norm_cl), P(denorm_cl), pb, h, w, ctypes.c_float(cfg.box_warp), P(o), P(d), P(depths), n, r, s,
                                                 mlp_a.ref(), mlp_b.ref(), P(rec), P(g_rec), P(g_norm_cl), P(g_denorm_cl),
                                                 *[P(t) for t in full], P(affine[0]) if affine else None, P(affine[1]) if affine else None,
                                                 affine[0].shape[0] if affine else 0, P(g_scale), P(g_shift), stream), "nfe_field_bwd")
            for depths, s, g_rec in ((dc, s_c, g_rec_c), (df, s_f, g_rec_f)):
                if not s or fused:
                    continue
                total = n * r * s
                geom = (pb, h, w, ctypes.c_float(cfg.box_warp), P(o), P(d), P(depths), n, r, s)
                # 2. decoder inputs
                fd = torch.empty((total, 32), device=dev)
                _lib.check(lib.nfe_feature_mean_fwd(P(denorm_cl), *geom, P(fd), stream), "nfe_feature_mean_fwd")
                fn = None
                if norm_cl is not None:
                    fn = torch.empty((total, 32), device=dev)
                    _lib.check(lib.nfe_feature_mean_fwd(P(norm_cl), *geom, P(fn), stream), "nfe_feature_mean_fwd")
                g_flat = g_rec.reshape(total, 48)
                g_fd = torch.empty_like(fd)
                g_fn = torch.empty_like(fn) if fn is not None else None
                # 3. decoder backward on library GEMMs, in row chunks
                for lo in range(0, total, MAX_ROWS):
                    hi = min(total, lo + MAX_ROWS)
                    with torch.enable_grad():
                        xd = fd[lo:hi].detach().requires_grad_(True)
                        xn = fn[lo:hi].detach().requires_grad_(True) if fn is not None else None
                        sigma, seg, rgb = _decode(kind, seq_a, seq_b, xn, xd)
                        outs, gouts = [sigma, rgb], [g_flat[lo:hi, :1], g_flat[lo:hi, 16:48]]
                        if seg is not None:
                            outs.append(seg)
                            gouts.append(g_flat[lo:hi, 1:16])
                        inputs = [xd] + ([xn] if xn is not None else []) + [params[i] for i in live]
                        res = torch.autograd.grad(outs, inputs, gouts, allow_unused=True)
                    g_fd[lo:hi] = res[0]
                    k = 1
                    if xn is not None:
                        g_fn[lo:hi] = res[1]
                        k = 2
                    for i, g in zip(live, res[k:]):
                        if g is not None:
                            g_params[i] += g
                # 4. gather backward
                _lib.check(lib.nfe_feature_mean_bwd(P(g_fd), *geom, P(g_denorm_cl), stream), "nfe_feature_mean_bwd")
                if g_fn is not None:
                    _lib.check(lib.nfe_feature_mean_bwd(P(g_fn), *geom, P(g_norm_cl), stream), "nfe_feature_mean_bwd")
            # 5. back to the reference layout

            def to_ref(g_cl, shape):
                out = torch.empty(shape, device=dev)
                _lib.check(lib.nfe_planes_from_channel_last(P(g_cl), g_cl.shape[0] * 3, 32, h * w, P(out), stream), "nfe_planes_from_channel_last")
                return out
            norm_shape, plane_shape = ctx.plane_shapes
            g_planes = to_ref(g_denorm_cl, plane_shape) if (g_denorm_cl is not None and ctx.needs_input_grad[2]) else None
            g_norm = to_ref(g_norm_cl, norm_shape) if (g_norm_cl is not None and ctx.needs_input_grad[1]) else None
            if affine is not None:
                g_scale = g_scale.reshape(ctx.stat_shapes[0]) if ctx.needs_input_grad[3] else None
                g_shift = g_shift.reshape(ctx.stat_shapes[1]) if ctx.needs_input_grad[4] else None
        return (None, g_norm, g_planes, g_scale, g_shift) + tuple(g_params)


class NormalizeFunction(torch.autograd.Function):
    """normalize_plane with its analytic backward (triplane.py:56-65): n = (x - mean) / (std + 1e-8)."""

    @staticmethod
    def forward(ctx, planes):
        mean, std = ops.plane_stats(planes)
        if planes.dim() == 4 and planes.shape[1] == 96 and planes.is_contiguous() and planes.dtype == torch.float32:
            norm = ops.plane_normalize_staged(planes.detach(), mean, std)      # the raw set is staged on demand (single-gather identity: never)
        else:
            norm = ops.plane_normalize(planes, mean, std)
        ctx.save_for_backward(norm, std)
        ctx.hw = planes.shape[-1] * planes.shape[-2]
        return norm, mean, std

    @staticmethod
    def backward(ctx, g_norm, g_mean, g_std):
        norm, std = ctx.saved_tensors
        # d n_j / d x_i = (delta_ij - 1/N)/d - n_j * n_i * d / ((N-1) * std * d)   (std is the unbiased one), i.e.
        #   gx = (g - mean(g))/d - norm*sum(g*norm)/((N-1)*std) + g_mean/N + g_std*norm*d/((N-1)*std)
        # in two streaming CUDA passes (csrc/nfe_planes.cu)
        def f32(t):
            return None if t is None else t.contiguous().float()
        g, gm, gs = f32(g_norm), f32(g_mean), f32(g_std)
        slabs = norm.numel() // ctx.hw
        gx = torch.empty_like(norm)
        sums = torch.empty((slabs, 2), device=norm.device, dtype=torch.float64) if g is not None else None
        with torch.cuda.device(norm.device):
            _lib.check(_lib.load().nfe_plane_normalize_bwd(ops._ptr(g), ops._ptr(norm), ops._ptr(std), ops._ptr(gm), ops._ptr(gs), slabs, ctx.hw,
                                                           ops._ptr(sums), ops._ptr(gx), torch.cuda.current_stream(norm.device).cuda_stream),
                       "nfe_plane_normalize_bwd")
        return gx


class DenormalizeFunction(torch.autograd.Function):
    """denormalize_plane (triplane.py:66-68): out = planes * std + mean, statistics broadcast over H, W (and over the
    batch when they belong to a single item)."""

    @staticmethod
    def forward(ctx, planes, mean, std):
        ctx.save_for_backward(planes, std)
        ctx.stat_shape = mean.shape
        return ops.plane_denormalize(planes, mean, std)

    @staticmethod
    def backward(ctx, g):
        planes, std = ctx.saved_tensors

        def reduce_to(t, shape):
            t = t.sum(dim=(-1, -2), keepdim=True)
            if shape[0] == 1 and t.shape[0] != 1:
                t = t.sum(dim=0, keepdim=True)
            return t.reshape(shape)
        return (g * std if ctx.needs_input_grad[0] else None,
                reduce_to(g, ctx.stat_shape) if ctx.needs_input_grad[1] else None,
                reduce_to(g * planes, ctx.stat_shape) if ctx.needs_input_grad[2] else None)
