"""Backbone / super-resolution plugins (first step): drop-in forms of the reference's two custom ops,

    bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda')      torch_utils/ops/bias_act.py:51-88
    upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda')              torch_utils/ops/upfirdn2d.py:117-163
    setup_filter, filter2d, upsample2d, downsample2d                                              upfirdn2d.py:67-115, 277-391

with the same arguments, defaults, autograd structure (bias_act: first AND second order, exactly the reference's two nested
Functions; upfirdn2d: the backward is the same op with up/down exchanged and the filter flipped) and error behaviour
(assertions for malformed arguments, RuntimeError for what the plugin's TORCH_CHECKs reject).  The arithmetic runs in
csrc/nfe_stylegan_ops.cu through the C ABI (nfe_bias_act, nfe_upfirdn2d); fp32, fp16 and bf16 activations.  CUDA tensors only —
`impl='ref'` is accepted for signature compatibility and runs the same kernels: this package has no PyTorch fallback.
"""
import numpy as np
import torch

from . import _lib
from .ops import _Guard, _ptr, _stream

_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


class _Act:
    def __init__(self, def_alpha, def_gain, cuda_idx, ref, has_2nd_grad):
        self.def_alpha, self.def_gain, self.cuda_idx, self.ref, self.has_2nd_grad = def_alpha, def_gain, cuda_idx, ref, has_2nd_grad


# bias_act.py:23-33 (the table the reference's layers read def_gain from, e.g. networks_stylegan2.py:109,300)
activation_funcs = {
    'linear':   _Act(0,   1,          1, '',  False),
    'relu':     _Act(0,   np.sqrt(2), 2, 'y', False),
    'lrelu':    _Act(0.2, np.sqrt(2), 3, 'y', False),
    'tanh':     _Act(0,   1,          4, 'y', True),
    'sigmoid':  _Act(0,   1,          5, 'y', True),
    'elu':      _Act(0,   1,          6, 'y', True),
    'selu':     _Act(0,   1,          7, 'y', True),
    'softplus': _Act(0,   1,          8, 'y', True),
    'swish':    _Act(0,   np.sqrt(2), 9, 'x', True),
}


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (this path has no CPU fallback)")
    if t.dtype not in _DTYPES:
        raise RuntimeError(f"{what}: dtype must be float32, float16 or bfloat16, got {t.dtype}")


def _dense(t, memory_format):
    return t.contiguous(memory_format=memory_format)


def _bias_act_call(x, b, xref, yref, dy, grad, dim, spec, alpha, gain, clamp):
    """One launch of nfe_bias_act on tensors that share ONE dense memory order (bias_act.cpp:34-99)."""
    y = torch.empty_like(x)             # preserves the memory format of a dense x
    if x.numel() == 0:
        return y
    size_b, step_b, bp = 1, 1, None
    if b is not None:
        if b.dtype != x.dtype:
            raise RuntimeError("bias_act: b must have the same dtype as x")
        if b.ndim != 1 or not 0 <= dim < x.ndim or b.shape[0] != x.shape[dim]:
            raise RuntimeError("bias_act: b must be a vector with the same number of elements as x has along dim")
        size_b, step_b, bp = b.shape[0], x.stride(dim), _ptr(b)
    with _Guard(x):
        _lib.check(_lib.load().nfe_bias_act(_ptr(x), bp, _ptr(xref) if xref is not None else None, _ptr(yref) if yref is not None else None,
                                            _ptr(dy) if dy is not None else None, _ptr(y), x.numel(), size_b, step_b, _DTYPES[x.dtype], grad,
                                            spec.cuda_idx, alpha, gain, clamp, _stream(x)), "nfe_bias_act")
    return y


_bias_act_cache = {}


def _bias_act_cuda(dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """The reference's cached pair of autograd Functions (bias_act.py:126-209): BiasActCuda saves x / b / y as the activation's
    `ref` and `has_2nd_grad` ask, its backward is BiasActCudaGrad, whose own backward gives the second-order terms."""
    assert clamp is None or clamp >= 0
    spec = activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)
    key = (dim, act, alpha, gain, clamp)
    if key in _bias_act_cache:
        return _bias_act_cache[key]

    class BiasActCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, b):
            ctx.memory_format = torch.channels_last if x.ndim == 4 and x.stride(1) == 1 else torch.contiguous_format
            x = _dense(x, ctx.memory_format)
            b = b.contiguous() if b is not None else None
            y = x
            if act != 'linear' or gain != 1 or clamp >= 0 or b is not None:
                y = _bias_act_call(x, b, None, None, None, 0, dim, spec, alpha, gain, clamp)
            keep_x = 'x' in spec.ref or spec.has_2nd_grad
            ctx.has_b = b is not None
            ctx.save_for_backward(x if keep_x else None, b if (keep_x and b is not None) else None, y if 'y' in spec.ref else None)
            return y

        @staticmethod
        def backward(ctx, dy):
            dy = _dense(dy, ctx.memory_format)
            x, b, y = ctx.saved_tensors
            dx = db = None
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                dx = dy
                if act != 'linear' or gain != 1 or clamp >= 0:
                    dx = BiasActCudaGrad.apply(dy, x, b, y)
            if ctx.has_b and ctx.needs_input_grad[1]:
                db = dx.sum([i for i in range(dx.ndim) if i != dim])
            return dx, db

    class BiasActCudaGrad(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x, b, y):
            ctx.memory_format = torch.channels_last if dy.ndim == 4 and dy.stride(1) == 1 else torch.contiguous_format
            dx = _bias_act_call(dy, b, x, y, None, 1, dim, spec, alpha, gain, clamp)
            ctx.save_for_backward(dy if spec.has_2nd_grad else None, x, b, y)
            return dx

        @staticmethod
        def backward(ctx, d_dx):
            d_dx = _dense(d_dx, ctx.memory_format)
            dy, x, b, y = ctx.saved_tensors
            d_dy = d_x = d_b = None
            if ctx.needs_input_grad[0]:
                d_dy = BiasActCudaGrad.apply(d_dx, x, b, y)
            if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
                d_x = _bias_act_call(d_dx, b, x, y, dy, 2, dim, spec, alpha, gain, clamp)
            if spec.has_2nd_grad and b is not None and ctx.needs_input_grad[2]:
                d_b = d_x.sum([i for i in range(d_x.ndim) if i != dim])
            return d_dy, d_x, d_b, None

    _bias_act_cache[key] = BiasActCuda
    return BiasActCuda


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    """Fused bias + activation + gain + clamp (bias_act.py:51-88); first and second order gradients."""
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    assert act in activation_funcs, act
    _require_cuda(x, "bias_act")
    return _bias_act_cuda(dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp).apply(x, b)


# ---------------------------------------------------------------------------------------------- upfirdn2d
def _parse_scaling(scaling):
    if isinstance(scaling, int):
        scaling = [scaling, scaling]
    assert isinstance(scaling, (list, tuple))
    assert all(isinstance(x, int) for x in scaling)
    sx, sy = scaling
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple))
    assert all(isinstance(x, int) for x in padding)
    if len(padding) == 2:
        padx, pady = padding
        padding = [padx, padx, pady, pady]
    padx0, padx1, pady0, pady1 = padding
    return padx0, padx1, pady0, pady1


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    fw, fh = int(f.shape[-1]), int(f.shape[0])
    assert fw >= 1 and fh >= 1
    return fw, fh


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """FIR filter in the form upfirdn2d() takes (upfirdn2d.py:67-115): [taps] when separable (>= 8 taps by default), else [fh, fw]."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in [0, 1, 2]
    assert f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = f.ger(f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


def _upfirdn2d_call(x, f2d, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain):
    """One launch of nfe_upfirdn2d; x in any strided layout, the result in x's memory format (upfirdn2d.cpp:20-107)."""
    n, c, in_h, in_w = x.shape
    fh, fw = f2d.shape
    up_w, up_h = in_w * upx + padx0 + padx1, in_h * upy + pady0 + pady1
    if up_w < fw or up_h < fh:
        raise RuntimeError("upfirdn2d: the up-sampled, padded image is smaller than the filter")
    out_w, out_h = (up_w - fw + downx) // downx, (up_h - fh + downy) // downy
    channels_last = x.ndim == 4 and x.stride(1) == 1 and c > 1
    y = torch.empty((n, c, out_h, out_w), dtype=x.dtype, device=x.device, memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    if y.numel() == 0:
        return y
    f2d = f2d.to(device=x.device, dtype=torch.float32).contiguous()
    i64x4 = _lib.c_i64 * 4
    with _Guard(x):
        _lib.check(_lib.load().nfe_upfirdn2d(_ptr(x), _ptr(f2d), _ptr(y), n, c, in_h, in_w, out_h, out_w, fh, fw, i64x4(*x.stride()), i64x4(*y.stride()),
                                             upx, upy, downx, downy, padx0, padx1, pady0, pady1, int(bool(flip_filter)), float(gain),
                                             _DTYPES[x.dtype], _stream(x)), "nfe_upfirdn2d")
    return y


_upfirdn2d_cache = {}


def _upfirdn2d_cuda(up=1, down=1, padding=0, flip_filter=False, gain=1):
    """The reference's cached autograd Function (upfirdn2d.py:218-273): a separable filter runs as two one-dimensional passes, the
    backward is upfirdn2d with up and down exchanged, the complementary padding and the filter flipped."""
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    key = (upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
    if key in _upfirdn2d_cache:
        return _upfirdn2d_cache[key]

    class Upfirdn2dCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, f):
            assert isinstance(x, torch.Tensor) and x.ndim == 4
            if f is None:
                f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
            if f.ndim == 1 and f.shape[0] == 1:
                f = f.square().unsqueeze(0)          # separable-1 -> full 1x1
            assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
            if f.ndim == 2:
                y = _upfirdn2d_call(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
            else:
                y = _upfirdn2d_call(x, f.unsqueeze(0), upx, 1, downx, 1, padx0, padx1, 0, 0, flip_filter, 1.0)
                y = _upfirdn2d_call(y, f.unsqueeze(1), 1, upy, 1, downy, 0, 0, pady0, pady1, flip_filter, gain)
            ctx.save_for_backward(f)
            ctx.x_shape = x.shape
            return y

        @staticmethod
        def backward(ctx, dy):
            f, = ctx.saved_tensors
            _, _, ih, iw = ctx.x_shape
            _, _, oh, ow = dy.shape
            fw, fh = _get_filter_size(f)
            p = [fw - padx0 - 1, iw * upx - ow * downx + padx0 - upx + 1, fh - pady0 - 1, ih * upy - oh * downy + pady0 - upy + 1]
            dx = None
            if ctx.needs_input_grad[0]:
                dx = _upfirdn2d_cuda(up=down, down=up, padding=p, flip_filter=(not flip_filter), gain=gain).apply(dy, f)
            assert not ctx.needs_input_grad[1]
            return dx, None

    _upfirdn2d_cache[key] = Upfirdn2dCuda
    return Upfirdn2dCuda


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Pad, up-sample, FIR-filter and down-sample a batch of 2-D images (upfirdn2d.py:117-163)."""
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    _require_cuda(x, "upfirdn2d")
    assert f is None or (isinstance(f, torch.Tensor) and f.dtype == torch.float32 and not f.requires_grad)
    return _upfirdn2d_cuda(up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain).apply(x, f)


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """upfirdn2d.py:277-311: same-size FIR filtering."""
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + fw // 2, padx1 + (fw - 1) // 2, pady0 + fh // 2, pady1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """upfirdn2d.py:315-350: the result is `up` times the input size."""
    upx, upy = _parse_scaling(up)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw + upx - 1) // 2, padx1 + (fw - upx) // 2, pady0 + (fh + upy - 1) // 2, pady1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """upfirdn2d.py:354-389: the result is the input size divided by `down`."""
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw - downx + 1) // 2, padx1 + (fw - downx) // 2, pady0 + (fh - downy + 1) // 2, pady1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)
