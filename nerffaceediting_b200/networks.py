"""Backbone / super-resolution convolution stack, inference forward: drop-in forms of

    modulated_conv2d                                   training/networks_stylegan2.py:34-91
    FullyConnectedLayer (all activations)              networks_stylegan2.py:96-130
    MappingNetwork                                     networks_stylegan2.py:193-270
    SynthesisLayer, ToRGBLayer                         networks_stylegan2.py:276-357
    SynthesisBlock ('orig' and 'skip'), SynthesisNetwork   networks_stylegan2.py:365-526
    Generator (the StyleGAN2 backbone)                 networks_stylegan2.py:529-557
    SuperresolutionHybrid8XDC / 8X / 4X / 2X           training/superresolution.py:29-125,264-290

with the reference's constructor arguments, parameter / buffer names (state dicts load unchanged) and call signatures.  Every
convolution runs in csrc/nfe_modconv.cu through the C ABI (`nfe_modulated_conv2d`): a tcgen05 implicit GEMM over channels-last
activations with the layer's noise / bias / activation / clamp fused into its epilogue; `up=2` layers are four phase convolutions
plus one fused filter pass.  fp16 blocks use fp16 tensor-core operands like the reference's; fp32 blocks (and `force_fp32=True`)
use bf16 hi/lo split operands (three MMAs per product).  CUDA only, forward only: a call that needs gradients raises.
"""
import numpy as np
import torch

from . import _lib
from . import stylegan_ops as sg
from .ops import _Guard, _ptr, _stream, resize_bilinear

_DT = {torch.float32: 0, torch.float16: 1}
_ACT_IDX = {'linear': 1, 'relu': 2, 'lrelu': 3}


def _no_grad(*tensors):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise RuntimeError("nerffaceediting_b200.networks: the convolution stack is forward-only; call it under torch.no_grad()")


_LAYOUT_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def _channels_last(x):
    """x [N,C,H,W] in channels-last memory order; a contiguous-NCHW CUDA tensor goes through the library's tiled transpose
    (nfe_layout_convert) instead of torch's strided copy."""
    if x.is_contiguous(memory_format=torch.channels_last):
        return x
    if x.is_cuda and x.is_contiguous() and x.dtype in _LAYOUT_DT and x.shape[0] <= 65535 and x.numel() > 0:
        n, c, h, w = x.shape
        y = torch.empty_like(x, memory_format=torch.channels_last)
        with _Guard(x):
            _lib.check(_lib.load().nfe_layout_convert(_ptr(x), _ptr(y), n, c, h * w, _LAYOUT_DT[x.dtype], 1, _stream(x)), "nfe_layout_convert")
        return y
    return x.contiguous(memory_format=torch.channels_last)


def _accumulate_image(img, y):
    """networks_stylegan2.py:456-457: `y = y.to(float32, contiguous_format); img = img.add_(y) if img is not None else y`, the conversion
    and the add as one pass (nfe_image_accumulate) when y comes channels-last out of the convolution kernel."""
    if (img is not None and y.is_cuda and y.ndim == 4 and y.dtype in (torch.float32, torch.float16) and img.dtype == torch.float32
            and img.shape == y.shape and img.is_contiguous() and y.is_contiguous(memory_format=torch.channels_last)
            and not y.is_contiguous() and y.shape[0] <= 65535 and y.numel() > 0):
        n, c, h, w = y.shape
        with _Guard(y):
            _lib.check(_lib.load().nfe_image_accumulate(_ptr(y), _ptr(img), n, c, h * w, _LAYOUT_DT[y.dtype], _stream(y)), "nfe_image_accumulate")
        return img
    y = y.to(dtype=torch.float32, memory_format=torch.contiguous_format)
    return img.add_(y) if img is not None else y


def _nchw(x):
    """x [N,C,H,W] as a contiguous-NCHW tensor (the layout the reference's own code expects back)."""
    if x.is_contiguous():
        return x
    if x.is_cuda and x.is_contiguous(memory_format=torch.channels_last) and x.dtype in _LAYOUT_DT and x.shape[0] <= 65535 and x.numel() > 0:
        n, c, h, w = x.shape
        y = torch.empty(x.shape, dtype=x.dtype, device=x.device)
        with _Guard(x):
            _lib.check(_lib.load().nfe_layout_convert(_ptr(x), _ptr(y), n, c, h * w, _LAYOUT_DT[x.dtype], 0, _stream(x)), "nfe_layout_convert")
        return y
    return x.contiguous()


def _modconv(x, weight, styles, noise, up, resample_filter, demodulate, flip_weight, bias, act, gain, clamp):
    """One call of nfe_modulated_conv2d; x [N,C,H,W] (any layout; staged channels-last), result channels-last in x.dtype.
    weight [O,I,k,k] (one for the batch) or [N,O,I,k,k] (per-item weights: the grouped form of a fused modulated_conv2d);
    styles [N,I] or None (= ones)."""
    if not x.is_cuda:
        raise RuntimeError("modulated_conv2d: expected a CUDA tensor (this path has no CPU fallback)")
    if x.dtype not in _DT:
        raise RuntimeError(f"modulated_conv2d: activations must be float32 or float16, got {x.dtype}")
    n, c, h, w = x.shape
    o, ci, kh, kw = weight.shape[-4:]
    assert ci == c and kh == kw
    assert weight.ndim == 4 or (weight.ndim == 5 and weight.shape[0] == n)
    assert styles is None or styles.shape == (n, c)
    x = _channels_last(x)
    weight = weight.detach().to(torch.float32).contiguous()          # dense [O,I,k,k] whatever memory format the parameter was made in
    styles = styles.detach().to(torch.float32).contiguous() if styles is not None else None
    y = torch.empty((n, o, h * up, w * up), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    nz, nz_stride = None, 0
    if noise is not None:
        nz = noise.detach().to(device=x.device, dtype=torch.float32)
        if nz.ndim == 0:
            nz = nz.expand(h * up, w * up)
        if nz.ndim == 4:
            assert nz.shape[1] == 1 and nz.shape[0] in (1, n)
            nz_stride = h * up * w * up if nz.shape[0] == n and n > 1 else 0
        nz = nz.contiguous()
        assert nz.shape[-2:] == (h * up, w * up)
    b = bias.detach().to(torch.float32).contiguous() if bias is not None else None
    f = None
    if up == 2:
        f = resample_filter.to(device=x.device, dtype=torch.float32)
        if f.ndim == 1:
            f = f.ger(f)
        f = f.contiguous()
    args = _lib.NfeModconvArgs(x=_ptr(x), weight=_ptr(weight), styles=_ptr(styles), noise=_ptr(nz), noise_batch_stride=nz_stride, bias=_ptr(b),
                               y=_ptr(y), filter=_ptr(f), fh=0 if f is None else f.shape[0], fw=0 if f is None else f.shape[1],
                               batch=n, in_ch=c, out_ch=o, in_h=h, in_w=w, ksize=kh, up=up, demodulate=int(bool(demodulate)),
                               flip_weight=int(bool(flip_weight)), act=_ACT_IDX[act], alpha=0.2 if act == 'lrelu' else 0.0, gain=float(gain),
                               clamp=float(clamp) if clamp is not None else -1.0, dtype=_DT[x.dtype],
                               weight_batch_stride=o * ci * kh * kw if weight.ndim == 5 else 0)
    lib = _lib.load()
    with _Guard(x):
        need = lib.nfe_modconv_workspace_bytes(args)
        if need < 0:
            _lib.check(1, "nfe_modulated_conv2d")
        ws = torch.empty(max(int(need), 256), dtype=torch.uint8, device=x.device)
        _lib.check(lib.nfe_modulated_conv2d(args, _ptr(ws), ws.numel(), _stream(x)), "nfe_modulated_conv2d")
    return y


def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True, flip_weight=True,
                     fused_modconv=True):
    """networks_stylegan2.py:34-91.  `fused_modconv` selects between two algebraically equal formulations in the reference; here the
    per-sample fold always happens in the weight-packing kernel.  Supported: the reference's own call sites — 3x3 or 1x1 kernels,
    up in {1, 2}, down = 1, padding = kernel // 2."""
    _no_grad(x, weight, styles, noise)
    assert down == 1, "modulated_conv2d: down-sampling convolutions are discriminator-side and not provided"
    assert up in (1, 2)
    assert padding == weight.shape[-1] // 2, "modulated_conv2d: padding must be kernel_size // 2 (what the reference's layers pass)"
    return _modconv(x, weight, styles, noise, up, resample_filter, demodulate, flip_weight, None, 'linear', 1.0, None)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """torch_utils/ops/conv2d_resample.py:48-143 for the cases the generator's layers produce (inference): 3x3 / 1x1 kernels, up in
    {1, 2}, down = 1, padding = kernel // 2, and either groups = 1 (one weight for the batch: the non-fused formulation, Conv2dLayer)
    or the fused modulated_conv2d's grouped form — x [1, N*I, H, W], w [N*O, I, k, k], groups = N (networks_stylegan2.py:84-88), which
    runs as N items with per-item weights.  The result has the caller's layout conventions ([N, O, H', W'] / [1, N*O, H', W'])."""
    _no_grad(x, w)
    kh, kw = w.shape[-2:]
    pad = [padding] * 4 if isinstance(padding, int) else ([padding[0], padding[0], padding[1], padding[1]] if len(padding) == 2 else list(padding))
    assert down == 1 and up in (1, 2) and kh == kw and all(p == kh // 2 for p in pad) and not flip_filter, "conv2d_resample: unsupported configuration"
    if groups == 1:
        return _modconv(x, w, None, None, up, f, False, flip_weight, None, 'linear', 1.0, None)
    assert x.shape[0] == 1 and x.shape[1] % groups == 0 and w.shape[0] % groups == 0
    n, i, o = groups, x.shape[1] // groups, w.shape[0] // groups
    y = _modconv(x.reshape(n, i, *x.shape[2:]), w.reshape(n, o, i, kh, kw), None, None, up, f, False, flip_weight, None, 'linear', 1.0, None)
    return _nchw(y).reshape(1, n * o, *y.shape[2:])


def normalize_2nd_moment(x, dim=1, eps=1e-8):
    return x * (x.square().mean(dim=dim, keepdim=True) + eps).rsqrt()


class FullyConnectedLayer(torch.nn.Module):
    """networks_stylegan2.py:96-130 (matrices of at most 512 x 512 on a handful of rows: library GEMM + the bias_act kernel)."""

    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1, bias_init=0):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.activation = activation
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x):
        w = self.weight.to(x.dtype) * self.weight_gain
        b = self.bias
        if b is not None:
            b = b.to(x.dtype)
            if self.bias_gain != 1:
                b = b * self.bias_gain
        if self.activation == 'linear' and b is not None:
            return torch.addmm(b.unsqueeze(0), x, w.t())
        return sg.bias_act(x.matmul(w.t()), b, act=self.activation)

    def extra_repr(self):
        return f'in_features={self.in_features:d}, out_features={self.out_features:d}, activation={self.activation:s}'


class MappingNetwork(torch.nn.Module):
    """networks_stylegan2.py:193-270."""

    def __init__(self, z_dim, c_dim, w_dim, num_ws, num_layers=8, embed_features=None, layer_features=None, activation='lrelu',
                 lr_multiplier=0.01, w_avg_beta=0.998):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.num_ws, self.num_layers, self.w_avg_beta = z_dim, c_dim, w_dim, num_ws, num_layers, w_avg_beta
        if embed_features is None:
            embed_features = w_dim
        if c_dim == 0:
            embed_features = 0
        if layer_features is None:
            layer_features = w_dim
        features_list = [z_dim + embed_features] + [layer_features] * (num_layers - 1) + [w_dim]
        if c_dim > 0:
            self.embed = FullyConnectedLayer(c_dim, embed_features)
        for idx in range(num_layers):
            setattr(self, f'fc{idx}', FullyConnectedLayer(features_list[idx], features_list[idx + 1], activation=activation, lr_multiplier=lr_multiplier))
        if num_ws is not None and w_avg_beta is not None:
            self.register_buffer('w_avg', torch.zeros([w_dim]))

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False):
        _no_grad(z, c)
        assert not update_emas, "MappingNetwork: update_emas is a training-time feature"
        x = None
        if self.z_dim > 0:
            x = normalize_2nd_moment(z.to(torch.float32))
        if self.c_dim > 0:
            y = normalize_2nd_moment(self.embed(c.to(torch.float32)))
            x = torch.cat([x, y], dim=1) if x is not None else y
        for idx in range(self.num_layers):
            x = getattr(self, f'fc{idx}')(x)
        if self.num_ws is not None:
            x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            assert self.w_avg_beta is not None
            if self.num_ws is None or truncation_cutoff is None:
                x = self.w_avg.lerp(x, truncation_psi)
            else:
                x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x


class SynthesisLayer(torch.nn.Module):
    """networks_stylegan2.py:276-330: modulated 3x3 convolution (optionally up-sampling by 2) + noise + bias + activation + clamp,
    ONE fused call here."""

    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, use_noise=True, activation='lrelu',
                 resample_filter=[1, 3, 3, 1], conv_clamp=None, channels_last=False):
        super().__init__()
        self.in_channels, self.out_channels, self.w_dim, self.resolution, self.up = in_channels, out_channels, w_dim, resolution, up
        self.use_noise, self.activation, self.conv_clamp = use_noise, activation, conv_clamp
        self.register_buffer('resample_filter', sg.setup_filter(resample_filter))
        self.padding = kernel_size // 2
        self.act_gain = sg.activation_funcs[activation].def_gain
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        if use_noise:
            self.register_buffer('noise_const', torch.randn([resolution, resolution]))
            self.noise_strength = torch.nn.Parameter(torch.zeros([]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))

    def forward(self, x, w, noise_mode='random', fused_modconv=True, gain=1):
        assert noise_mode in ['random', 'const', 'none']
        _no_grad(x, w, self.weight)
        in_resolution = self.resolution // self.up
        assert x.shape[1:] == (self.in_channels, in_resolution, in_resolution)
        styles = self.affine(w)
        noise = None
        if self.use_noise and noise_mode == 'random':
            noise = torch.randn([x.shape[0], 1, self.resolution, self.resolution], device=x.device) * self.noise_strength
        if self.use_noise and noise_mode == 'const':
            noise = self.noise_const * self.noise_strength
        act_gain = self.act_gain * gain
        act_clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        if self.activation in _ACT_IDX:
            return _modconv(x, self.weight, styles, noise, self.up, self.resample_filter, True, self.up == 1, self.bias, self.activation, act_gain, act_clamp)
        x = _modconv(x, self.weight, styles, noise, self.up, self.resample_filter, True, self.up == 1, None, 'linear', 1.0, None)
        return sg.bias_act(x, self.bias.to(x.dtype), act=self.activation, gain=act_gain, clamp=act_clamp)

    def extra_repr(self):
        return ' '.join([f'in_channels={self.in_channels:d}, out_channels={self.out_channels:d}, w_dim={self.w_dim:d},',
                         f'resolution={self.resolution:d}, up={self.up}, activation={self.activation:s}'])


class ToRGBLayer(torch.nn.Module):
    """networks_stylegan2.py:338-357: modulated 1x1 convolution without demodulation + bias + clamp."""

    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1, conv_clamp=None, channels_last=False):
        super().__init__()
        self.in_channels, self.out_channels, self.w_dim, self.conv_clamp = in_channels, out_channels, w_dim, conv_clamp
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))

    def forward(self, x, w, fused_modconv=True):
        _no_grad(x, w, self.weight)
        styles = self.affine(w) * self.weight_gain
        return _modconv(x, self.weight, styles, None, 1, None, False, True, self.bias, 'linear', 1.0, self.conv_clamp)

    def extra_repr(self):
        return f'in_channels={self.in_channels:d}, out_channels={self.out_channels:d}, w_dim={self.w_dim:d}'


class SynthesisBlock(torch.nn.Module):
    """networks_stylegan2.py:365-464, architectures 'orig' and 'skip' (the generator's and the super-resolution heads').  Activations
    stay channels-last between the layers whatever `fp16_channels_last` says: that is the layout the convolution kernel reads."""
    _up = 2          # SynthesisBlockNoUp (superresolution.py:158-253) is the same block with 1

    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture='skip', resample_filter=[1, 3, 3, 1],
                 conv_clamp=256, use_fp16=False, fp16_channels_last=False, fused_modconv_default=True, **layer_kwargs):
        assert architecture in ['orig', 'skip'], "SynthesisBlock: the 'resnet' architecture is not used by the reference's generator"
        super().__init__()
        self.in_channels, self.w_dim, self.resolution, self.img_channels, self.is_last = in_channels, w_dim, resolution, img_channels, is_last
        self.architecture, self.use_fp16, self.channels_last, self.fused_modconv_default = architecture, use_fp16, (use_fp16 and fp16_channels_last), fused_modconv_default
        self.register_buffer('resample_filter', sg.setup_filter(resample_filter))
        self.num_conv = 0
        self.num_torgb = 0
        if in_channels == 0:
            self.const = torch.nn.Parameter(torch.randn([out_channels, resolution, resolution]))
        if in_channels != 0:
            up_kwargs = dict(up=2, resample_filter=resample_filter) if self._up == 2 else {}
            self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim=w_dim, resolution=resolution, conv_clamp=conv_clamp,
                                        channels_last=self.channels_last, **up_kwargs, **layer_kwargs)
            self.num_conv += 1
        self.conv1 = SynthesisLayer(out_channels, out_channels, w_dim=w_dim, resolution=resolution, conv_clamp=conv_clamp,
                                    channels_last=self.channels_last, **layer_kwargs)
        self.num_conv += 1
        if is_last or architecture == 'skip':
            self.torgb = ToRGBLayer(out_channels, img_channels, w_dim=w_dim, conv_clamp=conv_clamp, channels_last=self.channels_last)
            self.num_torgb += 1

    def forward(self, x, img, ws, force_fp32=False, fused_modconv=None, update_emas=False, **layer_kwargs):
        _no_grad(x, img, ws)
        assert ws.shape[1:] == (self.num_conv + self.num_torgb, self.w_dim)
        w_iter = iter(ws.unbind(dim=1))
        dtype = torch.float16 if self.use_fp16 and not force_fp32 else torch.float32
        if self.in_channels == 0:
            x = self.const.detach().to(dtype=dtype).unsqueeze(0).repeat([ws.shape[0], 1, 1, 1])
        else:
            assert x.shape[1:] == (self.in_channels, self.resolution // self._up, self.resolution // self._up)
            x = x.to(dtype=dtype)
        if self.in_channels == 0:
            x = self.conv1(x, next(w_iter), **layer_kwargs)
        else:
            x = self.conv0(x, next(w_iter), **layer_kwargs)
            x = self.conv1(x, next(w_iter), **layer_kwargs)
        if img is not None and self._up == 2:
            assert img.shape[1:] == (self.img_channels, self.resolution // 2, self.resolution // 2)
            img = sg.upsample2d(img, self.resample_filter)
        if self.is_last or self.architecture == 'skip':
            y = self.torgb(x, next(w_iter))
            img = _accumulate_image(img, y)
        assert x.dtype == dtype
        assert img is None or img.dtype == torch.float32
        return x, img

    def extra_repr(self):
        return f'resolution={self.resolution:d}, architecture={self.architecture:s}'


class SynthesisNetwork(torch.nn.Module):
    """networks_stylegan2.py:469-526: the tri-plane backbone (img_channels = 96 at 256 x 256 in the reference's configuration)."""

    def __init__(self, w_dim, img_resolution, img_channels, channel_base=32768, channel_max=512, num_fp16_res=4, **block_kwargs):
        assert img_resolution >= 4 and img_resolution & (img_resolution - 1) == 0
        super().__init__()
        self.w_dim, self.img_resolution, self.img_channels, self.num_fp16_res = w_dim, img_resolution, img_channels, num_fp16_res
        self.img_resolution_log2 = int(np.log2(img_resolution))
        self.block_resolutions = [2 ** i for i in range(2, self.img_resolution_log2 + 1)]
        channels_dict = {res: min(channel_base // res, channel_max) for res in self.block_resolutions}
        fp16_resolution = max(2 ** (self.img_resolution_log2 + 1 - num_fp16_res), 8)
        self.num_ws = 0
        for res in self.block_resolutions:
            in_channels = channels_dict[res // 2] if res > 4 else 0
            block = SynthesisBlock(in_channels, channels_dict[res], w_dim=w_dim, resolution=res, img_channels=img_channels,
                                   is_last=(res == self.img_resolution), use_fp16=(res >= fp16_resolution), **block_kwargs)
            self.num_ws += block.num_conv
            if res == self.img_resolution:
                self.num_ws += block.num_torgb
            setattr(self, f'b{res}', block)

    def forward(self, ws, update_emas=False, **block_kwargs):
        _no_grad(ws)
        assert ws.shape[1:] == (self.num_ws, self.w_dim)
        ws = ws.to(torch.float32)
        block_ws, w_idx = [], 0
        for res in self.block_resolutions:
            block = getattr(self, f'b{res}')
            block_ws.append(ws.narrow(1, w_idx, block.num_conv + block.num_torgb))
            w_idx += block.num_conv
        x = img = None
        for res, cur_ws in zip(self.block_resolutions, block_ws):
            x, img = getattr(self, f'b{res}')(x, img, cur_ws, **block_kwargs)
        return img


class Generator(torch.nn.Module):
    """networks_stylegan2.py:529-557: mapping + synthesis (the `StyleGAN2Backbone` of training/triplane.py:14,47)."""

    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, mapping_kwargs={}, **synthesis_kwargs):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.img_resolution, self.img_channels = z_dim, c_dim, w_dim, img_resolution, img_channels
        self.synthesis = SynthesisNetwork(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels, **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, num_ws=self.num_ws, **mapping_kwargs)

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False, **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.synthesis(ws, **synthesis_kwargs)


class _SuperresolutionHybrid(torch.nn.Module):
    """Shared body of superresolution.py:29-125,264-290: optional bilinear pre-resize of the feature image and its RGB slice to
    `input_resolution` (the library's resize kernel), then two synthesis blocks driven by the last w."""

    def __init__(self, channels, img_resolution, sr_num_fp16_res, sr_antialias, expected_resolution, input_resolution, block0, block1):
        super().__init__()
        assert img_resolution == expected_resolution
        self.input_resolution = input_resolution
        self.sr_antialias = sr_antialias
        self.block0, self.block1 = block0, block1

    def forward(self, rgb, x, ws, **block_kwargs):
        _no_grad(rgb, x, ws)
        ws = ws[:, -1:, :].repeat(1, 3, 1)
        if x.shape[-1] != self.input_resolution:
            size = (self.input_resolution, self.input_resolution)
            x = resize_bilinear(x.float().contiguous(), size, antialias=self.sr_antialias)
            rgb = resize_bilinear(rgb.float().contiguous(), size, antialias=self.sr_antialias)
        x, rgb = self.block0(x, rgb, ws, **block_kwargs)
        x, rgb = self.block1(x, rgb, ws, **block_kwargs)
        return rgb


def _sr_blocks(channels, sr_num_fp16_res, specs, block_kwargs):
    use_fp16 = sr_num_fp16_res > 0
    out = []
    for in_ch, out_ch, res, is_last, block_cls in specs:
        out.append(block_cls(in_ch, out_ch, w_dim=512, resolution=res, img_channels=3, is_last=is_last, use_fp16=use_fp16,
                             conv_clamp=(256 if use_fp16 else None), **block_kwargs))
    return out


class SuperresolutionHybrid8XDC(_SuperresolutionHybrid):
    """superresolution.py:264-290 (the 512 x 512 default, train.py:276-277): 128 -> 256 (256 ch) -> 512 (128 ch)."""

    def __init__(self, channels, img_resolution, sr_num_fp16_res, sr_antialias, num_fp16_res=4, conv_clamp=None, channel_base=None,
                 channel_max=None, **block_kwargs):
        b0, b1 = _sr_blocks(channels, sr_num_fp16_res, [(channels, 256, 256, False, SynthesisBlock), (256, 128, 512, True, SynthesisBlock)], block_kwargs)
        super().__init__(channels, img_resolution, sr_num_fp16_res, sr_antialias, 512, 128, b0, b1)


class SuperresolutionHybrid8X(_SuperresolutionHybrid):
    """superresolution.py:29-58: 128 -> 256 (128 ch) -> 512 (64 ch)."""

    def __init__(self, channels, img_resolution, sr_num_fp16_res, sr_antialias, num_fp16_res=4, conv_clamp=None, channel_base=None,
                 channel_max=None, **block_kwargs):
        b0, b1 = _sr_blocks(channels, sr_num_fp16_res, [(channels, 128, 256, False, SynthesisBlock), (128, 64, 512, True, SynthesisBlock)], block_kwargs)
        super().__init__(channels, img_resolution, sr_num_fp16_res, sr_antialias, 512, 128, b0, b1)
        self.register_buffer('resample_filter', sg.setup_filter([1, 3, 3, 1]))


class SynthesisBlockNoUp(SynthesisBlock):
    """superresolution.py:158-253: the synthesis block whose first convolution keeps the resolution; the RGB skip is not up-sampled."""
    _up = 1


class SuperresolutionHybrid4X(_SuperresolutionHybrid):
    """superresolution.py:62-90: 128 -> 128 (128 ch, no up-sampling) -> 256 (64 ch)."""

    def __init__(self, channels, img_resolution, sr_num_fp16_res, sr_antialias, num_fp16_res=4, conv_clamp=None, channel_base=None,
                 channel_max=None, **block_kwargs):
        b0, b1 = _sr_blocks(channels, sr_num_fp16_res, [(channels, 128, 128, False, SynthesisBlockNoUp), (128, 64, 256, True, SynthesisBlock)], block_kwargs)
        super().__init__(channels, img_resolution, sr_num_fp16_res, sr_antialias, 256, 128, b0, b1)
        self.register_buffer('resample_filter', sg.setup_filter([1, 3, 3, 1]))

    def forward(self, rgb, x, ws, **block_kwargs):
        _no_grad(rgb, x, ws)
        ws = ws[:, -1:, :].repeat(1, 3, 1)
        if x.shape[-1] < self.input_resolution:
            size = (self.input_resolution, self.input_resolution)
            x = resize_bilinear(x.float().contiguous(), size, antialias=self.sr_antialias)
            rgb = resize_bilinear(rgb.float().contiguous(), size, antialias=self.sr_antialias)
        x, rgb = self.block0(x, rgb, ws, **block_kwargs)
        x, rgb = self.block1(x, rgb, ws, **block_kwargs)
        return rgb


class SuperresolutionHybrid2X(_SuperresolutionHybrid):
    """superresolution.py:94-125: 64 -> 64 (128 ch, no up-sampling) -> 128 (64 ch)."""

    def __init__(self, channels, img_resolution, sr_num_fp16_res, sr_antialias, num_fp16_res=4, conv_clamp=None, channel_base=None,
                 channel_max=None, **block_kwargs):
        b0, b1 = _sr_blocks(channels, sr_num_fp16_res, [(channels, 128, 64, False, SynthesisBlockNoUp), (128, 64, 128, True, SynthesisBlock)], block_kwargs)
        super().__init__(channels, img_resolution, sr_num_fp16_res, sr_antialias, 128, 64, b0, b1)
        self.register_buffer('resample_filter', sg.setup_filter([1, 3, 3, 1]))
