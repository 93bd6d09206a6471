"""Decoders and plane statistics of the hot path (reference: training/triplane.py:56-68,167-270).

* OSGDecoder / SegmentationOSGDecoder / DisentangledOSGDecoder keep the reference's constructor
  `(n_features, options)`, parameter names (`net.0.weight`, `geo_net.2.bias`, ...) and call signatures, so
  state dicts load unchanged; stand-alone calls run the fused decoder kernel (csrc/nfe_field.cu).
* compute_mean_var / normalize_plane / denormalize_plane are the free-function form of the
  TriPlaneGenerator methods (triplane.py:56-68; notebook twins utils.py:146-158).
"""
import numpy as np
import torch

from . import ops


class FullyConnectedLayer(torch.nn.Module):
    """Parameter container with the semantics of training/networks_stylegan2.py:96-127 for the case the
    decoders use (activation='linear', bias=True): y = addmm(bias*bias_gain, x, (weight*weight_gain)^T)."""

    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1, bias_init=0):
        super().__init__()
        assert activation == 'linear' and bias, "only the linear, biased layer of the decoders is provided"
        self.in_features = in_features
        self.out_features = out_features
        self.activation = activation
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.full([out_features], np.float32(bias_init)))
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def extra_repr(self):
        return f'in_features={self.in_features:d}, out_features={self.out_features:d}, activation={self.activation:s}'


def _mlp(n_features, hidden, out, lr_mul):
    return torch.nn.Sequential(FullyConnectedLayer(n_features, hidden, lr_multiplier=lr_mul), torch.nn.Softplus(),
                               FullyConnectedLayer(hidden, out, lr_multiplier=lr_mul))


class OSGDecoder(torch.nn.Module):
    """mean over planes -> 32->64->33; sigma = y[0], rgb = sigmoid(y[1:])*1.002-0.001 (triplane.py:167-190)."""

    def __init__(self, n_features, options):
        super().__init__()
        self.hidden_dim = 64
        self.net = _mlp(n_features, self.hidden_dim, 1 + options['decoder_output_dim'], options['decoder_lr_mul'])

    def forward(self, sampled_features, ray_directions):
        ops._no_grad_needed(sampled_features, *self.parameters())
        return ops.decoder_fwd(ops.DEC_OSG, self.net, None, None, sampled_features)


class SegmentationOSGDecoder(torch.nn.Module):
    """`disable_alignment` ablation: net (32->64->33) and seg_net (32->64->15), both on the de-normalised
    features; sampled_norm_features is ignored (triplane.py:192-230)."""

    def __init__(self, n_features, options):
        super().__init__()
        self.hidden_dim = 64
        self.net = _mlp(n_features, self.hidden_dim, 1 + options['decoder_output_dim'], options['decoder_lr_mul'])
        self.seg_net = _mlp(n_features, self.hidden_dim, options['decoder_seg_dim'], options['decoder_lr_mul'])

    def forward(self, sampled_norm_features, sampled_denorm_features, ray_directions):
        ops._no_grad_needed(sampled_denorm_features, *self.parameters())
        return ops.decoder_fwd(ops.DEC_SEGMENTATION, self.net, self.seg_net, None, sampled_denorm_features)


class DisentangledOSGDecoder(torch.nn.Module):
    """geo_net (32->64->16) on normalised features -> sigma, 15 semantic logits; app_net (32->64->32) on
    de-normalised features -> rgb (triplane.py:232-270)."""

    def __init__(self, n_features, options):
        super().__init__()
        self.hidden_dim = 64
        self.geo_net = _mlp(n_features, self.hidden_dim, 1 + options['decoder_seg_dim'], options['decoder_lr_mul'])
        self.app_net = _mlp(n_features, self.hidden_dim, options['decoder_output_dim'], options['decoder_lr_mul'])

    def forward(self, sampled_norm_features, sampled_denorm_features, ray_directions):
        ops._no_grad_needed(sampled_norm_features, sampled_denorm_features, *self.parameters())
        return ops.decoder_fwd(ops.DEC_DISENTANGLED, self.geo_net, self.app_net, sampled_norm_features, sampled_denorm_features)


# ---------------------------------------------------------------------------------------- plane statistics
def compute_mean_var(planes):
    """(mean, std) over H,W with keepdim; `var` in the reference's naming is the STD (triplane.py:56-60)."""
    ops._no_grad_needed(planes)
    return ops.plane_stats(planes)


def normalize_plane(planes):
    """(planes - mean) / (std + 1e-8) -> (norm_planes, mean, std) (triplane.py:61-65)."""
    tri = planes.dim() == 4 and planes.shape[1] == 96 and planes.is_contiguous() and planes.dtype == torch.float32
    if torch.is_grad_enabled() and planes.requires_grad:
        from .autograd import NormalizeFunction
        norm, mean, std, norm_cl = NormalizeFunction.apply(planes)
        if norm_cl is not None:
            ops.register_normalized(planes, norm, mean, std, norm_cl)
            ops.provenance_attach(planes, std, 1e-8, mean)       # planes == norm*(std+1e-8) + mean: lets the renderer gather one set
        return norm, mean, std
    mean, std = ops.plane_stats(planes)
    if tri:
        # the generator's [N,96,H,W] tri-planes: also stage the planes for the renderer in the same pass
        return ops.plane_normalize_staged(planes, mean, std)[0], mean, std
    return ops.plane_normalize(planes, mean, std), mean, std


def denormalize_plane(planes, mean, var):
    """planes * std + mean (triplane.py:66-68).  Statistics may belong to another identity (appearance
    swap), or to one batch item broadcast over the batch (triplane.py:98-103)."""
    if torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in (planes, mean, var)):
        from .autograd import DenormalizeFunction
        out = DenormalizeFunction.apply(planes, mean, var)
        if torch.is_tensor(mean) and torch.is_tensor(var):
            ops.register_denormalized(out, planes, mean, var)
            ops.provenance_attach(out, var, 0.0, mean, norm_requires_grad=planes.requires_grad)
        return out
    return ops.plane_denormalize(planes, mean, var)
