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


class TriPlaneGenerator(torch.nn.Module):
    """The whole generator of training/triplane.py:18-165 (BASELINE configs[1]: mapping + StyleGAN2 tri-plane backbone + decoders +
    renderer + super-resolution), composed of this package's modules only: same constructor, attribute / state-dict names
    (`backbone.mapping.fc0.weight`, `backbone.synthesis.b64.conv1.weight`, `decoder.geo_net.0.weight`, `superresolution.block1...`) and
    methods (`mapping`, `synthesis`, `sample`, `sample_mixed`, `forward`), so `G2.load_state_dict(G.state_dict())` moves a checkpoint over.
    Inference (the convolution stack is forward-only); the plane statistics go through `normalize_plane`, which stages the planes and
    registers their provenance, so the renderer takes the single-gather path."""

    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, sr_num_fp16_res=0, mapping_kwargs={}, rendering_kwargs={},
                 sr_kwargs={}, disable_disentangle=False, disable_alignment=False, **synthesis_kwargs):
        super().__init__()
        from . import networks
        from .ray_sampler import RaySampler
        from .renderer import DisentangledImportanceRenderer
        self.z_dim, self.c_dim, self.w_dim, self.img_resolution, self.img_channels = z_dim, c_dim, w_dim, img_resolution, img_channels
        self.disable_disentangle, self.disable_alignment = disable_disentangle, disable_alignment
        assert not self.disable_alignment or disable_disentangle
        self.renderer = DisentangledImportanceRenderer()
        self.ray_sampler = RaySampler()
        self.backbone = networks.Generator(z_dim, c_dim, w_dim, img_resolution=256, img_channels=32 * 3, mapping_kwargs=mapping_kwargs, **synthesis_kwargs)
        sr_cls = getattr(networks, rendering_kwargs['superresolution_module'].split('.')[-1])        # 'training.superresolution.SuperresolutionHybrid8XDC'
        self.superresolution = sr_cls(channels=32, img_resolution=img_resolution, sr_num_fp16_res=sr_num_fp16_res,
                                      sr_antialias=rendering_kwargs['sr_antialias'], **sr_kwargs)
        dec_opts = {'decoder_lr_mul': rendering_kwargs.get('decoder_lr_mul', 1), 'decoder_output_dim': 32, 'decoder_seg_dim': 15}
        self.decoder = DisentangledOSGDecoder(32, dec_opts) if not self.disable_alignment else SegmentationOSGDecoder(32, dec_opts)
        self.neural_rendering_resolution = 64
        self.rendering_kwargs = rendering_kwargs
        self._last_planes = None

    compute_mean_var = staticmethod(compute_mean_var)
    normalize_plane = staticmethod(normalize_plane)
    denormalize_plane = staticmethod(denormalize_plane)

    def mapping(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False):
        if self.rendering_kwargs['c_gen_conditioning_zero']:
            c = torch.zeros_like(c)
        return self.backbone.mapping(z, c * self.rendering_kwargs.get('c_scale', 0), truncation_psi=truncation_psi,
                                     truncation_cutoff=truncation_cutoff, update_emas=update_emas)

    def _planes(self, ws, update_emas=False, **synthesis_kwargs):
        return self.backbone.synthesis(ws, update_emas=update_emas, **synthesis_kwargs)

    def synthesis(self, ws, c, neural_rendering_resolution=None, update_emas=False, cache_backbone=False, use_cached_backbone=False,
                  planes_mean=None, planes_var=None, **synthesis_kwargs):
        cam2world_matrix = c[:, :16].view(-1, 4, 4)
        intrinsics = c[:, 16:25].view(-1, 3, 3)
        if neural_rendering_resolution is None:
            neural_rendering_resolution = self.neural_rendering_resolution
        else:
            self.neural_rendering_resolution = neural_rendering_resolution
        ray_origins, ray_directions = self.ray_sampler(cam2world_matrix, intrinsics, neural_rendering_resolution)
        N, M, _ = ray_origins.shape
        if use_cached_backbone and self._last_planes is not None:
            planes = self._last_planes
        else:
            planes = self._planes(ws, update_emas=update_emas, **synthesis_kwargs)
        if not self.disable_disentangle:
            norm_planes, mean, var = self.normalize_plane(planes)
            if planes_mean is not None and planes_var is not None:              # appearance swap (triplane.py:98-103)
                if type(planes_mean) == int and type(planes_var) == int:
                    planes = self.denormalize_plane(norm_planes, mean[planes_mean][None, ...], var[planes_var][None, ...])
                else:
                    planes = self.denormalize_plane(norm_planes, planes_mean, planes_var)
        else:
            norm_planes = mean = var = None
        if cache_backbone:
            self._last_planes = planes
        if not self.disable_disentangle:
            norm_planes = norm_planes.view(len(norm_planes), 3, 32, norm_planes.shape[-2], norm_planes.shape[-1])
        planes = planes.view(len(planes), 3, 32, planes.shape[-2], planes.shape[-1])
        feature_samples, seg_samples, depth_samples, weights_samples = self.renderer(
            norm_planes if not self.disable_disentangle else planes, planes, self.decoder, ray_origins, ray_directions, self.rendering_kwargs)
        H = W = self.neural_rendering_resolution
        feature_image = feature_samples.permute(0, 2, 1).reshape(N, feature_samples.shape[-1], H, W).contiguous()
        seg_image = seg_samples.permute(0, 2, 1).reshape(N, seg_samples.shape[-1], H, W).contiguous()
        depth_image = depth_samples.permute(0, 2, 1).reshape(N, 1, H, W)
        rgb_image = feature_image[:, :3]
        sr_image = self.superresolution(rgb_image, feature_image, ws, noise_mode=self.rendering_kwargs['superresolution_noise_mode'],
                                        **{k: synthesis_kwargs[k] for k in synthesis_kwargs.keys() if k != 'noise_mode'})
        return {'image': sr_image, 'image_seg': seg_image, 'image_raw': rgb_image, 'image_depth': depth_image, 'plane_mean': mean, 'plane_var': var}

    def _run_model(self, ws, coordinates, directions, update_emas, synthesis_kwargs):
        planes = self._planes(ws, update_emas=update_emas, **synthesis_kwargs)
        norm_planes = None
        if not self.disable_disentangle:
            norm_planes, _, _ = self.normalize_plane(planes)
            norm_planes = norm_planes.view(len(norm_planes), 3, 32, norm_planes.shape[-2], norm_planes.shape[-1])
        planes = planes.view(len(planes), 3, 32, planes.shape[-2], planes.shape[-1])
        return self.renderer.run_model(norm_planes if not self.disable_disentangle else planes, planes, self.decoder, coordinates, directions,
                                       self.rendering_kwargs)

    def sample(self, coordinates, directions, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False, **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self._run_model(ws, coordinates, directions, update_emas, synthesis_kwargs)

    def sample_mixed(self, coordinates, directions, ws, truncation_psi=1, truncation_cutoff=None, update_emas=False, **synthesis_kwargs):
        return self._run_model(ws, coordinates, directions, update_emas, synthesis_kwargs)

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, neural_rendering_resolution=None, update_emas=False, cache_backbone=False,
                use_cached_backbone=False, planes_mean=None, planes_var=None, **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.synthesis(ws, c, update_emas=update_emas, neural_rendering_resolution=neural_rendering_resolution, cache_backbone=cache_backbone,
                              use_cached_backbone=use_cached_backbone, planes_mean=planes_mean, planes_var=planes_var, **synthesis_kwargs)
