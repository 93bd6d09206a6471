"""Module-path shadow of torch_utils/ops/conv2d_resample.py (:48-143).  What a generator's layers ask for at inference — 3x3 / 1x1
kernels, up in {1, 2}, padding = kernel // 2, groups = 1 or the fused modulated_conv2d's grouped per-sample form, fp16 / fp32 CUDA
tensors, no gradient — runs nerffaceediting_b200's tcgen05 implicit GEMM (csrc/nfe_modconv.cu, nfe_modulated_conv2d); everything else
(training, the discriminator's down-sampling convolutions, CPU tensors, odd channel counts) runs the reference's own function."""
import torch

from nerffaceediting_b200 import networks as _impl

from . import load_reference


def _fast(x, w, f, up, down, padding, groups, flip_filter):
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.ndim == 4 and x.dtype in (torch.float16, torch.float32) and w.dtype == x.dtype):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad):
        return False
    kh, kw = int(w.shape[-2]), int(w.shape[-1])
    pad = [padding] * 4 if isinstance(padding, int) else ([padding[0], padding[0], padding[1], padding[1]] if len(padding) == 2 else list(padding))
    if down != 1 or up not in (1, 2) or kh != kw or kh not in (1, 3) or (up == 2 and kh != 3) or any(p != kh // 2 for p in pad) or flip_filter:
        return False
    if up == 2 and (f is None or f.ndim not in (1, 2) or max(f.shape) > 4):
        return False
    if groups != 1 and (x.shape[0] != 1 or x.shape[1] % groups or w.shape[0] % groups):
        return False
    in_ch, out_ch = int(w.shape[1]), int(w.shape[0]) // groups
    return in_ch % 16 == 0 and (out_ch <= 256 or out_ch % 128 == 0) and (up == 1 or out_ch % 8 == 0)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    if _fast(x, w, f, up, down, padding, groups, flip_filter):
        return _impl.conv2d_resample(x, w, f=f, up=up, down=down, padding=padding, groups=groups, flip_weight=flip_weight, flip_filter=flip_filter)
    return load_reference("conv2d_resample").conv2d_resample(x=x, w=w, f=f, up=up, down=down, padding=padding, groups=groups,
                                                             flip_weight=flip_weight, flip_filter=flip_filter)
