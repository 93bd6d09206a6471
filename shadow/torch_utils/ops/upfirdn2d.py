"""Module-path shadow of torch_utils/ops/upfirdn2d.py: CUDA tensors run nerffaceediting_b200's kernels (differentiable, the backward
being the same op with the factors exchanged); CPU tensors and `impl='ref'` go to the reference's own pure-PyTorch path."""
from nerffaceediting_b200 import stylegan_ops as _impl
from nerffaceediting_b200.stylegan_ops import _get_filter_size, _parse_padding, _parse_scaling, setup_filter  # noqa: F401

from . import load_reference


def _cuda(x, impl):
    return impl == 'cuda' and x.is_cuda and x.dtype in _impl._DTYPES


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    if _cuda(x, impl):
        return _impl.upfirdn2d(x, f, up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain)
    return load_reference("upfirdn2d")._upfirdn2d_ref(x, f, up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain)


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    if _cuda(x, impl):
        return _impl.filter2d(x, f, padding=padding, flip_filter=flip_filter, gain=gain)
    return load_reference("upfirdn2d").filter2d(x, f, padding=padding, flip_filter=flip_filter, gain=gain, impl='ref')


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    if _cuda(x, impl):
        return _impl.upsample2d(x, f, up=up, padding=padding, flip_filter=flip_filter, gain=gain)
    return load_reference("upfirdn2d").upsample2d(x, f, up=up, padding=padding, flip_filter=flip_filter, gain=gain, impl='ref')


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    if _cuda(x, impl):
        return _impl.downsample2d(x, f, down=down, padding=padding, flip_filter=flip_filter, gain=gain)
    return load_reference("upfirdn2d").downsample2d(x, f, down=down, padding=padding, flip_filter=flip_filter, gain=gain, impl='ref')
