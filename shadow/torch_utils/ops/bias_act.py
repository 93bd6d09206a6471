"""Module-path shadow of torch_utils/ops/bias_act.py: CUDA tensors run nerffaceediting_b200's kernel (first- and second-order
gradients included, csrc/nfe_stylegan_ops.cu); CPU tensors and `impl='ref'` go to the reference's own pure-PyTorch path."""
from nerffaceediting_b200 import stylegan_ops as _impl
from nerffaceediting_b200.stylegan_ops import activation_funcs  # noqa: F401

from . import load_reference


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    if impl == 'cuda' and x.is_cuda and x.dtype in _impl._DTYPES:
        return _impl.bias_act(x, b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp)
    return load_reference("bias_act")._bias_act_ref(x=x, b=b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp)
