"""Drop-in for the reference's `networks/stylegan2/op` package
(scf/networks/stylegan2/op/__init__.py:1-2): same names, same argument meaning, same error behaviour."""
from .bias_act import FusedLeakyReLU, fused_leaky_relu, fused_bias_act
from .fir_resample import upfirdn2d, upfirdn2d_op

__all__ = ['FusedLeakyReLU', 'fused_leaky_relu', 'fused_bias_act', 'upfirdn2d', 'upfirdn2d_op']
