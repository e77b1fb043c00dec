"""Drop-in for the reference's `models.stylegan2.op` package (models/stylegan2/op/__init__.py:1-2)."""
from .fused_act import FusedLeakyReLU, fused_leaky_relu
from .upfirdn2d import upfirdn2d, upfirdn2d_native

__all__ = ["FusedLeakyReLU", "fused_leaky_relu", "upfirdn2d", "upfirdn2d_native"]
