"""where2edit_b200 -- B200-native (sm_100a) implementation of Where2edit's StyleGAN2 synthesis hot path.

Public surface = the reference's (models/stylegan2/model.py, attention/attention_model.py,
models/stylegan2/op): Generator, ModulatedConv2d, StyledConv, ToRGB, EqualLinear, PixelNorm,
EqualConv2d, upfirdn2d, fused_leaky_relu, FusedLeakyReLU.  `install_as_reference()` registers the
package under the reference's module names so existing scripts import it unchanged.
"""
from . import _native
from .model import (Blur, ConstantInput, Downsample, EqualConv2d, EqualLinear, Generator, ModulatedConv2d,
                    NoiseInjection, PixelNorm, ScaledLeakyReLU, StyledConv, ToRGB, Upsample, make_kernel)
from .op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d, upfirdn2d_native
from .graph import GraphedGenerator, GraphedStep

__version__ = "0.1.0"


def enable_weight_gradients(module, enabled=True):
    """Turn the (opt-in) convolution-weight gradient on for every ModulatedConv2d under `module` -- for networks
    that TRAIN StyledConv weights (ToRGB runs a fused kernel and is not covered), e.g. the cluster-style mapper's attention heads
    (attention/run_attention.py:725-735).  The generator stays frozen in every caller and does not need it."""
    for m in module.modules():
        if isinstance(m, ModulatedConv2d):
            m.weight_grad = bool(enabled)
    return module


def install_as_reference():
    """Make `from models.stylegan2.op import upfirdn2d`, `from models.stylegan2.model import Generator`,
    `from attention_model import Generator` and `from attention.attention_model import ...` resolve to
    this package (the import lines of attention/run_attention.py:29, mapper/latent_mappers.py:5, ...)."""
    import sys
    import types

    from . import model, op
    from .op import fused_act as _fa
    from .op import upfirdn2d as _up

    def pkg(name):
        """The reference's REAL package when it is importable (its other submodules -- models.facial_recognition,
        models.encoders, models.psp, ... -- must keep resolving; only the stylegan2 leaves are overridden);
        an empty stub package only when the reference tree is not on sys.path at all."""
        import importlib
        m = sys.modules.get(name)
        if m is None:
            try:
                m = importlib.import_module(name)
            except ImportError:
                m = types.ModuleType(name)
                m.__path__ = []
                sys.modules[name] = m
        return m

    pkg("models")
    pkg("models.stylegan2")
    pkg("attention")
    sys.modules["models.stylegan2.op"] = op
    sys.modules["models.stylegan2.op.upfirdn2d"] = _up
    sys.modules["models.stylegan2.op.fused_act"] = _fa
    sys.modules["models.stylegan2.model"] = model
    sys.modules["attention_model"] = model
    sys.modules["attention.attention_model"] = model
    sys.modules["models.stylegan2"].op = op
    sys.modules["models.stylegan2"].model = model
    sys.modules["attention"].attention_model = model


__all__ = ["Generator", "ModulatedConv2d", "StyledConv", "ToRGB", "EqualLinear", "EqualConv2d", "PixelNorm",
           "Blur", "Upsample", "Downsample", "NoiseInjection", "ConstantInput", "ScaledLeakyReLU", "make_kernel",
           "FusedLeakyReLU", "fused_leaky_relu", "upfirdn2d", "upfirdn2d_native", "install_as_reference",
           "enable_weight_gradients", "GraphedGenerator", "GraphedStep"]
