"""Latent mappers of the StyleCLIP-style edit loop: the caller-side MLP stacks
whose output `w_hat = w + 0.1 * mapper(w)` feeds Generator.forward (mapper/styleclip_mapper.py, BASELINE cfg3).

Mirrors mapper/latent_mappers.py:10-82 — same class names, constructor (`opts` with no_coarse_mapper /
no_medium_mapper / no_fine_mapper), attribute names (including the reference's `course_mapping` spelling, so
its checkpoints load) and forward semantics — on this package's EqualLinear / PixelNorm, whose fused
bias + leaky-ReLU runs in libw2e (w2e_bias_act_fwd/bwd); the 512x512 products are plain library GEMMs.
The reference's own file also works unmodified after `install_as_reference()`, since it only imports
`EqualLinear, PixelNorm` from models.stylegan2.model (:5).
"""
import torch
from torch import nn

from .model import EqualLinear, PixelNorm


class Mapper(nn.Module):
    """mapper/latent_mappers.py:10-29: PixelNorm (over dim 1) + 4 x EqualLinear(lr_mul=0.01, fused_lrelu)."""

    def __init__(self, opts, latent_dim=512):
        super().__init__()
        self.opts = opts
        layers = [PixelNorm()]
        for _ in range(4):
            layers.append(EqualLinear(latent_dim, latent_dim, lr_mul=0.01, activation="fused_lrelu"))
        self.mapping = nn.Sequential(*layers)

    def forward(self, x):
        return self.mapping(x)


class SingleMapper(nn.Module):
    """mapper/latent_mappers.py:32-44."""

    def __init__(self, opts):
        super().__init__()
        self.opts = opts
        self.mapping = Mapper(opts)

    def forward(self, x):
        return self.mapping(x)


class LevelsMapper(nn.Module):
    """mapper/latent_mappers.py:47-82: separate mappers for W+ rows [0,4), [4,8), [8,...); disabled levels
    contribute zeros."""

    def __init__(self, opts):
        super().__init__()
        self.opts = opts
        if not opts.no_coarse_mapper:
            self.course_mapping = Mapper(opts)
        if not opts.no_medium_mapper:
            self.medium_mapping = Mapper(opts)
        if not opts.no_fine_mapper:
            self.fine_mapping = Mapper(opts)

    def forward(self, x):
        x_coarse, x_medium, x_fine = x[:, :4, :], x[:, 4:8, :], x[:, 8:, :]
        x_coarse = torch.zeros_like(x_coarse) if self.opts.no_coarse_mapper else self.course_mapping(x_coarse)
        x_medium = torch.zeros_like(x_medium) if self.opts.no_medium_mapper else self.medium_mapping(x_medium)
        x_fine = torch.zeros_like(x_fine) if self.opts.no_fine_mapper else self.fine_mapping(x_fine)
        return torch.cat([x_coarse, x_medium, x_fine], dim=1)


__all__ = ["Mapper", "SingleMapper", "LevelsMapper"]
