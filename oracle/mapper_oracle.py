"""CPU restatement (numpy, fp32 storage / fp64 accumulation) of the reference's latent mappers — TEST
INFRASTRUCTURE (only tests/, smoke() and bench.py's cpu_baseline leg may import it).

Follows mapper/latent_mappers.py:10-82 over models/stylegan2/model.py:11-17 (PixelNorm, dim 1), :130-159
(EqualLinear, lr_mul scaling) and op/fused_act.py:23-39 (bias on the LAST dim for 3-D inputs).  Pinned by
tests/golden/mapper.npz (oracle/make_mapper_golden.py runs the unmodified reference LevelsMapper).
`mapper_state` generates the seeded weights both sides load (the 12.6 MB of weights are not stored).
"""
import math

import numpy as np

F32 = np.float32
LEVELS = ("course_mapping", "medium_mapping", "fine_mapping")   # the reference's attribute names (:55-59)


def mapper_state(seed=5, latent_dim=512, levels=LEVELS):
    """State dict (numpy fp32) of a LevelsMapper with the reference's key names: weights ~ randn / lr_mul as in
    EqualLinear.__init__ (model.py:136), biases ~ 0.5 * randn so that the bias path is live."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}
    for lv in levels:
        for i in range(1, 5):
            sd[f"{lv}.mapping.{i}.weight"] = (rng.standard_normal((latent_dim, latent_dim), dtype=F32) / F32(0.01)).astype(F32)
            sd[f"{lv}.mapping.{i}.bias"] = (50.0 * rng.standard_normal(latent_dim)).astype(F32)
    return sd


def mapper_forward(x, sd, prefix, lr_mul=0.01):
    """Mapper.forward on x [B, L, D] (latent_mappers.py:26-29)."""
    x = np.asarray(x, np.float64)
    x = x / np.sqrt(np.mean(x * x, axis=1, keepdims=True) + 1e-8)        # PixelNorm over dim 1 (model.py:16)
    d = x.shape[-1]
    scale = (1.0 / math.sqrt(d)) * lr_mul
    for i in range(1, 5):
        w = sd[f"{prefix}.mapping.{i}.weight"].astype(np.float64) * scale
        b = sd[f"{prefix}.mapping.{i}.bias"].astype(np.float64) * lr_mul
        x = x @ w.T + b
        x = np.where(x > 0, x, 0.2 * x) * math.sqrt(2.0)
    return x


def levels_mapper(x, sd, no_coarse=False, no_medium=False, no_fine=False):
    """LevelsMapper.forward (latent_mappers.py:61-82)."""
    x = np.asarray(x, F32)
    parts = []
    for sl, off, name in ((slice(0, 4), no_coarse, LEVELS[0]), (slice(4, 8), no_medium, LEVELS[1]),
                          (slice(8, None), no_fine, LEVELS[2])):
        xs = x[:, sl, :]
        parts.append(np.zeros_like(xs, dtype=np.float64) if off else mapper_forward(xs, sd, name))
    return np.concatenate(parts, axis=1).astype(F32)
