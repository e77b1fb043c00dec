"""Seeded synthetic state for the StyleGAN2 synthesis path (TEST INFRASTRUCTURE).

The reference ships no weights (pretrained_models/ReadME.md only links), so every parity
check runs on random-init state.  torch's RNG stream is not guaranteed stable across
versions, so the state is drawn from numpy's PCG64 instead and keyed exactly like the
reference ``Generator.state_dict()`` (reference constructors at
models/stylegan2/model.py:136-143, 220-224, 283, 297, 351, 420-423 and
models/stylegan2/op/fused_act.py:15).

Variant "plain" follows the reference init distributions (randn weights, modulation bias 1,
noise weight / activation bias / rgb bias 0).  Variant "perturbed" additionally draws
noise.weight, activate.bias and to_rgb bias from N(0, 0.1^2) so that those code paths are
live.
"""
import math

import numpy as np
import torch

CHANNELS_BASE = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256, 128: 128, 256: 64, 512: 32, 1024: 16}


def channels_for(size, channel_multiplier=2):
    # models/stylegan2/model.py:392-402
    ch = {}
    for res, c in CHANNELS_BASE.items():
        ch[res] = c if res <= 32 else c * channel_multiplier
    return ch


def blur_kernel_2d(taps=(1, 3, 3, 1), gain=1.0):
    # make_kernel, models/stylegan2/model.py:20-28
    k = np.asarray(taps, dtype=np.float32)
    k2 = k[None, :] * k[:, None]
    k2 = k2 / k2.sum()
    return torch.from_numpy((k2 * np.float32(gain)).astype(np.float32))


def make_state_dict(size, style_dim=512, n_mlp=8, channel_multiplier=2, seed=0,
                    perturbed=False, lr_mlp=0.01):
    """Returns an OrderedDict-like dict of fp32 CPU tensors with the reference's keys/shapes."""
    rng = np.random.Generator(np.random.PCG64(seed))
    prng = np.random.Generator(np.random.PCG64(seed + 1000003))

    def randn(*shape):
        return torch.from_numpy(rng.standard_normal(shape, dtype=np.float32))

    def small(*shape):
        if perturbed:
            return torch.from_numpy((0.1 * prng.standard_normal(shape)).astype(np.float32))
        return torch.zeros(*shape)

    ch = channels_for(size, channel_multiplier)
    log_size = int(math.log2(size))
    sd = {}
    for i in range(n_mlp):
        sd[f"style.{i + 1}.weight"] = randn(style_dim, style_dim) / lr_mlp
        sd[f"style.{i + 1}.bias"] = small(style_dim) / lr_mlp if perturbed else torch.zeros(style_dim)
    sd["input.input"] = randn(1, ch[4], 4, 4)

    def modconv(prefix, cin, cout, k, upsample):
        sd[f"{prefix}.weight"] = randn(1, cout, cin, k, k)
        if upsample:
            sd[f"{prefix}.blur.kernel"] = blur_kernel_2d(gain=4.0)
        sd[f"{prefix}.modulation.weight"] = randn(cin, style_dim)
        sd[f"{prefix}.modulation.bias"] = torch.ones(cin)

    def styled(prefix, cin, cout, upsample):
        modconv(f"{prefix}.conv", cin, cout, 3, upsample)
        sd[f"{prefix}.noise.weight"] = small(1)
        sd[f"{prefix}.activate.bias"] = small(cout)

    def torgb(prefix, cin, upsample):
        sd[f"{prefix}.bias"] = small(1, 3, 1, 1)
        if upsample:
            sd[f"{prefix}.upsample.kernel"] = blur_kernel_2d(gain=4.0)
        modconv(f"{prefix}.conv", cin, 3, 1, False)

    styled("conv1", ch[4], ch[4], False)
    torgb("to_rgb1", ch[4], False)
    cin = ch[4]
    for j, i in enumerate(range(3, log_size + 1)):
        cout = ch[2 ** i]
        styled(f"convs.{2 * j}", cin, cout, True)
        styled(f"convs.{2 * j + 1}", cout, cout, False)
        torgb(f"to_rgbs.{j}", cout, True)
        cin = cout
    num_layers = (log_size - 2) * 2 + 1
    for layer_idx in range(num_layers):
        res = (layer_idx + 5) // 2
        sd[f"noises.noise_{layer_idx}"] = randn(1, 1, 2 ** res, 2 ** res)
    return sd


def make_wplus(batch, n_latent, style_dim=512, seed=2):
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.standard_normal((batch, n_latent, style_dim), dtype=np.float32))


def make_z(batch, style_dim=512, seed=2):
    rng = np.random.Generator(np.random.PCG64(seed + 17))
    return torch.from_numpy(rng.standard_normal((batch, style_dim), dtype=np.float32))


def make_mask(batch, h, seed=3, binary=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    m = rng.random((batch, 1, h, h), dtype=np.float32)
    if binary:
        m = (m > 0.5).astype(np.float32)
    return torch.from_numpy(m)


def make_tensor(shape, seed, scale=1.0):
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((scale * rng.standard_normal(shape)).astype(np.float32))
