"""CPU: the oracle against the reference's outputs at the BASELINE configurations (tests/golden/fullsize.npz,
written by oracle/make_fullsize_golden.py from the UNMODIFIED reference): 256^2 batch 4, 1024^2 W+ forward, the
LevelsMapper edit blended at layer 13, and the 1024^2 autograd gradients."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, max_abs
from oracle import make_fullsize_golden as mfg
from oracle import make_golden as mg
from oracle import mapper_oracle as mo
from oracle import stylegan2_oracle as orc
from oracle import synth


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "fullsize.npz")))


@pytest.fixture(scope="module")
def sd1024():
    torch.set_num_threads(os.cpu_count() or 1)
    return synth.make_state_dict(1024, seed=0, perturbed=True, channel_multiplier=2)


def test_cfg1_oracle_matches_reference(golden):
    sd = synth.make_state_dict(256, seed=0, perturbed=True, channel_multiplier=2)
    wplus = synth.make_wplus(4, 14, seed=2)
    with torch.no_grad():
        img, _ = orc.generator_forward_ref(sd, [wplus], 256, input_is_latent=True)
    assert max_abs(mfg.grid(img), golden["cfg1/grid"]) <= 2e-5
    np.testing.assert_allclose(mfg.image_stats(img)[..., 1:], golden["cfg1/stats"][..., 1:], rtol=1e-6)


def test_cfg2_cfg3_oracle_matches_reference(golden, sd1024):
    sd = sd1024
    np.testing.assert_allclose(mg.sd_checksum(sd), golden["sd_checksum"], rtol=1e-12)
    wplus = synth.make_wplus(2, 18, seed=2)
    with torch.no_grad():
        img, _, _, feats = orc.generator_forward_ref(sd, [wplus], 1024, input_is_latent=True, return_features=True)
        assert max_abs(mfg.grid(img), golden["cfg2/grid"]) <= 2e-5
        np.testing.assert_allclose(mfg.image_stats(img)[..., 1:], golden["cfg2/stats"][..., 1:], rtol=1e-6)
        np.testing.assert_allclose(np.stack([mg.stats(f) for f in feats])[:, 1:], golden["cfg2/feat_stats"][:, 1:], rtol=1e-5)
        # cfg3: the numpy mapper oracle reproduces the reference's w_hat; the blended forward its image
        w_hat = wplus + 0.1 * torch.from_numpy(mo.levels_mapper(wplus.numpy(), mo.mapper_state()))
        want = golden["cfg3/w_hat"]
        assert max_abs(w_hat, want) <= 1e-5 * float(np.abs(want).max())
        w_ref = torch.from_numpy(want)
        _, _, styles = orc.generator_forward_ref(sd, [w_ref], 1024, input_is_latent=True, return_latents=True)
        img3, _, _, _ = orc.generator_forward_ref(sd, [styles], 1024, input_is_stylespace=True, return_features=True,
                                                  attention_layer=13, attention_map=synth.make_mask(2, 64, seed=3),
                                                  feature_map=feats)
    assert max_abs(mfg.grid(img3), golden["cfg3/grid"]) <= 2e-5
    np.testing.assert_allclose(mfg.image_stats(img3)[..., 1:], golden["cfg3/stats"][..., 1:], rtol=1e-6)


def test_cfg4_oracle_gradient_matches_reference_autograd(golden, sd1024):
    """dL/dW+ at 1024^2 by the oracle's fp32 autograd: as close to the reference's fp64 gradient as the
    reference's own fp32 run (the two share every library op, so they agree far inside that)."""
    wp = synth.make_wplus(2, 18, seed=2)[:1].clone().requires_grad_(True)
    img, _ = orc.generator_forward_ref(sd1024, [wp], 1024, input_is_latent=True)
    (img * mfg.upstream_grad((1, 3, 1024, 1024))).sum().backward()
    g32, g64 = golden["cfg4/grad_wplus_f32"], golden["cfg4/grad_wplus_f64"]
    scale = float(np.abs(g64).max())
    assert max_abs(wp.grad, g32) <= 1e-3 * scale
    assert max_abs(wp.grad, g64) <= 2 * max_abs(g32, g64) + 1e-4 * scale
