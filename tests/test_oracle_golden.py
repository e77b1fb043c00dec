"""Pins oracle/stylegan2_oracle.py against the committed outputs of the reference's own modules
(tests/golden/*.npz, produced by oracle/make_golden.py in the build container), and against
the live reference when /root/reference is present.  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import max_abs
from oracle import make_golden as mg
from oracle import stylegan2_oracle as orc
from oracle import synth

TOL = 2e-5   # fp32 reorder noise of the reference itself is ~1e-5 on |img|~10


def test_upfirdn2d_oracle_matches_reference_golden(golden_ops):
    for i, (shape, kspec, *geom) in enumerate(mg.UPFIRDN_CASES):
        x = synth.make_tensor(shape, 100 + i)
        k = mg.make_kernel_spec(kspec)
        ref = golden_ops[f"upfirdn_{i}"]
        got = orc.upfirdn2d_native_ref(x, k, *geom)
        assert tuple(got.shape) == ref.shape
        assert max_abs(got, ref) <= 1e-6, (i, geom)
        direct = orc.upfirdn2d_direct(x.double(), k.double(), *geom)
        assert max_abs(direct, ref) <= 2e-6, (i, geom)


def test_fused_leaky_relu_oracle_matches_reference_golden(golden_ops):
    for i, (shape, cdim) in enumerate([((5, 7), 7), ((3, 4, 6), 6), ((2, 5, 4, 3), 5)]):
        x = synth.make_tensor(shape, 200 + i)
        b = synth.make_tensor((cdim,), 210 + i)
        assert max_abs(orc.fused_leaky_relu_ref(x, b), golden_ops[f"flrelu_{i}"]) == 0.0
        assert max_abs(orc.fused_leaky_relu_ref(x, b, 0.1, 1.5), golden_ops[f"flrelu_{i}_args"]) == 0.0


MODCONV_CASES = [(8, 6, 3, False, True, 7), (8, 6, 3, True, True, 5), (8, 3, 1, False, False, 6),
                 (16, 16, 3, True, True, 4)]


def modconv_case_inputs(i):
    cin, cout, k, up, demod, h = MODCONV_CASES[i]
    weight = synth.make_tensor((1, cout, cin, k, k), 300 + i)
    mod_w = synth.make_tensor((cin, 12), 310 + i)
    mod_b = 1 + synth.make_tensor((cin,), 320 + i, 0.1)
    x = synth.make_tensor((2, cin, h, h), 330 + i)
    w = synth.make_tensor((2, 12), 340 + i)
    return weight, mod_w, mod_b, x, w


def test_modconv_oracle_matches_reference_golden(golden_ops):
    blur = synth.blur_kernel_2d(gain=4.0)
    for i, (cin, cout, k, up, demod, h) in enumerate(MODCONV_CASES):
        weight, mod_w, mod_b, x, w = modconv_case_inputs(i)
        y, s = orc.modulated_conv2d_ref(x, w, weight, mod_w, mod_b, demodulate=demod, upsample=up,
                                        blur_kernel=blur)
        assert max_abs(y, golden_ops[f"modconv_{i}_y"]) <= 1e-5
        assert max_abs(s, golden_ops[f"modconv_{i}_s"]) <= 1e-6
        y2, _ = orc.modulated_conv2d_ref(x, torch.from_numpy(golden_ops[f"modconv_{i}_s"]) * 1.1, weight,
                                         mod_w, mod_b, demodulate=demod, upsample=up, blur_kernel=blur,
                                         input_is_stylespace=True)
        assert max_abs(y2, golden_ops[f"modconv_{i}_y_ss"]) <= 1e-5


@pytest.fixture(scope="module")
def g32_state():
    sd = synth.make_state_dict(32, seed=0, perturbed=True)
    wplus = synth.make_wplus(2, 8, seed=2)
    return sd, wplus


def test_synth_state_is_reproducible(golden_g32, g32_state):
    sd, _ = g32_state
    np.testing.assert_allclose(mg.sd_checksum(sd), golden_g32["sd_checksum"], rtol=1e-12)


def test_generator_oracle_matches_reference_golden(golden_g32, g32_state):
    sd, wplus = g32_state
    g = golden_g32
    img, latent, styles, feats = orc.generator_forward_ref(sd, [wplus], 32, input_is_latent=True,
                                                           return_features=True)
    assert max_abs(img, g["img_wplus"]) <= TOL
    assert len(feats) == 11 and len(styles) == 11
    for i, f in enumerate(feats):
        assert max_abs(mg.sub(f), g[f"feat_{i}_sub"]) <= TOL
    for i, s in enumerate(styles):
        assert max_abs(s, g[f"style_{i}"]) <= 1e-6
    ref_styles = [torch.from_numpy(g[f"style_{i}"]) for i in range(11)]
    img_ss, none = orc.generator_forward_ref(sd, [ref_styles], 32, input_is_stylespace=True)
    assert none is None
    assert max_abs(img_ss, g["img_stylespace"]) <= TOL
    edited = [s * (1 + 0.05 * synth.make_tensor(tuple(s.shape), 500 + i)) for i, s in enumerate(ref_styles)]
    img_ed, _ = orc.generator_forward_ref(sd, [edited], 32, input_is_stylespace=True)
    assert max_abs(img_ed, g["img_edited"]) <= TOL
    for tag, layer, msize, binary in [("a", 7, 16, False), ("b", 6, 8, False), ("c", 9, 12, True),
                                      ("d", 1, 4, False)]:
        mask = synth.make_mask(2, msize, seed=3 + layer, binary=binary)
        img_b, _, _, feats_b = orc.generator_forward_ref(
            sd, [edited], 32, input_is_stylespace=True, return_features=True, attention_layer=layer,
            attention_map=mask, feature_map=feats)
        assert max_abs(img_b, g[f"img_blend_{tag}"]) <= TOL, tag
        got = np.stack([mg.stats(f) for f in feats_b])
        np.testing.assert_allclose(got, g[f"blend_{tag}_feat_stats"], rtol=2e-5, atol=1e-3)
    mask = synth.make_mask(2, 16, seed=10)
    w_ed = wplus + 0.1 * synth.make_tensor(tuple(wplus.shape), 600)
    img_bw, _, _, _ = orc.generator_forward_ref(sd, [w_ed], 32, input_is_latent=True, return_features=True,
                                                attention_layer=7, attention_map=mask, feature_map=feats)
    assert max_abs(img_bw, g["img_blend_wplus"]) <= TOL


def test_generator_oracle_z_paths(golden_g32, g32_state):
    sd, wplus = g32_state
    g = golden_g32
    z, z2 = synth.make_z(2, seed=2), synth.make_z(2, seed=3)
    mean_w = orc.mapping_ref(sd, synth.make_z(64, seed=9)).mean(0, keepdim=True)
    assert max_abs(mean_w, g["mean_w"]) <= 1e-5
    img, lat, sv = orc.generator_forward_ref(sd, [z], 32, truncation=0.7, truncation_latent=mean_w,
                                             return_latents=True)
    assert max_abs(lat, g["latent_z_trunc"]) <= 1e-5
    assert max_abs(img, g["img_z_trunc"]) <= TOL
    img_mix, _ = orc.generator_forward_ref(sd, [z, z2], 32, inject_index=3)
    assert max_abs(img_mix, g["img_z_mix"]) <= TOL
    noise = [synth.make_tensor((1, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)), 700 + i) for i in range(7)]
    img_n, _ = orc.generator_forward_ref(sd, [wplus], 32, input_is_latent=True, noise=noise)
    assert max_abs(img_n, g["img_noise_list"]) <= TOL


def test_generator_oracle_gradients(golden_g32, g32_state):
    sd, wplus = g32_state
    g = golden_g32
    upstream = synth.make_tensor((2, 3, 32, 32), 4) / (2 * 3 * 32 * 32)
    wp = wplus.clone().requires_grad_(True)
    img, _ = orc.generator_forward_ref(sd, [wp], 32, input_is_latent=True)
    (img * upstream).sum().backward()
    scale = np.abs(g["grad_wplus"]).max()
    assert max_abs(wp.grad, g["grad_wplus"]) <= 1e-4 * scale


def test_generator128_oracle_matches_reference_golden(golden_g128):
    sd = synth.make_state_dict(128, channel_multiplier=1, seed=5, perturbed=True)
    np.testing.assert_allclose(mg.sd_checksum(sd), golden_g128["sd_checksum"], rtol=1e-12)
    wplus = synth.make_wplus(1, 12, seed=6)
    img, _, _, feats = orc.generator_forward_ref(sd, [wplus], 128, input_is_latent=True, return_features=True)
    assert max_abs(img, golden_g128["img_wplus"]) <= TOL
    got = np.stack([mg.stats(f) for f in feats])
    np.testing.assert_allclose(got, golden_g128["feat_stats"], rtol=2e-5, atol=1e-3)


@pytest.mark.skipif(not os.path.isdir("/root/reference/models/stylegan2"), reason="reference tree absent")
def test_oracle_against_live_reference():
    """Build-container only: run the reference module and the oracle side by side on fresh seeds."""
    saved_cuda = torch.Tensor.cuda
    am, up_native, flrelu = mg.import_reference()
    try:
        sd = synth.make_state_dict(16, seed=11, perturbed=True)
        gen = am.Generator(16, 512, 8)
        gen.load_state_dict(sd, strict=True)
        wplus = synth.make_wplus(3, 6, seed=12)
        with torch.no_grad():
            ref_img, _, ref_styles, ref_feats = gen([wplus], input_is_latent=True, randomize_noise=False,
                                                    return_features=True)
        img, _, styles, feats = orc.generator_forward_ref(sd, [wplus], 16, input_is_latent=True,
                                                          return_features=True)
        assert max_abs(img, ref_img) <= TOL
        for a, b in zip(feats, ref_feats):
            assert max_abs(a, b) <= TOL
        x = synth.make_tensor((2, 3, 13, 11), 77)
        k = synth.make_tensor((4, 3), 78)
        assert max_abs(orc.upfirdn2d_native_ref(x, k, 2, 1, 1, 2, 1, 2, 0, 3),
                       up_native(x, k, 2, 1, 1, 2, 1, 2, 0, 3)) <= 1e-6
    finally:
        torch.Tensor.cuda = saved_cuda
        for p in ("/root/reference", "/root/reference/attention"):
            while p in sys.path:
                sys.path.remove(p)
