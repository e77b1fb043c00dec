"""bf16 tensor-core path (tcgen05 modulated convolution + channels-last kernels) against the CPU
oracle.  Tolerances are the north-star's bf16 bar: <= 2e-2 max-abs and >= 45 dB PSNR (peak-to-peak
2) after both images are divided by c = max|reference| (random-init images are not in [-1,1],
SURVEY.md section 0.6); per-op checks use the same normalisation with a tighter 1e-2 bound."""
import numpy as np
import pytest
import torch

import where2edit_b200 as w2e
from oracle import stylegan2_oracle as orc
from oracle import synth
from where2edit_b200 import _native as N
from where2edit_b200 import engine as E
from where2edit_b200 import functional as K

from conftest import max_abs, psnr_db

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def norm_err(test, ref):
    c = float(torch.as_tensor(ref).abs().max())
    return max_abs(torch.as_tensor(test).double() / c, torch.as_tensor(ref).double() / c)


class _Layer:
    """A StyledConv of the product package driven directly through the engine's launchers."""

    def __init__(self, cin, cout, up, seed):
        self.m = w2e.StyledConv(cin, cout, 3, 16, upsample=up)
        self.weight = synth.make_tensor((1, cout, cin, 3, 3), seed)
        with torch.no_grad():
            self.m.conv.weight.copy_(self.weight)
            self.m.noise.weight.fill_(0.3)
            self.m.activate.bias.copy_(0.2 * synth.make_tensor((cout,), seed + 1))
        self.m = self.m.to(DEV)


def run_layer(eng, layer, x, s, noise, nxt, act=True):
    """x fp32 NCHW, s [B,Cin] -> (out NCHW fp32, out_mod NCHW fp32) through the bf16 kernels."""
    m = layer.m
    b, cin, h, w = x.shape
    pw = eng._tc_weight(m.conv)
    d = K.demod_coefficients(s.to(DEV), pw.wsq)
    xs = eng._to_nhwc(x.to(DEV), s.to(DEV).contiguous(), b)
    nz = noise.to(DEV) if noise is not None else None
    nw = m.noise.weight.detach()
    bias = m.activate.bias.detach()
    nxt_d = nxt.to(DEV).contiguous() if nxt is not None else None
    if not m.conv.upsample:
        out, out_mod = eng._conv2(xs, pw, d, nz, nw, bias, nxt_d, True, nxt is not None, False,
                                  N.ACT_LRELU if act else N.ACT_NONE)
    else:
        z, _ = eng._conv2(xs, pw, d, None, None, None, None, True, False, True, N.ACT_NONE)
        out, out_mod = eng._blur(z, m.conv.blur.kernel, m.conv.blur.pad, bias, nz, nw, nxt_d, True, nxt is not None,
                                 (2 * h, 2 * w))
    eng.assert_ok()
    return eng._to_nchw(out).cpu(), (eng._to_nchw(out_mod).cpu() if out_mod is not None else None)


@pytest.fixture(scope="module")
def eng():
    gen = w2e.Generator(8, 512, 1).to(DEV)
    return E.SynthesisEngine(gen)


@pytest.mark.parametrize("b,cin,cout,h,up", [
    (2, 64, 64, 16, False),     # BK 64, one pixel tile per image row block
    (1, 32, 32, 40, False),     # BK 32 (64-byte swizzle), ragged 40x40 grid
    (3, 128, 48, 8, False),     # two images per tile, BN 16
    (9, 64, 32, 4, False),      # eight images per tile, ragged batch
    (1, 512, 512, 8, False),    # BN 256 x 2 column tiles, 72 pipeline iterations
    (2, 256, 128, 16, False),   # BN 128
    (2, 64, 64, 8, True),       # transposed x2: four parity classes + blur
    (1, 128, 64, 20, True),
    (5, 512, 512, 4, True),
    (1, 64, 32, 36, True),      # resident weights, ragged 37x37 class grid, two accumulator buffers
    (2, 32, 32, 72, False),     # many tiles per CTA slot: accumulator double-buffering wraps
])
def test_styled_conv_tc_matches_oracle(eng, b, cin, cout, h, up):
    layer = _Layer(cin, cout, up, 31)
    x = synth.make_tensor((b, cin, h, h), 32)
    s = 1 + 0.3 * synth.make_tensor((b, cin), 33)
    oh = 2 * h if up else h
    noise = synth.make_tensor((1, 1, oh, oh), 34)
    nxt = 1 + 0.3 * synth.make_tensor((b, cout), 35)
    got, got_mod = run_layer(eng, layer, x, s, noise, nxt)
    blur = synth.blur_kernel_2d(gain=4.0)
    ref, _ = orc.modulated_conv2d_ref(x.double(), s.double().reshape(b, 1, cin, 1, 1), layer.weight.double(), None,
                                      None, True, up, blur.double(), input_is_stylespace=True)
    ref = ref + 0.3 * noise.double()
    ref = orc.fused_leaky_relu_ref(ref, layer.m.activate.bias.detach().cpu().double())
    assert got.shape == ref.shape
    assert norm_err(got, ref) <= 1e-2
    assert norm_err(got_mod, ref * nxt.double().reshape(b, cout, 1, 1)) <= 1e-2


def test_torgb_nhwc_and_layouts(eng):
    b, c, h = 2, 64, 16
    x = synth.make_tensor((b, c, h, h), 41)
    s = 1 + 0.3 * synth.make_tensor((b, c), 42)
    skip = synth.make_tensor((b, 3, h // 2, h // 2), 43)
    m = w2e.ToRGB(c, 16)
    w = synth.make_tensor((1, 3, c, 1, 1), 44)
    with torch.no_grad():
        m.conv.weight.copy_(w)
        m.bias.copy_(0.1 * synth.make_tensor((1, 3, 1, 1), 45))
    m = m.to(DEV)
    xh = eng._to_nhwc(x.to(DEV), None, b)
    # layout round trip is exact up to the bf16 rounding of the input
    back = eng._to_nchw(xh).cpu()
    assert torch.equal(back, x.to(torch.bfloat16).float())
    rgb = eng._torgb(xh, m, s.to(DEV).contiguous(), skip.to(DEV)).cpu()
    sd = {"p.conv.weight": w, "p.bias": m.bias.detach().cpu(), "p.upsample.kernel": synth.blur_kernel_2d(gain=4.0)}
    ref, _ = orc._to_rgb({k: v.double() for k, v in sd.items()} | {"p.conv.modulation.weight": None,
                                                                   "p.conv.modulation.bias": None},
                         "p", x.to(torch.bfloat16).double(), s.double().reshape(b, 1, c, 1, 1), skip.double(), True)
    assert norm_err(rgb, ref) <= 1e-5
    rgb0 = eng._torgb(xh, m, s.to(DEV).contiguous(), None).cpu()
    ref0, _ = orc._to_rgb({k: v.double() for k, v in sd.items()} | {"p.conv.modulation.weight": None,
                                                                    "p.conv.modulation.bias": None},
                          "p", x.to(torch.bfloat16).double(), s.double().reshape(b, 1, c, 1, 1), None, True)
    assert norm_err(rgb0, ref0) <= 1e-5


@pytest.mark.parametrize("b,c,h,w", [(2, 32, 8, 8), (3, 64, 37, 37), (1, 128, 16, 24), (2, 512, 9, 7), (2, 32, 5, 5),
                                     (1, 48, 16, 16), (2, 256, 64, 64)])
def test_channels_last_to_nchw_is_an_exact_transpose(eng, b, c, h, w):
    """w2e_nhwc_to_nchw_f32 (captured feature maps go back to the caller through it): both kernels -- 64 x 64 tiles with
    16-byte loads, 32 x 32 tiles for the other shapes -- against a torch permute, ragged pixel counts included"""
    x = synth.make_tensor((b, h, w, c), 31).to(torch.bfloat16).to(DEV)
    got = eng._to_nchw(x)
    assert got.dtype == torch.float32 and torch.equal(got, x.float().permute(0, 3, 1, 2).contiguous())


def _check_image(img, ref):
    c = float(ref.abs().max())
    assert max_abs(img.double() / c, ref.double() / c) <= 2e-2
    assert psnr_db(img.double() / c, ref.double() / c, peak=2.0) >= 45.0


@pytest.mark.parametrize("size,cm,batch", [(32, 2, 3), (128, 2, 2)])
def test_generator_bf16_matches_oracle(size, cm, batch):
    sd = synth.make_state_dict(size, seed=0, perturbed=True, channel_multiplier=cm)
    gen = w2e.Generator(size, 512, 8, channel_multiplier=cm, precision="bf16")
    gen.load_state_dict(sd)
    gen = gen.to(DEV).eval()
    wplus = synth.make_wplus(batch, gen.n_latent, seed=2)
    ref, _, ref_styles, ref_feats = orc.generator_forward_ref(sd, [wplus], size, input_is_latent=True,
                                                              return_features=True)
    with torch.no_grad():
        img, latent, styles, feats = gen([wplus.to(DEV)], input_is_latent=True, randomize_noise=False,
                                         return_features=True)
    gen._engine.assert_ok()
    assert img.dtype == torch.float32 and tuple(img.shape) == tuple(ref.shape)
    _check_image(img.cpu(), ref)
    assert len(feats) == len(ref_feats) and len(styles) == len(ref_styles)
    for i, (f, r) in enumerate(zip(feats, ref_feats)):
        assert tuple(f.shape) == tuple(r.shape), i
        assert norm_err(f.cpu(), r) <= 3e-2, i
    for s, r in zip(styles, ref_styles):
        assert max_abs(s.cpu(), r) <= 1e-4 * max(1.0, float(r.abs().max()))
    # stylespace input reproduces the W+ image bit-exactly (same kernels, same styles)
    with torch.no_grad():
        img_ss, _ = gen([styles], input_is_stylespace=True, randomize_noise=False)
    assert torch.equal(img_ss, img)


def test_generator_bf16_blend_path():
    size = 32
    sd = synth.make_state_dict(size, seed=0, perturbed=True)
    gen = w2e.Generator(size, 512, 8, precision="bf16")
    gen.load_state_dict(sd)
    gen = gen.to(DEV).eval()
    wplus = synth.make_wplus(2, gen.n_latent, seed=2)
    _, _, ref_styles, ref_feats = orc.generator_forward_ref(sd, [wplus], size, input_is_latent=True,
                                                            return_features=True)
    edited = [s * (1 + 0.05 * synth.make_tensor(tuple(s.shape), 500 + i)) for i, s in enumerate(ref_styles)]
    mask = synth.make_mask(2, 16, seed=10)
    ref, _, _, ref_new = orc.generator_forward_ref(sd, [edited], size, input_is_stylespace=True, return_features=True,
                                                   attention_layer=7, attention_map=mask, feature_map=ref_feats)
    with torch.no_grad():
        img, _, _, new = gen([[s.to(DEV) for s in edited]], input_is_stylespace=True, randomize_noise=False,
                             return_features=True, attention_layer=7, attention_map=mask.to(DEV),
                             feature_map=[f.to(DEV) for f in ref_feats])
    gen._engine.assert_ok()
    _check_image(img.cpu(), ref)
    assert norm_err(new[6].cpu(), ref_new[6]) <= 3e-2
    # a mask of ones leaves the edited image, a mask of zeros restores the original features downstream
    with torch.no_grad():
        full, _ = gen([[s.to(DEV) for s in edited]], input_is_stylespace=True, randomize_noise=False)
        ones, _ = gen([[s.to(DEV) for s in edited]], input_is_stylespace=True, randomize_noise=False,
                      attention_layer=7, attention_map=torch.ones(2, 1, 16, 16, device=DEV),
                      feature_map=[f.to(DEV) for f in ref_feats])
    assert norm_err(ones.cpu(), full.cpu()) <= 2e-2


def test_bf16_batch_invariance():
    """no cross-sample term and no split-K: a batch of 3 equals three batches of 1, bitwise"""
    sd = synth.make_state_dict(32, seed=3, perturbed=True)
    gen = w2e.Generator(32, 512, 8, precision="bf16")
    gen.load_state_dict(sd)
    gen = gen.to(DEV).eval()
    wplus = synth.make_wplus(3, gen.n_latent, seed=4).to(DEV)
    with torch.no_grad():
        _, _, styles = gen([wplus], input_is_latent=True, randomize_noise=False, return_latents=True)
        full, _ = gen([styles], input_is_stylespace=True, randomize_noise=False)
        singles = torch.cat([gen([[s[i:i + 1] for s in styles]], input_is_stylespace=True,
                                 randomize_noise=False)[0] for i in range(3)])
    gen._engine.assert_ok()
    assert torch.equal(full, singles)


@pytest.mark.parametrize("b,cin,cout,h,with_skip,want_out,want_mod", [
    (2, 64, 64, 32, True, True, True), (1, 32, 32, 40, False, False, True), (2, 128, 256, 24, True, False, True),
    (1, 64, 32, 72, True, True, True),
    (2, 32, 32, 136, True, False, False),   # RGB-only last layer: 512-pixel tiles, two pixels per epilogue thread
    (2, 32, 32, 144, True, False, False),   # RGB-only last layer on pixel pairs (w2e_modconv_tc2_rgb_pair), ragged rows
    (1, 32, 32, 48, False, False, False),   # pixel pairs, no skip image
    (3, 32, 32, 80, True, False, False),    # pixel pairs, 5 x 3 tiles per image
    (2, 32, 32, 112, True, True, False),    # pixel pairs with the activation kept (the training forward's last layer)
    (1, 64, 64, 64, True, False, True),     # one staging slot per epilogue half
    (3, 128, 128, 48, True, False, True),   # weight ring + staged epilogue, 64-channel units
    (2, 512, 512, 24, True, True, True),    # two channel blocks adding their partial ToRGB sums
])
def test_conv_with_fused_torgb_matches_oracle(eng, b, cin, cout, h, with_skip, want_out, want_mod):
    """StyledConv + ToRGB in one launch (w2e_modconv_tc2_rgb) vs the oracle's two modules"""
    layer = _Layer(cin, cout, False, 51)
    rgbm = w2e.ToRGB(cout, 16)
    wr = synth.make_tensor((1, 3, cout, 1, 1), 52)
    with torch.no_grad():
        rgbm.conv.weight.copy_(wr)
        rgbm.bias.copy_(0.1 * synth.make_tensor((1, 3, 1, 1), 53))
    rgbm = rgbm.to(DEV)
    x = synth.make_tensor((b, cin, h, h), 54)
    s = 1 + 0.3 * synth.make_tensor((b, cin), 55)
    s_rgb = 1 + 0.3 * synth.make_tensor((b, cout), 56)
    nxt = 1 + 0.3 * synth.make_tensor((b, cout), 57)
    noise = synth.make_tensor((1, 1, h, h), 58)
    skip = synth.make_tensor((b, 3, h // 2, h // 2), 59) if with_skip else None
    m = layer.m
    pw = eng._tc_weight(m.conv)
    d = K.demod_coefficients(s.to(DEV), pw.wsq)
    xs = eng._to_nhwc(x.to(DEV), s.to(DEV).contiguous(), b)
    out, out_mod, rgb = eng._conv2_rgb(xs, pw, d, noise.to(DEV), m.noise.weight.detach(), m.activate.bias.detach(),
                                       nxt.to(DEV).contiguous() if want_mod else None, want_out, want_mod, rgbm,
                                       s_rgb.to(DEV).contiguous(), skip.to(DEV) if skip is not None else None)
    eng.assert_ok()
    ref, _ = orc.modulated_conv2d_ref(x.double(), s.double().reshape(b, 1, cin, 1, 1), layer.weight.double(), None,
                                      None, True, False, None, input_is_stylespace=True)
    ref = orc.fused_leaky_relu_ref(ref + 0.3 * noise.double(), m.activate.bias.detach().cpu().double())
    sd = {"p.conv.weight": wr.double(), "p.bias": rgbm.bias.detach().cpu().double(),
          "p.upsample.kernel": synth.blur_kernel_2d(gain=4.0).double(), "p.conv.modulation.weight": None,
          "p.conv.modulation.bias": None}
    ref_rgb, _ = orc._to_rgb(sd, "p", ref, s_rgb.double().reshape(b, 1, cout, 1, 1),
                             skip.double() if skip is not None else None, True)
    assert norm_err(rgb.cpu(), ref_rgb) <= 1e-2
    if want_mod:
        assert norm_err(eng._to_nchw(out_mod).cpu(), ref * nxt.double().reshape(b, cout, 1, 1)) <= 1e-2
    else:
        assert out_mod is None
    if want_out:
        assert norm_err(eng._to_nchw(out).cpu(), ref) <= 1e-2
    else:
        assert out is None


@pytest.mark.parametrize("b,h,rgb_dtype", [(2, 64, torch.float32), (1, 144, torch.bfloat16)])
def test_pixel_pair_last_layer_agrees_with_the_pixel_kernel(eng, b, h, rgb_dtype):
    """w2e_modconv_tc2_rgb_pair (N = 64 MMAs on pixel pairs) is the same arithmetic as w2e_modconv_tc2_rgb on the
    32-channel RGB-only layer up to the fp32 summation order inside the tensor core and the ToRGB dot product"""
    layer = _Layer(32, 32, False, 61)
    rgbm = w2e.ToRGB(32, 16)
    with torch.no_grad():
        rgbm.conv.weight.copy_(synth.make_tensor((1, 3, 32, 1, 1), 62))
        rgbm.bias.copy_(0.1 * synth.make_tensor((1, 3, 1, 1), 63))
    rgbm = rgbm.to(DEV)
    x = synth.make_tensor((b, 32, h, h), 64)
    s = 1 + 0.3 * synth.make_tensor((b, 32), 65)
    s_rgb = (1 + 0.3 * synth.make_tensor((b, 32), 66)).to(DEV).contiguous()
    noise = synth.make_tensor((1, 1, h, h), 67).to(DEV)
    skip = synth.make_tensor((b, 3, h // 2, h // 2), 68).to(DEV)
    m = layer.m
    pw = eng._tc_weight(m.conv)
    d = K.demod_coefficients(s.to(DEV), pw.wsq)
    xs = eng._to_nhwc(x.to(DEV), s.to(DEV).contiguous(), b)
    got = {}
    try:
        for pair in (True, False):
            eng.pair_mode = pair
            before = N.STATS.launches.get("w2e_modconv_tc2_rgb_pair", 0)
            _, _, rgb = eng._conv2_rgb(xs, pw, d, noise, m.noise.weight.detach(), m.activate.bias.detach(), None, False,
                                       False, rgbm, s_rgb, skip, rgb_dtype=rgb_dtype)
            eng.assert_ok()
            assert (N.STATS.launches.get("w2e_modconv_tc2_rgb_pair", 0) - before) == (1 if pair else 0)
            assert rgb.dtype == rgb_dtype
            got[pair] = rgb.float().cpu()
    finally:
        eng.pair_mode = True
    tol = 1e-5 if rgb_dtype == torch.float32 else 8e-3   # bf16 image: one rounding step of the stored value
    assert norm_err(got[True], got[False]) <= tol


@pytest.mark.parametrize("b,cin,cout,h,up", [(2, 64, 32, 40, True), (1, 128, 64, 24, True), (2, 256, 128, 20, True),
                                             (2, 64, 64, 48, False), (2, 32, 32, 72, False)])
def test_staged_and_direct_epilogues_agree_bitwise(eng, b, cin, cout, h, up):
    """the TMA-store epilogue (two MMA issuers, edge-tile tap masking, unit split) and the direct-store
    epilogue are two schedules of the same arithmetic: results must be bit-identical"""
    layer = _Layer(cin, cout, up, 71)
    x = synth.make_tensor((b, cin, h, h), 72)
    s = 1 + 0.3 * synth.make_tensor((b, cin), 73)
    noise = synth.make_tensor((1, 1, 2 * h if up else h, 2 * h if up else h), 74)
    nxt = 1 + 0.3 * synth.make_tensor((b, cout), 75)
    outs = []
    try:
        # (ts_mode, flags, cluster_log2): the per-call w2e_tc2_config switches (include/w2e.h)
        for ts_mode, flags, clus in [(1, 0, 0), (0, 0, 0), (1, 1, 0), (1, 2, 0), (1, 3, 0), (1, 0, 1), (0, 0, 1), (1, 8, 0),
                                     (1, 16, 0), (1, 32, 0), (1, 64, 0), (1, 66, 0), (1, 256, 0), (1, 258, 0), (1, 257, 0),
                                     (1, 512, 0)]:
            eng.tc2_cfg = N.tc2_config(ts_mode=ts_mode, flags=flags, cluster_log2=clus)
            outs.append(run_layer(eng, layer, x, s, noise, nxt))
    finally:
        eng.tc2_cfg = None
    for got, got_mod in outs[1:]:
        assert torch.equal(got, outs[0][0])
        assert torch.equal(got_mod, outs[0][1])


def test_style_plan_matches_per_layer_modules():
    """w2e_style_mod_all / w2e_style_demod_all (two launches for all 26 styled layers) against the per-layer
    EqualLinear modulation and demodulation of the module path"""
    sd = synth.make_state_dict(64, seed=5, perturbed=True)
    gen = w2e.Generator(64, 512, 8, precision="bf16")
    gen.load_state_dict(sd)
    gen = gen.to(DEV).eval()
    eng_ = gen._engine if getattr(gen, "_engine", None) is not None else E.SynthesisEngine(gen)
    wplus = synth.make_wplus(5, gen.n_latent, seed=6).to(DEV)
    layers = gen.styled_layers()
    rows = gen.latent_rows(False)
    with torch.no_grad():
        styles, demods = eng_._styles_and_demods(layers, rows, wplus, False, 5, torch.device(DEV))
        for (module, kind), row, s, d in zip(layers, rows, styles, demods):
            ref = module.conv.styles(wplus[:, row], False).double()
            assert max_abs(s.cpu().double(), ref.cpu()) <= 1e-5 * max(1.0, float(ref.abs().max()))
            if kind == "rgb":
                assert d is None
            else:
                pw = module.conv.packed()
                dref = torch.rsqrt((ref * ref) @ pw.wsq.double().t() + 1e-8)
                assert max_abs(d.cpu().double(), dref.cpu()) <= 1e-5 * float(dref.abs().max())
        # stylespace entry: the concatenated styles go through the same demodulation launch
        ss = [s.reshape(5, 1, -1, 1, 1) for s in styles]
        styles2, demods2 = eng_._styles_and_demods(layers, gen.latent_rows(True), ss, True, 5, torch.device(DEV))
        for a, b_ in zip(styles, styles2):
            assert torch.equal(a, b_)
        for a, b_ in zip(demods, demods2):
            assert (a is None and b_ is None) or torch.equal(a, b_)


@pytest.mark.parametrize("b,cin,cout,h,want_out", [(2, 64, 32, 40, True), (1, 128, 64, 24, False), (1, 64, 32, 72, False),
                                                   (3, 32, 32, 18, True), (1, 128, 64, 66, True)])
def test_fused_upconv_blur_matches_oracle(eng, b, cin, cout, h, want_out):
    """w2e_modconv_tc2_upblur: conv_transpose x2 + demod + 4x4 blur + noise + bias + lrelu in one launch,
    against the oracle and against the two-kernel path (which rounds the pre-blur tensor to bf16 as well)"""
    layer = _Layer(cin, cout, True, 81)
    x = synth.make_tensor((b, cin, h, h), 82)
    s = 1 + 0.3 * synth.make_tensor((b, cin), 83)
    noise = synth.make_tensor((1, 1, 2 * h, 2 * h), 84)
    nxt = 1 + 0.3 * synth.make_tensor((b, cout), 85)
    m = layer.m
    pw = eng._tc_weight(m.conv)
    d = K.demod_coefficients(s.to(DEV), pw.wsq)
    xs = eng._to_nhwc(x.to(DEV), s.to(DEV).contiguous(), b)
    fused = eng._upblur(xs, pw, d, m.conv.blur.kernel, m.conv.blur.pad, m.activate.bias.detach(), noise.to(DEV),
                        m.noise.weight.detach(), nxt.to(DEV).contiguous(), want_out, True)
    eng.assert_ok()
    assert fused is not None
    out, out_mod = fused
    blur = synth.blur_kernel_2d(gain=4.0)
    ref, _ = orc.modulated_conv2d_ref(x.double(), s.double().reshape(b, 1, cin, 1, 1), layer.weight.double(), None,
                                      None, True, True, blur.double(), input_is_stylespace=True)
    ref = orc.fused_leaky_relu_ref(ref + 0.3 * noise.double(), m.activate.bias.detach().cpu().double())
    assert tuple(out_mod.shape) == (b, 2 * h, 2 * h, cout)
    assert norm_err(eng._to_nchw(out_mod).cpu(), ref * nxt.double().reshape(b, cout, 1, 1)) <= 1e-2
    if want_out:
        assert norm_err(eng._to_nchw(out).cpu(), ref) <= 1e-2
    else:
        assert out is None
    two, two_mod = run_layer(eng, layer, x, s, noise, nxt)
    assert norm_err(eng._to_nchw(out_mod).cpu(), two_mod) <= 1e-2


@pytest.mark.parametrize("b,cin,cout,h,up", [(2, 64, 32, 20, False), (1, 128, 64, 17, True), (2, 32, 32, 36, True),
                                             (1, 512, 256, 8, True)])
def test_tensor_core_dgrad_matches_fp32_engine(b, cin, cout, h, up):
    """autograd path of precision='bf16': conv_dgrad_tc (flipped-tap kernel / four parity-class convolutions for the
    transposed conv) against the exact fp32 conv engine on the same upstream gradient"""
    m = w2e.StyledConv(cin, cout, 3, 16, upsample=up)
    with torch.no_grad():
        m.conv.weight.copy_(synth.make_tensor((1, cout, cin, 3, 3), 91))
    m = m.to(DEV)
    pw = m.conv.packed()
    zh = 2 * h + 1 if up else h
    gy = synth.make_tensor((b, cout, zh, zh), 92).to(DEV)
    d = (1 + 0.3 * synth.make_tensor((b, cout), 93)).to(DEV).contiguous()
    ref = K.conv_dgrad(gy, d, pw, 3, up, (h, h))
    got = K.conv_dgrad_tc(gy, d, pw, up, (h, h))
    K.tc_assert_ok()
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert norm_err(got.cpu(), ref.cpu()) <= 1e-2


@pytest.mark.parametrize("b,cin,cout,h,up", [
    (2, 64, 64, 16, False),      # 128-byte rows (32 fp32 channels per K chunk)
    (1, 48, 32, 40, False),      # 64-byte rows (Cin % 32 != 0), ragged grid
    (1, 512, 512, 8, False),     # 256-column tiles x 2, 16 K chunks
    (2, 256, 128, 20, True),     # transposed x2, weight ring
    (1, 64, 32, 36, True),       # transposed x2, resident weights
])
def test_tf32_modulated_conv_matches_oracle(b, cin, cout, h, up):
    """w2e_modconv_tc2_tf32 (fp32 tensors in HBM, tcgen05 kind::tf32, north star item 1) through
    functional.conv_forward_tc / conv_dgrad_tc against the fp64 oracle: 10-bit-mantissa operands, so ~1e-3 of the
    output's range -- an order of magnitude inside the bf16 kernel's error on the same problem."""
    pw = K.PackedWeight(synth.make_tensor((1, cout, cin, 3, 3), 81).to(DEV), 1 / (cin * 9) ** 0.5, None)
    x = synth.make_tensor((b, cin, h, h), 82)
    s = 1 + 0.3 * synth.make_tensor((b, cin), 83)
    d = K.demod_coefficients(s.to(DEV), pw.wsq)
    got = K.conv_forward_tc(x.to(DEV), s.to(DEV), d, pw, up, "tf32")
    got16 = K.conv_forward_tc(x.to(DEV), s.to(DEV), d, pw, up, "bf16") if cin % 32 == 0 else None
    K.tc_assert_ok()
    w64 = synth.make_tensor((1, cout, cin, 3, 3), 81).double()
    xs = x.double() * s.double()[:, :, None, None]
    wt = w64[0] / (cin * 9) ** 0.5
    if up:
        ref = torch.nn.functional.conv_transpose2d(xs, wt.transpose(0, 1), stride=2)
    else:
        ref = torch.nn.functional.conv2d(xs, wt, padding=1)
    ref = ref * d.double().cpu()[:, :, None, None]
    e32 = norm_err(got.cpu(), ref)
    e16 = norm_err(got16.cpu(), ref) if got16 is not None else 1.0
    assert got.dtype == torch.float32 and tuple(got.shape) == tuple(ref.shape)
    assert e32 <= 2e-3 and e32 < 0.5 * e16, (e32, e16)
    # dgrad on the same kernel (flipped taps / the four parity classes of the transposed convolution)
    gy = synth.make_tensor(tuple(ref.shape), 84)
    gx = K.conv_dgrad_tc(gy.to(DEV), d, pw, up, (h, h), "tf32")
    K.tc_assert_ok()
    xs_ = xs.clone().requires_grad_(True)
    out = (torch.nn.functional.conv_transpose2d(xs_, wt.transpose(0, 1), stride=2) if up
           else torch.nn.functional.conv2d(xs_, wt, padding=1)) * d.double().cpu()[:, :, None, None]
    (out * gy.double()).sum().backward()
    assert norm_err(gx.cpu(), xs_.grad) <= 2e-3


def test_tf32_generator_between_bf16_and_fp32(golden_g128):
    """Generator(precision='tf32') at 128^2 against the reference golden: inside the bf16 tolerance and closer to the
    reference than the bf16 engine; gradients flow through the tf32 dgrad."""
    sd = synth.make_state_dict(128, channel_multiplier=1, seed=5, perturbed=True)
    gen = w2e.Generator(128, 512, 8, channel_multiplier=1, precision="tf32")
    gen.load_state_dict(sd, strict=True)
    gen = gen.to(DEV).eval()
    for p in gen.parameters():
        p.requires_grad_(False)
    wplus = synth.make_wplus(1, 12, seed=6).to(DEV)
    ref = torch.from_numpy(golden_g128["img_wplus"])
    c = float(ref.abs().max())
    with torch.no_grad():
        img, _ = gen([wplus], input_is_latent=True, randomize_noise=False)
        gen.set_precision("bf16")
        img16, _ = gen([wplus], input_is_latent=True, randomize_noise=False)
    gen.assert_ok()
    e32, e16 = max_abs(img.cpu() / c, ref / c), max_abs(img16.float().cpu() / c, ref / c)
    assert e32 <= 5e-3 and e32 < e16 <= 2e-2, (e32, e16)
    gen.set_precision("tf32")
    wp = wplus.clone().requires_grad_(True)
    img, _ = gen([wp], input_is_latent=True, randomize_noise=False)
    img.square().mean().backward()
    gen.assert_ok()
    g_tf32 = wp.grad.double().flatten().cpu()
    gen.set_precision("fp32")
    wp2 = wplus.clone().requires_grad_(True)
    img, _ = gen([wp2], input_is_latent=True, randomize_noise=False)
    img.square().mean().backward()
    g32 = wp2.grad.double().flatten().cpu()
    cos = float(torch.dot(g_tf32, g32) / (g_tf32.norm() * g32.norm()))
    assert cos >= 0.9995, cos
