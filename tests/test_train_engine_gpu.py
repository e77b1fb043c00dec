"""The channels-last bf16 forward + backward engine (train_engine.py; BASELINE config 4) against the reference's autograd
gradients (tests/golden/generator32.npz, written by the unmodified reference) and against this package's exact fp32
path; its kernels (grad_assemble, rowdot, sum4, the strided-view convolution) against plain torch formulas."""
import numpy as np
import pytest
import torch

import where2edit_b200 as w2e
from conftest import max_abs
from oracle import synth
from where2edit_b200 import _native as N
from where2edit_b200 import functional as K

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build(size, cm=2, precision="bf16"):
    gen = w2e.Generator(size, 512, 8, channel_multiplier=cm, precision=precision)
    gen.load_state_dict(synth.make_state_dict(size, seed=0, perturbed=True, channel_multiplier=cm), strict=True)
    gen = gen.to(DEV).eval()
    for p in gen.parameters():
        p.requires_grad_(False)
    return gen


def cos_rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float(torch.dot(a, b) / (a.norm() * b.norm())), float((a - b).norm() / b.norm())


def test_wplus_gradient_matches_reference_autograd(golden_g32):
    gen = build(32)
    wplus = synth.make_wplus(2, 8, seed=2)
    upstream = (synth.make_tensor((2, 3, 32, 32), 4) / (2 * 3 * 32 * 32)).to(DEV)
    wp = wplus.to(DEV).requires_grad_(True)
    img, _ = gen([wp], input_is_latent=True, randomize_noise=False)
    assert img.dtype == torch.float32 and img.grad_fn is not None and "Synthesis" in type(img.grad_fn).__name__
    (img * upstream).sum().backward()
    gen.assert_ok()
    c = float(np.abs(golden_g32["img_wplus"]).max())
    assert max_abs(img.detach().cpu() / c, torch.from_numpy(golden_g32["img_wplus"]) / c) <= 2e-2
    cos, rel = cos_rel(wp.grad, torch.from_numpy(golden_g32["grad_wplus"]))
    assert cos >= 0.99 and rel <= 0.12, (cos, rel)
    # deterministic: a second run is bit-identical
    wp2 = wplus.to(DEV).requires_grad_(True)
    img2, _ = gen([wp2], input_is_latent=True, randomize_noise=False)
    (img2 * upstream).sum().backward()
    assert torch.equal(img2, img) and torch.equal(wp2.grad, wp.grad)


@pytest.mark.parametrize("size,cm", [(64, 2), (256, 1)])
def test_stylespace_gradients_track_the_fp32_path(size, cm):
    """every one of the per-layer style gradients (3x3 convolutions: direct + demodulation term; ToRGB) against the
    exact fp32 kernels, which tests/test_generator_gpu.py pins to the reference's autograd"""
    gen = build(size, cm)
    w = synth.make_wplus(2, gen.n_latent, seed=2).to(DEV)
    g = torch.Generator().manual_seed(4)
    upstream = (torch.randn(2, 3, size, size, generator=g) / (3 * size * size)).to(DEV)
    with torch.no_grad():
        _, _, styles = gen([w], input_is_latent=True, randomize_noise=False, return_latents=True)
    grads = {}
    for mode in ("engine", "fp32"):
        gen.set_precision("bf16" if mode == "engine" else "fp32")
        st = [s.detach().clone().requires_grad_(True) for s in styles]
        img, _ = gen([st], input_is_stylespace=True, randomize_noise=False)
        (img * upstream).sum().backward()
        grads[mode] = [s.grad for s in st]
    gen.set_precision("bf16")
    gen.assert_ok()
    total = cos_rel(torch.cat([x.flatten() for x in grads["engine"]]), torch.cat([x.flatten() for x in grads["fp32"]]))
    assert total[0] >= 0.99 and total[1] <= 0.12, total
    for i, (a, b) in enumerate(zip(grads["engine"], grads["fp32"])):
        assert a.shape == b.shape
        cos, rel = cos_rel(a, b)
        assert cos >= 0.97, (i, cos, rel)
    # batch invariance of the backward: sample 1 alone gives the same gradient rows (the kernels and their reductions
    # are bit-identical per sample; the small [B,C] demodulation algebra goes through library GEMMs whose algorithm
    # may depend on the batch, hence a tolerance of a few fp32 ulps of the gradient's range)
    st1 = [s.detach()[1:2].clone().requires_grad_(True) for s in styles]
    img1, _ = gen([st1], input_is_stylespace=True, randomize_noise=False)
    (img1 * upstream[1:2]).sum().backward()
    for a, b in zip(st1, grads["engine"]):
        assert max_abs(a.grad.cpu(), b[1:2].cpu()) <= 1e-5 * float(b.abs().max())


def test_blended_stylespace_gradients_match_reference_autograd(golden_g32):
    """the reference's training forward (run_attention.py:1245): stylespace codes + region blend at a convolution layer
    and the following ToRGB; gradients w.r.t. the 11 styles AND the attention map against the reference's autograd"""
    g = golden_g32
    gen = build(32)
    wplus = synth.make_wplus(2, 8, seed=2).to(DEV)
    upstream = (synth.make_tensor((2, 3, 32, 32), 4) / (2 * 3 * 32 * 32)).to(DEV)
    gen.set_precision("fp32")
    with torch.no_grad():
        _, _, _, feats = gen([wplus], input_is_latent=True, randomize_noise=False, return_features=True)
    gen.set_precision("bf16")
    ref_styles = [torch.from_numpy(g[f"style_{i}"]) for i in range(11)]
    edited = [(s * (1 + 0.05 * synth.make_tensor(tuple(s.shape), 500 + i))).to(DEV).requires_grad_(True)
              for i, s in enumerate(ref_styles)]
    mask = synth.make_mask(2, 16, seed=10).to(DEV).requires_grad_(True)
    img, _ = gen([edited], input_is_stylespace=True, randomize_noise=False, attention_layer=7, attention_map=mask,
                 feature_map=feats)
    assert "Synthesis" in type(img.grad_fn).__name__
    (img * upstream).sum().backward()
    gen.assert_ok()
    got = torch.cat([s.grad.flatten() for s in edited])
    want = torch.cat([torch.from_numpy(g[f"grad_style_{i}"]).flatten() for i in range(11)])
    cos, rel = cos_rel(got, want)
    assert cos >= 0.99 and rel <= 0.12, (cos, rel)
    cos_m, rel_m = cos_rel(mask.grad, torch.from_numpy(g["grad_mask"]))
    assert cos_m >= 0.99 and rel_m <= 0.12, (cos_m, rel_m)
    # a blend at a ToRGB layer only (layer 8 = the RGB image at 16^2), and at an up-convolution (layer 6)
    for layer, msize in ((8, 16), (6, 8)):
        st = [s.detach().clone().requires_grad_(True) for s in edited]
        mk = synth.make_mask(2, msize, seed=3 + layer).to(DEV).requires_grad_(True)
        out, _ = gen([st], input_is_stylespace=True, randomize_noise=False, attention_layer=layer, attention_map=mk,
                     feature_map=feats)
        (out * upstream).sum().backward()
        gen.set_precision("fp32")
        st32 = [s.detach().clone().requires_grad_(True) for s in edited]
        mk32 = mk.detach().clone().requires_grad_(True)
        out32, _ = gen([st32], input_is_stylespace=True, randomize_noise=False, attention_layer=layer, attention_map=mk32,
                       feature_map=feats)
        (out32 * upstream).sum().backward()
        gen.set_precision("bf16")
        c1, r1 = cos_rel(torch.cat([s.grad.flatten() for s in st]), torch.cat([s.grad.flatten() for s in st32]))
        c2, r2 = cos_rel(mk.grad, mk32.grad)
        assert c1 >= 0.99 and c2 >= 0.99, (layer, c1, r1, c2, r2)
    gen.assert_ok()


def test_module_path_is_kept_where_the_engine_does_not_apply():
    gen = build(32)
    w = synth.make_wplus(1, 8, seed=2).to(DEV)
    gen.bf16_backward = "modules"
    wp = w.clone().requires_grad_(True)
    img, _ = gen([wp], input_is_latent=True, randomize_noise=False)
    assert "Synthesis" not in type(img.grad_fn).__name__
    gen.bf16_backward = "engine"
    gen.conv1.conv.weight.requires_grad_(True)          # a trainable generator parameter: not the engine's case
    wp = w.clone().requires_grad_(True)
    img, _ = gen([wp], input_is_latent=True, randomize_noise=False)
    assert "Synthesis" not in type(img.grad_fn).__name__
    img.square().mean().backward()
    assert torch.isfinite(wp.grad).all()


def test_backward_kernels_against_torch_formulas():
    lib = N.load()
    torch.manual_seed(0)
    b, h, w, c = 3, 20, 12, 64
    act = torch.randn(b, h, w, c, device=DEV).to(torch.bfloat16)
    gxs = torch.randn(b, h, w, c, device=DEV).to(torch.bfloat16)
    s_next, s_rgb, d = (1 + 0.3 * torch.randn(b, c, device=DEV) for _ in range(3))
    g_rgb = torch.randn(b, 3, h, w, device=DEV)
    w_rgb = torch.randn(3, c, device=DEV)
    noise = torch.randn(1, 1, h, w, device=DEV)
    nw = torch.tensor([0.3], device=DEV)
    bias = 0.1 * torch.randn(c, device=DEV)
    gz = torch.empty_like(act)
    sums = torch.empty(b, 3, c, device=DEV)
    ws = torch.empty(int(lib.w2e_grad_assemble_workspace(b, h * w, c)), device=DEV)
    N.check(lib.w2e_grad_assemble_nhwc(N.ptr(gxs), N.ptr(s_next), N.ptr(act), N.ptr(g_rgb), N.ptr(w_rgb), N.ptr(s_rgb),
                                       N.ptr(noise), N.ptr(nw), 1, N.ptr(bias), N.ptr(d), N.ACT_LRELU, N.ptr(gz), N.ptr(sums),
                                       N.ptr(ws), b, h * w, c, N.stream_ptr()), "grad_assemble")
    a64, g64 = act.double(), gxs.double()
    t = torch.einsum("bohw,oc->bhwc", g_rgb.double(), w_rgb.double())
    ga = g64 * s_next.double()[:, None, None] + t * s_rgb.double()[:, None, None]
    pos = a64 > 0
    gp = ga * torch.where(pos, 2 ** 0.5, 0.2 * 2 ** 0.5)
    y = a64 / torch.where(pos, 2 ** 0.5, 0.2 * 2 ** 0.5)
    r3 = (gp * (y - 0.3 * noise.double()[0, 0, :, :, None] - bias.double())).sum((1, 2))
    want_gz = gp * d.double()[:, None, None]
    scale = float(want_gz.abs().max())
    assert max_abs(gz.double().cpu(), want_gz.cpu()) <= 1e-2 * scale          # bf16 rounding of the result
    for r, want in enumerate(((g64 * a64).sum((1, 2)), (t * a64).sum((1, 2)), r3)):
        assert max_abs(sums[:, r].cpu(), want.cpu()) <= 1e-4 * float(want.abs().max()), r
    # rowdot (with a batch-broadcast second operand) and sum4
    dot = torch.empty(b, c, device=DEV)
    N.check(lib.w2e_rowdot_nhwc(N.ptr(gxs), N.ptr(act[:1].contiguous()), 1, N.ptr(dot), N.ptr(ws), b, h * w, c, N.stream_ptr()),
            "rowdot")
    want = (g64 * a64[:1]).sum((1, 2))
    assert max_abs(dot.cpu(), want.cpu()) <= 1e-4 * float(want.abs().max())
    ys = [torch.randn(b, h + 1 - p, w + 1 - q, c, device=DEV).to(torch.bfloat16) for p in (0, 1) for q in (0, 1)]
    out = torch.empty(b, h, w, c, device=DEV, dtype=torch.bfloat16)
    N.check(lib.w2e_sum4_nhwc(*[N.ptr(y) for y in ys], N.ptr(out), b, h, w, c, N.stream_ptr()), "sum4")
    want = sum(y[:, :h, :w].float() for y in ys).to(torch.bfloat16)
    assert torch.equal(out, want)
    # gradient of the ToRGB skip upsample against autograd of the public upfirdn2d
    skip = torch.randn(2, 3, 10, 14, device=DEV, requires_grad=True)
    kern = w2e.make_kernel([1, 3, 3, 1]).to(DEV) * 4
    up = w2e.upfirdn2d(skip, kern, up=2, down=1, pad=(2, 1))
    gup = torch.randn_like(up)
    up.backward(gup)
    got = torch.empty(2, 3, 10, 14, device=DEV)
    N.check(lib.w2e_skip_grad(N.ptr(gup.contiguous()), N.ptr(got), N.host_floats([0.25, 0.75, 0.75, 0.25]), 6, 10, 14,
                              N.stream_ptr()), "skip_grad")
    assert max_abs(got.cpu(), skip.grad.cpu()) <= 1e-5


@pytest.mark.parametrize("cin,cout,h", [(64, 32, 20), (256, 128, 24), (512, 512, 8), (64, 32, 40), (256, 128, 32),
                                        (128, 64, 36), (512, 256, 32), (512, 512, 33), (64, 32, 129)])
def test_transposed_conv_dgrad_by_parity_classes(cin, cout, h):
    """four strided-view launches with 4 / 2 / 2 / 1 taps + sum4 == the gradient of conv_transpose2d(stride 2)"""
    from where2edit_b200 import train_engine
    gen = w2e.Generator(8, 512, 1, precision="bf16").to(DEV)
    eng = train_engine.TrainEngine(gen)
    b = 2
    weight = synth.make_tensor((1, cout, cin, 3, 3), 91).to(DEV)
    pw = K.PackedWeight(weight, 1 / (cin * 9) ** 0.5, None)
    gz = torch.randn(b, 2 * h + 1, 2 * h + 1, cout, device=DEV).to(torch.bfloat16)
    # 64 -> 32 channels: ONE launch (w2e_modconv_tc2_dgrad_up: four class tiles, one accumulator); otherwise per class
    before = N.STATS.launches.get("w2e_modconv_tc2_dgrad_up", 0)
    before_k = N.STATS.launches.get("w2e_modconv_tc2_dgrad_up_k", 0)
    fused = eng._dgrad_up(gz, pw, h, h)
    eng.assert_ok()
    assert N.STATS.launches.get("w2e_modconv_tc2_dgrad_up", 0) - before == (1 if (cin, cout) == (64, 32) and h > 16 else 0)
    # >= 64 output channels of the forward, >= 16 rows: the K-loop form (w2e_modconv_tc2_dgrad_up_k), also ONE launch
    assert (N.STATS.launches.get("w2e_modconv_tc2_dgrad_up_k", 0) - before_k) == (1 if cout in (64, 128) and h >= 16 else 0)
    eng.dgrad_up_fused = False
    got = eng._dgrad_up(gz, pw, h, h)          # (h >= 32: accumulated in place by TMA reduce-add)
    eng.assert_ok()
    eng.dgrad_up_in_place = False
    parts = eng._dgrad_up(gz, pw, h, h)
    assert max_abs(got.float().cpu(), parts.float().cpu()) <= 2e-2 * float(parts.float().abs().max())
    x = torch.zeros(b, cin, h, h, device=DEV, dtype=torch.float64, requires_grad=True)
    wt = (weight[0].double() / (cin * 9) ** 0.5).to(torch.bfloat16).double()
    y = torch.nn.functional.conv_transpose2d(x, wt.transpose(0, 1), stride=2)
    (y * gz.double().permute(0, 3, 1, 2)).sum().backward()
    want = x.grad.permute(0, 2, 3, 1)
    scale = float(want.abs().max())
    assert max_abs(got.double().cpu(), want.cpu()) <= 1.5e-2 * scale
    # the fused launch rounds once (fp32 accumulation of all nine taps), the per-class launches once per class
    assert max_abs(fused.double().cpu(), want.cpu()) <= 8e-3 * scale


@pytest.mark.parametrize("cin,cout,h", [(512, 256, 32), (512, 512, 33), (256, 128, 17)])
def test_transposed_conv_dgrad_k_loop_form_at_every_width(cin, cout, h):
    """w2e_modconv_tc2_dgrad_up_k directly (the engine only routes 64 / 128-channel layers through it): wide layers and a
    ragged 17-row grid against the gradient of conv_transpose2d(stride 2)"""
    from where2edit_b200 import train_engine
    gen = w2e.Generator(8, 512, 1, precision="bf16").to(DEV)
    eng = train_engine.TrainEngine(gen)
    b = 2
    weight = synth.make_tensor((1, cout, cin, 3, 3), 93).to(DEV)
    pw = K.PackedWeight(weight, 1 / (cin * 9) ** 0.5, None)
    gz = torch.randn(b, 2 * h + 1, 2 * h + 1, cout, device=DEV).to(torch.bfloat16)
    out = torch.empty((b, h, h, cin), device=DEV, dtype=torch.bfloat16)
    masks = sum(train_engine._CLASS_TAPS[(c >> 1, c & 1)] << (9 * c) for c in range(4))
    N.check(N.load().w2e_modconv_tc2_dgrad_up_k(N.ptr(gz), N.ptr(pw.tc_dgrad_up_k()), masks, None, N.ptr(out),
                                                N.ptr(eng.error_flag(gz.device)), b, cout, cin, h, h, None, N.stream_ptr()),
            "modconv_tc2_dgrad_up_k")
    eng.assert_ok()
    x = torch.zeros(b, cin, h, h, device=DEV, dtype=torch.float64, requires_grad=True)
    wt = (weight[0].double() / (cin * 9) ** 0.5).to(torch.bfloat16).double()
    y = torch.nn.functional.conv_transpose2d(x, wt.transpose(0, 1), stride=2)
    (y * gz.double().permute(0, 3, 1, 2)).sum().backward()
    want = x.grad.permute(0, 2, 3, 1)
    assert max_abs(out.double().cpu(), want.cpu()) <= 8e-3 * float(want.abs().max())


@pytest.mark.parametrize("b,h", [(2, 48), (1, 144)])
def test_last_layer_dgrad_on_pixel_pairs_matches_the_pixel_kernel(b, h):
    """w2e_modconv_tc2_pair (the 32 -> 32 dgrad as N = 64 MMAs on pixel pairs) against the ordinary kernel: same bf16
    products, only the fp32 summation order inside the tensor core differs"""
    from where2edit_b200 import train_engine
    gen = w2e.Generator(8, 512, 1, precision="bf16").to(DEV)
    eng = train_engine.TrainEngine(gen)
    weight = synth.make_tensor((1, 32, 32, 3, 3), 95).to(DEV)
    pw = K.PackedWeight(weight, 1 / (32 * 9) ** 0.5, None)
    gz = torch.randn(b, h, h, 32, device=DEV).to(torch.bfloat16)
    before = N.STATS.launches.get("w2e_modconv_tc2_pair", 0)
    got = eng._dgrad_plain(gz, pw)
    eng.assert_ok()
    assert N.STATS.launches.get("w2e_modconv_tc2_pair", 0) - before == 1
    eng.pair_mode = False
    want = eng._dgrad_plain(gz, pw)
    eng.assert_ok()
    assert N.STATS.launches.get("w2e_modconv_tc2_pair", 0) - before == 1
    scale = float(want.float().abs().max())
    assert max_abs(got.float().cpu(), want.float().cpu()) <= 8e-3 * scale      # one bf16 rounding step of the stored value
    assert float((got.float() - want.float()).abs().mean()) <= 2e-4 * scale
