"""GPU parity of the public ops (through the C ABI) against the oracle and the reference goldens."""
import numpy as np
import pytest
import torch

import where2edit_b200 as w2e
from where2edit_b200 import functional as K
from conftest import max_abs
from oracle import make_golden as mg
from oracle import stylegan2_oracle as orc
from oracle import synth
from test_oracle_golden import MODCONV_CASES, modconv_case_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_upfirdn2d_golden_cases(golden_ops):
    for i, (shape, kspec, *geom) in enumerate(mg.UPFIRDN_CASES):
        x = synth.make_tensor(shape, 100 + i)
        k = mg.make_kernel_spec(kspec)
        got = w2e.upfirdn2d_native(x.to(DEV), k.to(DEV), *geom).cpu()
        assert tuple(got.shape) == golden_ops[f"upfirdn_{i}"].shape
        assert max_abs(got, golden_ops[f"upfirdn_{i}"]) <= 2e-6, (i, geom)


@pytest.mark.parametrize("shape,up,down,pad", [
    ((2, 5, 33, 33), 1, 1, (1, 1)), ((1, 3, 129, 129), 1, 1, (1, 1)), ((3, 2, 65, 63), 1, 1, (2, 2)),
    ((2, 3, 16, 16), 2, 1, (2, 1)), ((1, 4, 64, 64), 1, 2, (1, 1)), ((1, 2, 5, 5), 1, 1, (1, 1)),
    ((1, 1, 257, 300), 1, 1, (1, 1)), ((1, 2, 70, 70), 1, 1, (0, 3)),
])
def test_upfirdn2d_blur_kernel_sweep(shape, up, down, pad):
    """ragged / non-tile-multiple sizes on the separable fast path and the generic path"""
    x = synth.make_tensor(shape, 7)
    k = synth.blur_kernel_2d(gain=float(up * up))
    ref = orc.upfirdn2d_ref(x.double(), k.double(), up=up, down=down, pad=pad)
    got = w2e.upfirdn2d(x.to(DEV), k.to(DEV), up=up, down=down, pad=pad).cpu()
    assert got.shape == ref.shape
    assert max_abs(got, ref) <= 2e-6


def test_upfirdn2d_bf16_and_gradients():
    x = synth.make_tensor((2, 3, 17, 17), 8)
    k = synth.blur_kernel_2d(gain=4.0)
    ref = orc.upfirdn2d_ref(x.bfloat16().double(), k.double(), pad=(1, 1))
    got = w2e.upfirdn2d(x.to(DEV).bfloat16(), k.to(DEV), pad=(1, 1))
    assert got.dtype == torch.bfloat16
    assert max_abs(got.float().cpu(), ref) <= 2e-2
    for up, down, pad in [(1, 1, (1, 1)), (2, 1, (2, 1)), (1, 2, (1, 1))]:
        xc = synth.make_tensor((2, 2, 12, 12), 9).double().requires_grad_(True)
        yc = orc.upfirdn2d_ref(xc, k.double(), up=up, down=down, pad=pad)
        gy = synth.make_tensor(tuple(yc.shape), 10)
        yc.backward(gy.double())
        xg = xc.detach().float().to(DEV).requires_grad_(True)
        yg = w2e.upfirdn2d(xg, k.to(DEV), up=up, down=down, pad=pad)
        yg.backward(gy.to(DEV))
        assert max_abs(xg.grad.cpu(), xc.grad) <= 2e-6


def test_fused_leaky_relu_golden_and_layouts(golden_ops):
    for i, (shape, cdim) in enumerate([((5, 7), 7), ((3, 4, 6), 6), ((2, 5, 4, 3), 5)]):
        x = synth.make_tensor(shape, 200 + i)
        b = synth.make_tensor((cdim,), 210 + i)
        got = w2e.fused_leaky_relu(x.to(DEV), b.to(DEV)).cpu()
        assert max_abs(got, golden_ops[f"flrelu_{i}"]) <= 1e-6
        got = w2e.fused_leaky_relu(x.to(DEV), b.to(DEV), 0.1, 1.5).cpu()
        assert max_abs(got, golden_ops[f"flrelu_{i}_args"]) <= 1e-6
    # vectorised path (inner % 4 == 0), large, and module form
    x = synth.make_tensor((3, 16, 32, 32), 1)
    m = w2e.FusedLeakyReLU(16).to(DEV)
    with torch.no_grad():
        m.bias.copy_(synth.make_tensor((16,), 2))
    assert max_abs(m(x.to(DEV)).cpu(), orc.fused_leaky_relu_ref(x, m.bias.cpu().detach())) <= 1e-6
    x2 = synth.make_tensor((4, 5, 512), 3)
    b2 = synth.make_tensor((512,), 4)
    assert max_abs(w2e.fused_leaky_relu(x2.to(DEV), b2.to(DEV)).cpu(), orc.fused_leaky_relu_ref(x2, b2)) <= 1e-6


def test_fused_leaky_relu_gradients():
    for shape, cdim in [((6, 10), 10), ((3, 4, 8), 8), ((2, 6, 5, 5), 6)]:
        x = synth.make_tensor(shape, 5).double().requires_grad_(True)
        b = synth.make_tensor((cdim,), 6).double().requires_grad_(True)
        y = orc.fused_leaky_relu_ref(x, b)
        gy = synth.make_tensor(shape, 7)
        y.backward(gy.double())
        xg = x.detach().float().to(DEV).requires_grad_(True)
        bg = b.detach().float().to(DEV).requires_grad_(True)
        w2e.fused_leaky_relu(xg, bg).backward(gy.to(DEV))
        assert max_abs(xg.grad.cpu(), x.grad) <= 1e-6
        assert max_abs(bg.grad.cpu(), b.grad) <= 1e-5


def test_modulated_conv_golden_cases(golden_ops):
    for i, (cin, cout, k, up, demod, h) in enumerate(MODCONV_CASES):
        weight, mod_w, mod_b, x, w = modconv_case_inputs(i)
        m = w2e.ModulatedConv2d(cin, cout, k, 12, demodulate=demod, upsample=up)
        with torch.no_grad():
            m.weight.copy_(weight)
            m.modulation.weight.copy_(mod_w)
            m.modulation.bias.copy_(mod_b)
        m = m.to(DEV)
        with torch.no_grad():
            y, s = m(x.to(DEV), w.to(DEV))
            assert tuple(s.shape) == (2, 1, cin, 1, 1)
            assert max_abs(y.cpu(), golden_ops[f"modconv_{i}_y"]) <= 2e-5, i
            assert max_abs(s.cpu(), golden_ops[f"modconv_{i}_s"]) <= 1e-5
            y2, s2 = m(x.to(DEV), torch.from_numpy(golden_ops[f"modconv_{i}_s"]).to(DEV) * 1.1,
                       input_is_stylespace=True)
            assert max_abs(y2.cpu(), golden_ops[f"modconv_{i}_y_ss"]) <= 2e-5, i


@pytest.mark.parametrize("cin,cout,h,up", [(70, 67, 19, False), (64, 130, 9, True), (24, 8, 33, False), (40, 33, 6, True)])
def test_modulated_conv_ragged_shapes_and_gradients(cin, cout, h, up):
    """non-tile-multiple channel counts / sizes; gradients w.r.t. input and style vs oracle autograd"""
    weight = synth.make_tensor((1, cout, cin, 3, 3), 21)
    x = synth.make_tensor((2, cin, h, h), 22)
    s = 1 + 0.3 * synth.make_tensor((2, 1, cin, 1, 1), 23)
    blur = synth.blur_kernel_2d(gain=4.0)
    xd, sd_ = x.double().requires_grad_(True), s.double().requires_grad_(True)
    ref, _ = orc.modulated_conv2d_ref(xd, sd_, weight.double(), None, None, True, up, blur.double(),
                                      input_is_stylespace=True)
    gy = synth.make_tensor(tuple(ref.shape), 24)
    ref.backward(gy.double())
    m = w2e.ModulatedConv2d(cin, cout, 3, 16, upsample=up)
    with torch.no_grad():
        m.weight.copy_(weight)
    m = m.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    sg = s.to(DEV).requires_grad_(True)
    y, _ = m(xg, sg, input_is_stylespace=True)
    assert max_abs(y.detach().cpu(), ref.detach()) <= 2e-5
    y.backward(gy.to(DEV))
    gscale = max(1.0, xd.grad.abs().max().item())
    assert max_abs(xg.grad.cpu(), xd.grad) <= 2e-5 * gscale
    sscale = max(1.0, sd_.grad.abs().max().item())
    assert max_abs(sg.grad.cpu(), sd_.grad) <= 5e-5 * sscale


def test_to_rgb_and_gradients():
    cin, h = 24, 16
    m = w2e.ToRGB(cin, 16)
    with torch.no_grad():
        m.conv.weight.copy_(synth.make_tensor((1, 3, cin, 1, 1), 31))
        m.bias.copy_(synth.make_tensor((1, 3, 1, 1), 32, 0.1))
    x = synth.make_tensor((2, cin, h, h), 33)
    s = 1 + 0.3 * synth.make_tensor((2, 1, cin, 1, 1), 34)
    skip = synth.make_tensor((2, 3, h // 2, h // 2), 35)
    sd = {"t.conv.weight": m.conv.weight.detach().double(), "t.conv.modulation.weight": None,
          "t.conv.modulation.bias": None, "t.bias": m.bias.detach().double(),
          "t.upsample.kernel": m.upsample.kernel.double()}
    xd, sd_, kd = x.double().requires_grad_(True), s.double().requires_grad_(True), skip.double().requires_grad_(True)
    ref, _ = orc._to_rgb(sd, "t", xd, sd_, kd, True)
    gy = synth.make_tensor(tuple(ref.shape), 36)
    ref.backward(gy.double())
    m = m.to(DEV)
    xg, sg, kg = (t.to(DEV).requires_grad_(True) for t in (x, s, skip))
    out, _ = m(xg, sg, kg, input_is_stylespace=True)
    assert max_abs(out.detach().cpu(), ref.detach()) <= 1e-5
    out.backward(gy.to(DEV))
    assert max_abs(xg.grad.cpu(), xd.grad) <= 1e-5
    assert max_abs(sg.grad.cpu(), sd_.grad) <= 1e-4
    assert max_abs(kg.grad.cpu(), kd.grad) <= 1e-5
    with torch.no_grad():
        out0, _ = m(xg, sg, None, input_is_stylespace=True)
    ref0, _ = orc._to_rgb(sd, "t", x.double(), s.double(), None, True)
    assert max_abs(out0.cpu(), ref0) <= 1e-5


@pytest.mark.parametrize("h,mh", [(16, 16), (16, 8), (32, 12), (8, 20), (64, 64)])
def test_mask_blend_is_bit_exact_and_differentiable(h, mh):
    e = synth.make_tensor((2, 5, h, h), 41)
    o = synth.make_tensor((2, 5, h, h), 42)
    m = synth.make_mask(2, mh, seed=43)
    ref = orc.mask_blend_ref(e, o, m)                       # fp32, same op order as the reference
    got = K.mask_blend(e.to(DEV), o.to(DEV), m.to(DEV)).cpu()
    assert torch.equal(got, ref), "mask indexing / blend must be bit-exact in fp32"
    import torch.nn.functional as F
    assert torch.equal(orc.nearest_resize_ref(m, h), F.interpolate(m, h))
    ed, md = e.double().requires_grad_(True), m.double().requires_grad_(True)
    out = orc.mask_blend_ref(ed, o.double(), md)
    g = synth.make_tensor((2, 5, h, h), 44)
    out.backward(g.double())
    eg, mg_ = e.to(DEV).requires_grad_(True), m.to(DEV).requires_grad_(True)
    K.mask_blend(eg, o.to(DEV), mg_).backward(g.to(DEV))
    assert max_abs(eg.grad.cpu(), ed.grad) <= 1e-6
    assert max_abs(mg_.grad.cpu(), md.grad) <= 1e-4
