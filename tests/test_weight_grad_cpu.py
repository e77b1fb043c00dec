"""The opt-in convolution-weight gradient of ModulatedConv2d (functional.modconv_weight_grad + the identity tap
that hands it to autograd) is plain device-agnostic torch, so its math is checked here on CPU against autograd
through the reference formulation (oracle.modulated_conv2d_ref, models/stylegan2/model.py:234-276)."""
import math
import os
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import stylegan2_oracle as orc  # noqa: E402
from where2edit_b200 import functional as K  # noqa: E402


def reference_pre_blur(x, s, weight, demodulate, upsample):
    """Reference formulation up to (not including) the Blur of the up path: model.py:240-257 / :262-274."""
    b, cin, h, w = x.shape
    _, cout, _, k, _ = weight.shape
    if not upsample:
        return orc.modulated_conv2d_ref(x, s.reshape(b, 1, cin, 1, 1), weight, None, None, demodulate=demodulate,
                                        input_is_stylespace=True)[0]
    wmod = (1 / math.sqrt(cin * k * k)) * weight * s.reshape(b, 1, cin, 1, 1)
    if demodulate:
        wmod = wmod * torch.rsqrt(wmod.pow(2).sum([2, 3, 4]) + 1e-8).reshape(b, cout, 1, 1, 1)
    wt = wmod.transpose(1, 2).reshape(b * cin, cout, k, k)
    out = F.conv_transpose2d(x.reshape(1, b * cin, h, w), wt, padding=0, stride=2, groups=b)
    return out.reshape(b, cout, out.shape[-2], out.shape[-1])


@pytest.mark.parametrize("k,upsample,demodulate", [(1, False, True), (3, False, True), (3, True, True),
                                                   (1, False, False), (3, True, False)])
def test_weight_gradient_matches_reference_autograd(k, upsample, demodulate):
    torch.manual_seed(10 * k + upsample)
    b, cin, cout, h = 3, 5, 4, 6
    x = torch.randn(b, cin, h, h, dtype=torch.float64)
    s = 1 + 0.3 * torch.randn(b, cin, dtype=torch.float64)
    weight = torch.randn(1, cout, cin, k, k, dtype=torch.float64, requires_grad=True)
    y_ref = reference_pre_blur(x, s, weight, demodulate, upsample)
    head = torch.randn_like(y_ref)
    (gw_ref,) = torch.autograd.grad((y_ref * head).sum(), weight)

    scale = 1 / math.sqrt(cin * k * k)
    d = None
    if demodulate:
        d = torch.rsqrt(((scale * weight.detach()[0]) ** 2).sum((2, 3))[None] .mul(s.square()[:, None, :]).sum(-1) + 1e-8)
    y = y_ref.detach()
    gw = K.modconv_weight_grad(x, s, d, y, head, weight.detach(), scale, k, upsample)
    assert gw.shape == weight.shape
    assert float((gw - gw_ref).abs().max()) <= 1e-10 * max(1.0, float(gw_ref.abs().max()))

    # the tap: identity on y, weight gradient delivered through autograd, upstream gradient passed through
    w_param = weight.detach().clone().requires_grad_(True)
    y_leaf = y.clone().requires_grad_(True)
    out = K.weight_grad_tap(y_leaf, w_param, x, s, d, scale, k, upsample)
    assert torch.equal(out, y_leaf)
    (out * head).sum().backward()
    assert float((w_param.grad - gw_ref).abs().max()) <= 1e-10 * max(1.0, float(gw_ref.abs().max()))
    assert torch.equal(y_leaf.grad, head)


def test_enable_weight_gradients_marks_every_modulated_conv():
    import where2edit_b200 as w2e
    m = torch.nn.Sequential(w2e.StyledConv(8, 8, 1, 16), w2e.StyledConv(8, 4, 3, 16, upsample=True))
    assert not any(c.conv.weight_grad for c in m)
    w2e.enable_weight_gradients(m)
    assert all(c.conv.weight_grad for c in m)
    w2e.enable_weight_gradients(m, False)
    assert not any(c.conv.weight_grad for c in m)


def test_trained_weights_are_repacked_after_a_data_write():
    """Optimisers that write through `.data` (the reference's Ranger, mapper/training/ranger.py:155-162) do not bump the
    tensor's version counter; a module whose weight is being trained must not serve a cached layout of the old weight."""
    import where2edit_b200 as w2e
    m = w2e.ModulatedConv2d(8, 4, 1, 8)
    m.weight_grad = True
    before = m.packed().rgb.clone()
    v = m.weight._version
    m.weight.data.copy_(m.weight.data * 2)
    assert m.weight._version == v                        # the hazard: nothing tells the cache that the weight changed
    assert torch.allclose(m.packed().rgb, 2 * before)
    m.weight_grad = False                                # frozen modules keep the cache (keyed on pointer / version / device)
    assert m.packed() is m.packed()
