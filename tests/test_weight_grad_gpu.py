"""Opt-in training of StyledConv parameters (where2edit_b200.enable_weight_gradients): gradients of the
convolution weight, modulation layer, noise weight and activation bias against autograd through the CPU oracle
(the reference formulation) -- the cluster-style mapper's attention heads (attention/run_attention.py:725-735)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import where2edit_b200 as w2e  # noqa: E402
from oracle import stylegan2_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("cin,cout,k,upsample,stylespace", [(16, 8, 1, False, True), (12, 6, 3, False, False),
                                                            (8, 8, 3, True, False)])
def test_trainable_styled_conv_parameter_gradients(cin, cout, k, upsample, stylespace):
    torch.manual_seed(3 + k)
    b, h, style_dim = 2, 8, 16
    m = w2e.StyledConv(cin, cout, k, style_dim, upsample=upsample)
    with torch.no_grad():
        m.noise.weight.fill_(0.3)
        m.activate.bias.copy_(0.2 * torch.randn(cout))
    m = w2e.enable_weight_gradients(m).to(DEV).train()
    x = torch.randn(b, cin, h, h)
    style = (1 + 0.3 * torch.randn(b, 1, cin, 1, 1)) if stylespace else torch.randn(b, style_dim)
    oh = 2 * h if upsample else h
    noise = torch.randn(b, 1, oh, oh)
    head = torch.randn(b, cout, oh, oh)

    # oracle (CPU, autograd through the reference formulation)
    sd = {f"h.{n}": p.detach().cpu().clone().requires_grad_(True) for n, p in m.named_parameters()}
    for n, buf in m.named_buffers():
        sd[f"h.{n}"] = buf.detach().cpu().clone()
    out_o, _ = orc._styled_conv(sd, "h", x, style, noise, upsample, stylespace)
    (out_o * head).sum().backward()

    out, _ = m(x.to(DEV), style.to(DEV), noise=noise.to(DEV), input_is_stylespace=stylespace)
    c = float(out_o.detach().abs().max())
    assert float((out.detach().cpu() - out_o.detach()).abs().max()) <= 1e-4 * max(c, 1.0)
    (out * head.to(DEV)).sum().backward()
    checked = 0
    for n, p in m.named_parameters():
        want = sd[f"h.{n}"].grad
        if want is None:             # modulation parameters are unused with stylespace inputs
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        assert p.grad is not None, f"{n} received no gradient"
        scale = max(float(want.abs().max()), 1e-6)
        assert float((p.grad.cpu() - want).abs().max()) <= 2e-4 * scale, n
        checked += 1
    assert checked >= 3              # conv.weight, noise.weight, activate.bias (+ modulation.*)
