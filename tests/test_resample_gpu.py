"""GPU parity of clip_resample (where2edit_b200/resample.py -> libw2e.so) against the reference goldens, the
numpy oracle, and - at the full 1024^2 size - the materialising formulation evaluated by torch on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import resample_oracle as rs  # noqa: E402
from where2edit_b200 import resample  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def golden_resample():
    return np.load(os.path.join(ROOT, "tests", "golden", "resample.npz"))


@pytest.mark.parametrize("name", ["s64", "s32", "s160", "ragged"])
def test_clip_resample_matches_reference_golden(golden_resample, name):
    g = golden_resample
    scale, pool = int(g[f"{name}/scale"]), int(g[f"{name}/pool"])
    x = torch.from_numpy(g[f"{name}/x"]).to(DEV).requires_grad_(True)
    y = resample.clip_resample(x, scale, pool)
    assert tuple(y.shape) == g[f"{name}/y"].shape
    np.testing.assert_allclose(y.detach().cpu().numpy(), g[f"{name}/y"], rtol=0, atol=2e-6)
    (y * torch.from_numpy(g[f"{name}/gy"]).to(DEV)).sum().backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), g[f"{name}/gx"], rtol=0, atol=2e-6 * np.abs(g[f"{name}/gx"]).max())


@pytest.mark.parametrize("shape,scale,pool", [((1, 1, 5, 7), 7, 3), ((2, 3, 16, 16), 3, 8), ((1, 2, 9, 9), 1, 2),
                                              ((1, 1, 4, 4), 9, 2), ((0, 3, 8, 8), 7, 2), ((1, 1, 3, 50), 2, 6), ((1, 2, 12, 20), 7, 9),
                                              ((1, 3, 256, 256), 7, 8)])
def test_clip_resample_matches_oracle(shape, scale, pool):
    """Windows wider and narrower than a source pixel, scale 1 (pure pooling), ragged remainders, empty batch."""
    rng = np.random.default_rng(sum(shape) + scale)
    x = rng.standard_normal(shape).astype(np.float32)
    xt = torch.from_numpy(x).to(DEV).requires_grad_(True)
    y = resample.clip_resample(xt, scale, pool)
    y_o = rs.clip_resample(x, scale, pool)
    assert tuple(y.shape) == y_o.shape
    np.testing.assert_allclose(y.detach().cpu().numpy(), y_o, rtol=0, atol=2e-6)
    gy = rng.standard_normal(y_o.shape).astype(np.float32)
    (y * torch.from_numpy(gy).to(DEV)).sum().backward()
    g_o = rs.clip_resample_backward(gy, shape[-2:], scale, pool)
    np.testing.assert_allclose(xt.grad.cpu().numpy(), g_o, rtol=0, atol=2e-6 * max(1.0, float(np.abs(g_o).max()) if g_o.size else 1.0))


def test_clip_resample_rejects_bad_arguments():
    x = torch.randn(1, 3, 8, 8, device=DEV)
    with pytest.raises(ValueError):
        resample.clip_resample(x, 2.5, 2)
    with pytest.raises(ValueError):
        resample.clip_resample(x, 1, 9)
    with pytest.raises(ValueError):
        resample.clip_resample(x[0], 7, 2)
    with pytest.raises(RuntimeError):
        resample.clip_resample(x.cpu(), 7, 2)


def test_clip_resample_at_1024():
    """1024^2 images, 7x up, 32x32 pool -> 224^2 (the reference's configuration): against the materialised
    [1,3,7168,7168] formulation on one image, and constant / linearity / adjoint properties on the batch."""
    torch.manual_seed(8)
    mod = resample.ClipResample(1024)
    x = torch.randn(4, 3, 1024, 1024, device=DEV)
    y = mod(x)
    assert y.shape == (4, 3, 224, 224)
    # fp64 evaluation of the materialised formulation: the one-pass result is within fp32 rounding of it ...
    ref64 = torch.nn.functional.avg_pool2d(torch.nn.functional.interpolate(x[1:2].double(), scale_factor=7), 32)
    assert float((y[1:2].double() - ref64).abs().max()) < 5e-7
    # ... and the fp32 materialised formulation (1024-term fp32 sums per pixel) within its own summation error
    ref = torch.nn.functional.avg_pool2d(torch.nn.functional.interpolate(x[1:2], scale_factor=7), 32)
    assert float((y[1:2] - ref).abs().max()) < 3e-5
    assert float((ref.double() - ref64).abs().max()) > float((y[1:2].double() - ref64).abs().max())
    del ref, ref64
    const = mod(torch.full((1, 3, 1024, 1024), 0.37, device=DEV))
    assert float((const - 0.37).abs().max()) < 1e-6
    x2 = torch.randn_like(x)
    assert float((mod(2 * x - 3 * x2) - (2 * y - 3 * mod(x2))).abs().max()) < 1e-5
    xg = x.clone().requires_grad_(True)
    gy = torch.randn_like(y)
    (mod(xg) * gy).sum().backward()
    lhs = float((y.double() * gy.double()).sum())
    rhs = float((x.double() * xg.grad.double()).sum())
    assert abs(lhs - rhs) < 1e-6 * max(1.0, abs(lhs)) + 1e-3
    assert abs(float(xg.grad.double().sum()) - float(gy.double().sum())) < 1e-2   # weights of a window sum to 1
