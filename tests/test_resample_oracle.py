"""oracle/resample_oracle.py against vectors from the reference's own CLIPLoss.upsample / .avg_pool modules."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import resample_oracle as rs  # noqa: E402

CASES = ["s64", "s32", "s160", "ragged"]


@pytest.fixture(scope="module")
def golden_resample():
    return np.load(os.path.join(ROOT, "tests", "golden", "resample.npz"))


@pytest.mark.parametrize("name", CASES)
def test_resample_oracle_matches_reference(golden_resample, name):
    g = golden_resample
    scale, pool = int(g[f"{name}/scale"]), int(g[f"{name}/pool"])
    y = rs.clip_resample(g[f"{name}/x"], scale, pool)
    assert y.shape == g[f"{name}/y"].shape
    np.testing.assert_allclose(y, g[f"{name}/y"], rtol=0, atol=2e-6)
    gx = rs.clip_resample_backward(g[f"{name}/gy"], g[f"{name}/x"].shape[-2:], scale, pool)
    np.testing.assert_allclose(gx, g[f"{name}/gx"], rtol=0, atol=2e-6 * np.abs(g[f"{name}/gx"]).max())


def test_nearest_index_is_integer_division():
    """floor(dst * fp32(1/7)) == dst // 7 over the whole 7168-wide upsampled axis of a 1024^2 image."""
    x = np.arange(1024, dtype=np.float32)[None, :].repeat(2, 0)
    up = rs.upsample_nearest(x, 7)
    assert np.array_equal(up[0], (np.arange(7168) // 7).astype(np.float32))
