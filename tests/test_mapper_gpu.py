"""GPU parity of where2edit_b200.mappers.LevelsMapper against the reference golden and the numpy oracle."""
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import mapper_oracle as mo  # noqa: E402
from where2edit_b200 import mappers  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build(no_fine):
    opts = types.SimpleNamespace(no_coarse_mapper=False, no_medium_mapper=False, no_fine_mapper=no_fine)
    m = mappers.LevelsMapper(opts)
    levels = [lv for lv, off in zip(mo.LEVELS, (False, False, no_fine)) if not off]
    m.load_state_dict({k: torch.from_numpy(v) for k, v in mo.mapper_state(levels=levels).items()}, strict=True)
    return m.to(DEV)


@pytest.mark.parametrize("name,no_fine", [("all", False), ("no_fine", True)])
def test_levels_mapper_matches_reference_golden(name, no_fine):
    g = np.load(os.path.join(ROOT, "tests", "golden", "mapper.npz"))
    m = build(no_fine)
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    y = m(x)
    scale = np.abs(g[f"{name}/y"]).max()
    np.testing.assert_allclose(y.detach().cpu().numpy() / scale, g[f"{name}/y"] / scale, rtol=0, atol=1e-5)
    (y * torch.from_numpy(g[f"{name}/head"]).to(DEV)).sum().backward()
    gs = np.abs(g[f"{name}/gx"]).max()
    np.testing.assert_allclose(x.grad.cpu().numpy() / gs, g[f"{name}/gx"] / gs, rtol=0, atol=2e-5)
    if name == "all":   # parameter gradients (the mapper is trained by the edit loop): subsample + sums
        for pn, p in m.named_parameters():
            want = g[f"{name}/grad/{pn}"]
            got = p.grad.reshape(-1)[::997].cpu().numpy()
            sc = max(float(np.abs(want).max()), 1e-12)
            np.testing.assert_allclose(got / sc, want / sc, rtol=0, atol=5e-5, err_msg=pn)
            sums = g[f"{name}/gradsum/{pn}"]
            assert abs(float(p.grad.double().abs().sum()) - sums[1]) <= 1e-4 * sums[1], pn


def test_levels_mapper_matches_oracle_on_a_ragged_batch():
    """Batch 3, 14 latent rows (a 256^2 generator's W+): the fine level gets 6 rows."""
    rng = np.random.default_rng(9)
    x = rng.standard_normal((3, 14, 512)).astype(np.float32)
    y_o = mo.levels_mapper(x, mo.mapper_state())
    with torch.no_grad():
        y = build(False)(torch.from_numpy(x).to(DEV)).cpu().numpy()
    scale = np.abs(y_o).max()
    np.testing.assert_allclose(y / scale, y_o / scale, rtol=0, atol=1e-5)
