"""oracle/mapper_oracle.py against the unmodified reference LevelsMapper (tests/golden/mapper.npz)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import mapper_oracle as mo  # noqa: E402


@pytest.mark.parametrize("name,no_fine", [("all", False), ("no_fine", True)])
def test_levels_mapper_oracle_matches_reference(name, no_fine):
    g = np.load(os.path.join(ROOT, "tests", "golden", "mapper.npz"))
    levels = [lv for lv, off in zip(mo.LEVELS, (False, False, no_fine)) if not off]
    y = mo.levels_mapper(g["x"], mo.mapper_state(levels=levels), no_fine=no_fine)
    scale = np.abs(g[f"{name}/y"]).max()
    np.testing.assert_allclose(y / scale, g[f"{name}/y"] / scale, rtol=0, atol=2e-6)
    if no_fine:
        assert not y[:, 8:].any() and y[:, :8].any()


def test_mapper_state_is_reproducible():
    a, b = mo.mapper_state(), mo.mapper_state()
    assert sorted(a) == sorted(b) and len(a) == 24 and all(np.array_equal(a[k], b[k]) for k in a)
    assert a["course_mapping.mapping.1.weight"].shape == (512, 512)
