"""The numpy restatement of the region-mask construction (oracle/region_oracle.py) against vectors produced by
the unmodified reference forward (tests/golden/region.npz, oracle/make_region_golden.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import region_oracle as ro  # noqa: E402

CASES = ["same_res", "upsampled"]


@pytest.fixture(scope="module")
def golden_region():
    return np.load(os.path.join(ROOT, "tests", "golden", "region.npz"))


def case(golden, name):
    return {k.split("/", 1)[1]: golden[k] for k in golden.files if k.startswith(name + "/")}


def test_pairwise_distance_matches_reference(golden_region):
    d = ro.pairwise_distance(golden_region["pd/a"], golden_region["pd/b"])
    np.testing.assert_allclose(d, golden_region["pd/dis"], rtol=2e-6, atol=0)


@pytest.mark.parametrize("name", CASES)
def test_cluster_assignment_matches_reference(golden_region, name):
    c = case(golden_region, name)
    ids, dis = ro.assign_clusters(c["feature"], c["centres"], int(c["size"]), int(c["clusters"]))
    np.testing.assert_allclose(dis, c["dis"], rtol=1e-5)
    b, h = c["ids_lowres"].shape[:2]
    low = c["ids_lowres"] + np.arange(b)[:, None, None] * int(c["clusters"])
    idx = ro.nearest_index(int(c["size"]), h)
    assert np.array_equal(ids, low[:, idx][:, :, idx])          # integer ids: bit-exact
    assert len(np.unique(ids)) > int(c["clusters"])              # several clusters populated in each sample


@pytest.mark.parametrize("name", CASES)
def test_region_attention_matches_reference(golden_region, name):
    c = case(golden_region, name)
    k = int(c["clusters"])
    ids, _ = ro.assign_clusters(c["feature"], c["centres"], int(c["size"]), k)
    final, same, loss_reg, loss_tv = ro.region_attention(c["each"], ids, k)
    np.testing.assert_allclose(final, c["final"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(loss_reg, c["loss_reg"], rtol=2e-6)
    np.testing.assert_allclose(loss_tv, c["loss_tv"], rtol=1e-5)
    thr = np.where(same < 0.8, 0, same)
    assert (thr == 0).any() and (thr > 0).any()                  # both sides of the threshold exercised
    # backward: gradient at the logits of attention_last for the seeded head (0.7 * loss_reg + 1.3 * loss_tv)
    g_each = ro.region_attention_backward(c["head"], 0.7, 1.3, c["each"], ids, k)
    g_logits = g_each * c["each"] * (1 - c["each"])
    scale = np.abs(c["g_logits"]).max()
    np.testing.assert_allclose(g_logits / scale, c["g_logits"] / scale, rtol=0, atol=2e-5)


def test_gaussian_kernel_is_normalised_and_symmetric():
    k2 = ro.gaussian_kernel2d()
    assert abs(float(k2.sum()) - 1) < 1e-6 and np.array_equal(k2, k2.T) and np.array_equal(k2, k2[::-1, ::-1])
