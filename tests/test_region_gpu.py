"""GPU parity of the region-mask construction (where2edit_b200/region.py -> libw2e.so) against the reference
goldens (tests/golden/region.npz) and the numpy oracle (oracle/region_oracle.py), plus size-independent
properties at the BASELINE cfg3 shape (B=64, 64x64 attention, 512-channel feature, 20 clusters)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import region_oracle as ro  # noqa: E402
from where2edit_b200 import region  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def golden_region():
    return np.load(os.path.join(ROOT, "tests", "golden", "region.npz"))


def case(golden, name):
    return {k.split("/", 1)[1]: golden[k] for k in golden.files if k.startswith(name + "/")}


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("name", ["same_res", "upsampled"])
def test_region_mask_matches_reference_golden(golden_region, name):
    c = case(golden_region, name)
    k, size = int(c["clusters"]), int(c["size"])
    ids = region.assign_clusters(t(c["feature"]), t(c["centres"]), size, k)
    b, h = c["ids_lowres"].shape[:2]
    low = c["ids_lowres"] + np.arange(b)[:, None, None] * k
    idx = ro.nearest_index(size, h)
    assert ids.dtype == torch.int64 and np.array_equal(ids.cpu().numpy(), low[:, idx][:, :, idx])   # bit-exact ids
    logits = t(c["logits"]).requires_grad_(True)
    each = torch.sigmoid(logits + float(c["bias"]))
    final, same, loss_reg, loss_tv = region.region_attention(each, ids, k)
    assert final.shape == (b, 1, size, size) and same.shape == (b, size, size) and loss_reg.shape == (1,) and loss_tv.shape == ()
    np.testing.assert_allclose(final.detach().cpu().numpy(), c["final"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(loss_reg.detach().cpu().numpy(), c["loss_reg"], rtol=2e-6)
    np.testing.assert_allclose(loss_tv.item(), c["loss_tv"], rtol=1e-5)
    assert not same.requires_grad
    ((final * t(c["head"])).sum() + 0.7 * loss_reg.sum() + 1.3 * loss_tv).backward()
    scale = np.abs(c["g_logits"]).max()
    np.testing.assert_allclose(logits.grad.cpu().numpy() / scale, c["g_logits"] / scale, rtol=0, atol=2e-5)


@pytest.mark.parametrize("b,s,k,used", [(1, 3, 1, 1), (2, 5, 4, 4), (3, 37, 7, 5), (2, 64, 40, 40), (1, 130, 9, 9)])
def test_region_attention_matches_oracle(b, s, k, used):
    """Minimum size for the reflect padding, ragged sizes, empty clusters (used < k), more clusters than warps,
    pixels claimed by no cluster (id >= B*K keeps the reference's initial 1.0)."""
    rng = np.random.default_rng(100 * s + k)
    each = rng.uniform(0.55, 0.99, (b, s, s)).astype(np.float32)
    ids = rng.integers(0, used, (b, s, s)) + np.arange(b)[:, None, None] * k
    if s >= 5:
        ids[:, 0, :3] = b * k + 5       # unclaimed pixels
    final_o, same_o, reg_o, tv_o = ro.region_attention(each, ids, k)
    e = t(each).requires_grad_(True)
    final, same, reg, tv = region.region_attention(e, t(ids.astype(np.int64)), k, validate=True)
    np.testing.assert_allclose(final.detach().cpu().numpy(), final_o, rtol=0, atol=2e-6)
    np.testing.assert_allclose(same.cpu().numpy(), same_o, rtol=0, atol=1e-6)
    np.testing.assert_allclose(reg.detach().cpu().numpy(), reg_o, rtol=3e-6, atol=1e-7)
    np.testing.assert_allclose(tv.item(), tv_o, rtol=1e-5, atol=1e-9)
    head = rng.standard_normal(final_o.shape).astype(np.float32)
    ((final * t(head)).sum() + 0.3 * reg.sum() - 2.0 * tv).backward()
    g_o = ro.region_attention_backward(head, 0.3, -2.0, each, ids, k)
    scale = np.abs(g_o).max()
    np.testing.assert_allclose(e.grad.cpu().numpy() / scale, g_o / scale, rtol=0, atol=2e-5)


@pytest.mark.parametrize("b,c,h,k,size", [(2, 32, 2, 3, 4), (3, 64, 9, 11, 9), (2, 48, 12, 20, 30), (1, 512, 16, 17, 16)])
def test_cluster_assignment_matches_oracle(b, c, h, k, size):
    rng = np.random.default_rng(7 * c + h)
    feat = rng.standard_normal((b, c, h, h)).astype(np.float32)
    d = c + 2 * (c // 16)
    flat = ro.cluster_features(feat)
    centres = flat[rng.integers(0, flat.shape[0], k)] + 0.3 * rng.standard_normal((k, d)).astype(np.float32)
    ids_o, dis = ro.assign_clusters(feat, centres, size, k)
    ids = region.assign_clusters(t(feat), t(centres.astype(np.float32)), size, k).cpu().numpy()
    srt = np.sort(dis, axis=1)
    assert ((srt[:, 1] - srt[:, 0]) > 2e-5 * srt[:, 0]).all()    # no near-ties (beyond fp32 summation error) in this draw
    assert np.array_equal(ids, ids_o)


def test_region_rejects_bad_arguments():
    each = torch.rand(2, 8, 8, device=DEV)
    ids = torch.zeros(2, 8, 8, device=DEV, dtype=torch.int64)
    with pytest.raises(ValueError):
        region.region_attention(each, ids.int(), 4)
    with pytest.raises(ValueError):
        region.region_attention(each[:, :2, :2], ids[:, :2, :2], 4)
    with pytest.raises(ValueError):
        region.region_attention(each, ids, 4, validate=True)          # sample 1 points into sample 0's clusters
    with pytest.raises(RuntimeError):
        region.region_attention(each.cpu(), ids.cpu(), 4)             # no CPU path
    with pytest.raises(ValueError):
        region.assign_clusters(torch.randn(1, 32, 4, 4, device=DEV), torch.randn(3, 32, device=DEV), 4)
    f, s, r, tv = region.region_attention(each[:0], ids[:0], 4)
    assert f.shape == (0, 1, 8, 8) and s.shape == (0, 8, 8)


def test_region_mask_properties_at_cfg3_shape():
    """B=64, 64x64 attention on a 512-channel 64x64 feature, 20 clusters (BASELINE cfg3 blend layer 13)."""
    torch.manual_seed(3)
    b, c, h, k = 64, 512, 64, 20
    feat = torch.randn(b, c, h, h, device=DEV)
    d = c + 2 * (c // 16)
    centres = torch.randn(k, d, device=DEV) * 0.2
    ids = region.assign_clusters(feat, centres, h, k)
    own = torch.arange(b, device=DEV).view(b, 1, 1) * k
    assert bool(((ids >= own) & (ids < own + k)).all())
    # the chosen centre is the nearest one: fp64 distances of a pixel subset, evaluated by plain torch
    pix = torch.randint(0, h * h, (256,), device=DEV)
    pos = torch.arange(h, device=DEV).float() * 2 / float(h - 1) - 1
    f = feat[5].reshape(c, -1)[:, pix].t().double()
    full = torch.cat([f, pos[pix % h].double()[:, None].expand(-1, c // 16), pos[pix // h].double()[:, None].expand(-1, c // 16)], 1)
    dist = ((full[:, None, :] - centres.double()[None]) ** 2).sum(-1)
    chosen = (ids[5].reshape(-1)[pix] - 5 * k)
    assert bool((dist.gather(1, chosen[:, None])[:, 0] <= dist.min(1).values * (1 + 1e-6)).all())
    # constant attention: every mean, the painted map and (above the threshold) the blurred map equal it
    const = torch.full((b, h, h), 0.9, device=DEV)
    final, same, reg, tv = region.region_attention(const, ids, k)
    assert float((same - 0.9).abs().max()) < 1e-6 and float((final - 0.9).abs().max()) < 2e-6 and float(tv) < 1e-12
    final_lo, _, reg_lo, _ = region.region_attention(const * 0.5, ids, k)
    assert float(final_lo.abs().max()) == 0.0 and float(reg_lo) == 0.0           # all below threshold / margin
    # batch invariance (fixed reduction order): a permuted batch gives bitwise-permuted results
    each = torch.rand(b, h, h, device=DEV) * 0.5 + 0.5
    perm = torch.randperm(b, device=DEV)
    ids_local = ids - own
    f1, s1, _, _ = region.region_attention(each, ids, k)
    f2, s2, _, _ = region.region_attention(each[perm], ids_local[perm] + own, k)
    assert torch.equal(f1[perm], f2) and torch.equal(s1[perm], s2)
    # gradient of loss_reg alone: 1 / (B * count) on clusters whose mean exceeds the margin, else 0
    e = each.clone().requires_grad_(True)
    _, same_e, reg_e, _ = region.region_attention(e, ids, k)
    reg_e.sum().backward()
    cnt = torch.zeros(b * k, device=DEV).index_add_(0, ids.reshape(-1), torch.ones(b * h * h, device=DEV))
    expect = torch.where(same_e > 0.7, 1.0 / (b * cnt[ids]), torch.zeros_like(same_e))
    assert float((e.grad - expect).abs().max()) < 1e-9
    # the mask is a partition-wise mean: total attention mass is preserved
    assert abs(float(same_e.double().sum() - each.double().sum())) < 5e-2
