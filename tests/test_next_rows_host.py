"""Host-side logic of the SURVEY.md section 8f rows that needs no GPU: blur taps, module hyper-parameters,
state-dict layout, and the loud failure on CPU tensors (there is no CPU path)."""
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
from oracle import region_oracle as ro  # noqa: E402
from where2edit_b200 import mappers, region, resample  # noqa: E402


def test_gaussian_taps_equal_torchvision_and_oracle():
    taps = np.array(region.gaussian_taps(), np.float32).reshape(5, 5)
    assert np.array_equal(taps, ro.gaussian_kernel2d())
    tvf = pytest.importorskip("torchvision.transforms._functional_tensor")
    k2 = tvf._get_gaussian_kernel2d([5, 5], [1.1, 1.1], dtype=torch.float32, device=torch.device("cpu"))
    np.testing.assert_allclose(taps, k2.numpy(), rtol=0, atol=1e-8)


def test_gaussian_blur_oracle_matches_torchvision():
    tv = pytest.importorskip("torchvision.transforms.functional")
    x = torch.rand(2, 1, 9, 9)
    ref = tv.gaussian_blur(x, 5).numpy()[:, 0]
    np.testing.assert_allclose(ro.gaussian_blur(x.numpy()[:, 0]), ref, rtol=0, atol=1e-6)


def test_no_cpu_path_for_the_next_rows():
    each, ids = torch.rand(1, 8, 8), torch.zeros(1, 8, 8, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="no CPU path"):
        region.region_attention(each, ids, 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        region.assign_clusters(torch.randn(1, 16, 4, 4), torch.randn(2, 18), 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        resample.clip_resample(torch.randn(1, 3, 32, 32), 7, 1)


def test_clip_resample_module_mirrors_cliploss_hyperparameters():
    m = resample.ClipResample(1024)
    assert (m.scale_factor, m.kernel_size) == (7, 32)          # criteria/clip_loss.py:10-11
    assert resample.ClipResample(256).kernel_size == 8


def test_levels_mapper_state_dict_layout():
    opts = types.SimpleNamespace(no_coarse_mapper=False, no_medium_mapper=False, no_fine_mapper=False)
    sd = mappers.LevelsMapper(opts).state_dict()
    want = mo.mapper_state()
    assert sorted(sd) == sorted(want) and all(tuple(sd[k].shape) == want[k].shape for k in want)
    opts.no_medium_mapper = True
    assert not any(k.startswith("medium_mapping") for k in mappers.LevelsMapper(opts).state_dict())
    ref_root = os.environ.get("W2E_REFERENCE", "/root/reference")
    if os.path.isdir(ref_root):     # build container only: the reference's own class has the same keys
        sys.path.insert(0, ref_root)
        try:
            from mapper.latent_mappers import LevelsMapper as RefLevels
        finally:
            sys.path.remove(ref_root)
        opts.no_medium_mapper = False
        ref_sd = RefLevels(opts).state_dict()
        assert sorted(ref_sd) == sorted(sd) and all(ref_sd[k].shape == sd[k].shape for k in sd)


def test_c_abi_rejects_bad_arguments_without_touching_the_gpu():
    """Argument validation happens before any CUDA call: null pointers / bad shapes return an error code and
    set w2e_last_error_string (no kernel is enqueued, so this runs on a CPU-only box)."""
    from where2edit_b200 import _native as N
    lib = N.load()
    one = 0x1000   # any non-null address: validation fails on the shapes before it could be dereferenced
    cases = [
        (lib.w2e_box_resample_fwd(None, None, 1, 8, 8, 7, 2, None), "box_resample_fwd"),
        (lib.w2e_box_resample_fwd(one, one, 1, 8, 8, 0, 2, None), "box_resample_fwd"),       # scale 0
        (lib.w2e_box_resample_bwd(one, one, 1, 2, 2, 1, 9, None), "box_resample_bwd"),       # window larger than the image
        (lib.w2e_cluster_assign(None, None, None, None, 1, 16, 4, 2, 1, 4, None), "cluster_assign"),
        (lib.w2e_cluster_assign(one, one, one, one, 1, 16, 1, 2, 1, 4, None), "cluster_assign"),   # h = 1
        (lib.w2e_region_mask_fwd(None, None, None, None, None, None, None, None, 1, 8, 2, 0.8, 0.7, None), "region_mask_fwd"),
        (lib.w2e_region_mask_fwd(one, one, one, one, one, one, one, one, 1, 2, 2, 0.8, 0.7, None), "region_mask_fwd"),  # S < 3
        (lib.w2e_region_mask_bwd(None, None, None, None, None, None, None, None, 1, 8, 2, 0.7, None), "region_mask_bwd"),
    ]
    for code, what in cases:
        assert code != 0, what
    assert "region_mask_bwd" in N.last_error()
    assert lib.w2e_box_resample_fwd(one, one, 0, 8, 8, 7, 2, None) == 0      # empty batch: nothing to do, no launch


def test_cluster_style_mapper_state_dict_equals_the_reference_layout():
    """Keys and shapes of mappers.ClusterStyleMapper == those of the reference class as recorded by
    oracle/make_cluster_mapper_golden.py (so reference checkpoints load with strict=True)."""
    import json
    g = np.load(os.path.join(ROOT, "tests", "golden", "cluster_mapper.npz"))
    for name in ("same_res", "upsampled"):
        size, clusters, cluster_layer, attention_layer = (int(v) for v in g[f"{name}/cfg"])
        m = mappers.ClusterStyleMapper(8, 1024, 512, attention_layer=attention_layer, cluster_layer=cluster_layer,
                                       clusters=clusters, cluster_dim=576)
        assert {k: list(v.shape) for k, v in m.state_dict().items()} == json.loads(str(g[f"{name}/keys"]))
        with pytest.raises(ValueError):
            m.store_clusters(torch.zeros(clusters + 1, 576))


def test_save_image_oracle_equals_torchvision():
    """oracle/save_image_oracle.py (the uint8 image output's checker) against torchvision.utils itself: make_grid's
    normalisation followed by save_image's conversion, on values that cover both clamps and every level"""
    import torchvision
    from oracle import save_image_oracle as so
    g = torch.Generator().manual_seed(3)
    img = (torch.rand(1, 3, 64, 96, generator=g) * 2.6 - 1.3)
    img[0, 0, 0, :8] = torch.tensor([-1.0, 1.0, 0.0, -0.999999, 0.999999, 1e-8, -1.5, 1.5])
    grid = torchvision.utils.make_grid(img.clone(), nrow=1, normalize=True, value_range=(-1, 1), padding=0)
    want = grid.mul(255).add_(0.5).clamp_(0, 255).to(torch.uint8).numpy()
    got = so.quantise_ref(img[0].numpy())
    assert got.dtype == np.uint8 and np.array_equal(got, want)
    assert len(np.unique(got)) == 256
    # the host-side helper of the package is the same arithmetic (used when the image did not come out of the fused epilogue)
    from where2edit_b200 import functional as K
    assert np.array_equal(K.quantize_u8(img[0]).numpy(), want)


def test_new_entry_points_reject_bad_arguments_without_touching_the_gpu():
    """w2e_modconv_tc2_rgb_pair / _dgrad_up / w2e_attn_heads_fwd/_bwd: validation before any CUDA call"""
    import ctypes
    from where2edit_b200 import _native as N
    lib = N.load()
    one = 0x1000
    INVALID = 1
    # pair kernel: the width must be a multiple of 16
    assert lib.w2e_modconv_tc2_rgb_pair(one, one, one, one, None, None, 0, None, None, 1, 64, 24, 1, one, one, None, None, None,
                                        one, 0, None, None) == INVALID
    assert "multiple of 16" in N.last_error()
    # fused transposed dgrad: null pointers
    assert lib.w2e_modconv_tc2_dgrad_up(None, None, None, None, None, 1, 32, 64, 32, 32, None, None) == INVALID
    # attention heads: head count, null tables, a head without feature map
    ptrs = (ctypes.c_void_p * 1)(one)
    null = (ctypes.c_void_p * 1)(None)
    ints = (ctypes.c_int * 1)(32)
    assert lib.w2e_attn_heads_fwd(0, ptrs, ints, ints, ptrs, ptrs, ptrs, None, None, one, one, 1, 16, None) == INVALID
    assert lib.w2e_attn_heads_fwd(25, ptrs, ints, ints, ptrs, ptrs, ptrs, None, None, one, one, 1, 16, None) == INVALID
    assert lib.w2e_attn_heads_fwd(1, None, ints, ints, ptrs, ptrs, ptrs, None, None, one, one, 1, 16, None) == INVALID
    assert lib.w2e_attn_heads_fwd(1, null, ints, ints, ptrs, ptrs, ptrs, None, None, one, one, 1, 16, None) == INVALID
    assert lib.w2e_attn_heads_fwd(1, ptrs, ints, ints, ptrs, ptrs, ptrs, None, ptrs, one, one, 1, 16, None) == INVALID  # noise, no weight
    assert "noise" in N.last_error()
    assert lib.w2e_attn_heads_bwd(1, ptrs, ints, ints, ptrs, ptrs, ptrs, None, None, one, one, one, one, one, None, None, 1, 16,
                                  None) == INVALID
