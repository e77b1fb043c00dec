"""One whole edit step through the rows of SURVEY.md section 8 in the order the reference runs them
(attention/run_attention.py:1097-1110, 1240-1259): original forward with feature capture -> cluster assignment ->
region attention map -> blended edited forward -> CLIP pre-resample, composed on the GPU and compared with the
same composition of the CPU oracles; then a backward through everything."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import where2edit_b200 as w2e  # noqa: E402
from oracle import region_oracle, resample_oracle, synth  # noqa: E402
from oracle import stylegan2_oracle as orc  # noqa: E402
from where2edit_b200 import region, resample  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_edit_step_matches_composed_oracles():
    size, k, layer = 32, 6, 7
    sd = synth.make_state_dict(size, seed=0, perturbed=True)
    gen = w2e.Generator(size, 512, 8)
    gen.load_state_dict(sd, strict=True)
    gen = gen.to(DEV).eval()
    wplus = synth.make_wplus(2, gen.n_latent, seed=2)
    rng = np.random.default_rng(4)

    # ---- oracle composition (CPU)
    _, _, styles_o, feats_o = orc.generator_forward_ref(sd, [wplus], size, input_is_latent=True, return_features=True)
    blend_o = feats_o[layer - 1].numpy()
    s = blend_o.shape[-1]
    centres = region_oracle.cluster_features(blend_o)[rng.integers(0, 2 * s * s, k)] + \
        0.05 * rng.standard_normal((k, blend_o.shape[1] + 2 * (blend_o.shape[1] // 16))).astype(np.float32)
    centres = centres.astype(np.float32)
    ids_o, _ = region_oracle.assign_clusters(blend_o, centres, s, k)
    logits = (1.4 + rng.standard_normal((2, s, s))).astype(np.float32)
    each_o = 1 / (1 + np.exp(-logits.astype(np.float64)))
    final_o, _, reg_o, tv_o = region_oracle.region_attention(each_o.astype(np.float32), ids_o, k)
    edited_o = [st * (1 + 0.05 * synth.make_tensor(tuple(st.shape), 700 + i)) for i, st in enumerate(styles_o)]
    img_o, _ = orc.generator_forward_ref(sd, [edited_o], size, input_is_stylespace=True, attention_layer=layer,
                                         attention_map=torch.from_numpy(final_o), feature_map=feats_o)
    clip_o = resample_oracle.clip_resample(img_o.numpy(), 7, size // 32)

    # ---- the same step on the GPU
    with torch.no_grad():
        _, _, styles, feats = gen([wplus.to(DEV)], input_is_latent=True, randomize_noise=False, return_features=True)
    ids = region.assign_clusters(feats[layer - 1], torch.from_numpy(centres).to(DEV), s, k)
    assert np.array_equal(ids.cpu().numpy(), ids_o)
    lg = torch.from_numpy(logits).to(DEV).requires_grad_(True)
    final, _, reg, tv = region.region_attention(torch.sigmoid(lg), ids, k)
    assert float((final.detach().cpu() - torch.from_numpy(final_o)).abs().max()) <= 2e-6
    assert (final_o > 0).any() and (final_o == 0).any()            # both sides of the threshold in this draw
    edited = [(st * (1 + 0.05 * synth.make_tensor(tuple(st.shape), 700 + i).to(DEV))).requires_grad_(True)
              for i, st in enumerate(styles)]
    img, _ = gen([edited], input_is_stylespace=True, randomize_noise=False, attention_layer=layer,
                 attention_map=final, feature_map=feats)
    c = float(img_o.abs().max())
    assert float((img.detach().cpu() - img_o).abs().max()) <= 1e-4 * max(c, 1.0)
    clip = resample.ClipResample(size)(img)
    assert clip.shape == (2, 3, 224, 224)
    assert float(np.abs(clip.detach().cpu().numpy() - clip_o).max()) <= 1e-4 * max(c, 1.0)

    # ---- backward through resample -> generator (blend) -> region mask -> sigmoid, and to the styles
    (clip.square().mean() + 0.1 * reg.sum() + 0.1 * tv).backward()
    assert torch.isfinite(lg.grad).all() and float(lg.grad.abs().max()) > 0
    assert all(torch.isfinite(e.grad).all() for e in edited) and float(edited[0].grad.abs().max()) > 0
