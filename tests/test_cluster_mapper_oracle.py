"""oracle/cluster_mapper_oracle.py (the cluster-style mapper's whole forward: style mappers, text-conditioned 1x1
StyledConv attention heads on the generator's captured features, region mask, losses) against the unmodified
reference class (tests/golden/cluster_mapper.npz, oracle/make_cluster_mapper_golden.py)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cluster_mapper_oracle as cmo  # noqa: E402
from oracle import stylegan2_oracle as orc  # noqa: E402
from oracle import synth  # noqa: E402


@pytest.fixture(scope="module")
def captured():
    sd = synth.make_state_dict(32, seed=0, perturbed=True)
    wplus = synth.make_wplus(2, 8, seed=2)
    _, _, styles, feats = orc.generator_forward_ref(sd, [wplus], 32, input_is_latent=True, return_features=True)
    feats = list(feats) + [sd["input.input"].repeat(2, 1, 1, 1)]           # run_attention.py:1110
    return styles, feats


@pytest.mark.parametrize("name", ["same_res", "upsampled"])
def test_cluster_mapper_oracle_matches_reference(captured, name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "cluster_mapper.npz"))
    styles, feats = captured
    size, clusters, cluster_layer, attention_layer = (int(v) for v in g[f"{name}/cfg"])
    text = torch.from_numpy(g[f"{name}/text"])
    x = [torch.cat([text.unsqueeze(1), s[:, :, :, 0, 0]], dim=-1) for s in styles]     # run_attention.py:1240
    state = cmo.SeededState()
    out, final, (loss_delta, loss_reg, loss_tv), extras = cmo.cluster_mapper_forward(
        state, x, feats, size, torch.from_numpy(g[f"{name}/centres"]), float(g[f"{name}/bias"]), layers=8,
        attention_layer=attention_layer, cluster_layer=cluster_layer)
    got = torch.stack([s[:, 0, :, 0, 0] for s in out]).numpy()
    want = g[f"{name}/styles_out"]
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-5 * np.abs(want).max())
    np.testing.assert_allclose(final.numpy(), g[f"{name}/final"], rtol=0, atol=5e-6)
    np.testing.assert_allclose([float(loss_delta), float(loss_reg), float(loss_tv)], g[f"{name}/losses"], rtol=2e-5)
    assert len(np.unique(extras["ids"])) > clusters                    # several clusters live in both samples
    # every parameter the restatement asked for exists in the reference module with that shape
    ref_keys = json.loads(str(g[f"{name}/keys"]))
    for key, value in state.used.items():
        assert key in ref_keys and ref_keys[key] == list(value.shape), key
    # ... and it asked for everything the reference forward uses (all but the unused CA_NET branches and buffers)
    unused = [k for k in ref_keys if k not in state.used and not k.startswith("mapper_textca_")
              and k not in ("initial_bias", "initial_state")]
    assert unused == [], unused[:5]
