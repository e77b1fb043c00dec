"""CPU restatement (torch CPU ops + the numpy region oracle) of the cluster-style mapper's whole forward —
TEST INFRASTRUCTURE (only tests/, smoke() and bench.py's cpu_baseline leg may import it).

Follows `FullSpaceMapperFEATClusterLinStyle_Net` (attention/run_attention.py:703-893): constructor layout
:705-747, forward :755-893 — style mappers (`mapper_text_c`, `mapper_c`, `mapper_all_c`, :812-825), the 1x1
`StyledConv` attention heads on the captured generator features with text-conditioned stylespace inputs
(:801-806, :829-840, :842-849), the region-mask block (oracle/region_oracle.py) and the three losses.
Pinned by tests/golden/cluster_mapper.npz (oracle/make_cluster_mapper_golden.py runs the unmodified reference
class with the seeded parameters generated here).

Parameters are not stored (23 M floats): `seeded_value(key, shape)` derives every tensor from its state-dict key,
the golden script writes the same values into the reference module, and the test checks that every key / shape
this restatement asks for exists in the reference's own state_dict.
"""
import zlib

import numpy as np
import torch
import torch.nn.functional as F

from . import region_oracle
from . import stylegan2_oracle as orc

LAYER_NUM = [0, 2, 3, 5, 6, 8, 9, 11, 12, 14, 15, 17, 18, 20, 21, 23, 24]                     # :710
STYLE_LAYERS = [0, 2, 2, 3, 5, 5, 6, 8, 8, 9, 11, 11, 12, 14, 14, 15, 17, 17, 18, 20, 20, 21, 23, 23, 24, 26, 26]   # :711


def seeded_value(key, shape, seed=0):
    """Deterministic fp32 tensor for a state-dict key: weights ~ N(0,1); modulation biases 1 + 0.1 N; other biases
    0.1 N; noise weights 0 (StyledConv draws fresh noise when none is given, model.py:286-288 — a zero weight keeps
    the forward deterministic, as at the reference's initialisation)."""
    rng = np.random.Generator(np.random.PCG64(zlib.crc32(key.encode()) + 7919 * seed))
    if key.endswith("noise.weight"):
        return torch.zeros(shape)
    v = torch.from_numpy(rng.standard_normal(tuple(shape)).astype(np.float32))
    if key.endswith("modulation.bias"):
        return 1 + 0.1 * v
    if key.endswith("bias"):
        return 0.1 * v
    return v


class SeededState:
    """Lazy state dict: `get(key, shape)` generates (and records) the seeded tensor of a key."""

    def __init__(self, seed=0):
        self.seed = seed
        self.used = {}

    def get(self, key, shape):
        shape = tuple(int(s) for s in shape)
        if key not in self.used:
            self.used[key] = seeded_value(key, shape, self.seed)
        assert tuple(self.used[key].shape) == shape, key
        return self.used[key]


def mapper_dims(channel_multiplier=1):
    cm = channel_multiplier
    return [512] * 12 + [256 * cm] * 3 + [128 * cm] * 3 + [64 * cm] * 3 + [32 * cm] * 3 + [16 * cm] * 3   # :709


def _linear(state, prefix, x, in_dim, out_dim, activation=None):
    """EqualLinear(in_dim, out_dim, lr_mul=1[, activation]) (models/stylegan2/model.py:130-159)."""
    return orc.equal_linear_ref(x, state.get(f"{prefix}.weight", (out_dim, in_dim)), state.get(f"{prefix}.bias", (out_dim,)),
                                1.0, activation)


def _head(state, prefix, feature, cin, cout, style):
    """StyledConv(cin, cout, 1, cin) called with a stylespace input and no noise argument (:805, :837, :845)."""
    sd = {
        f"{prefix}.conv.weight": state.get(f"{prefix}.conv.weight", (1, cout, cin, 1, 1)),
        f"{prefix}.conv.modulation.weight": state.get(f"{prefix}.conv.modulation.weight", (cin, cin)),
        f"{prefix}.conv.modulation.bias": state.get(f"{prefix}.conv.modulation.bias", (cin,)),
        f"{prefix}.noise.weight": state.get(f"{prefix}.noise.weight", (1,)),
        f"{prefix}.activate.bias": state.get(f"{prefix}.activate.bias", (cout,)),
    }
    b, _, h, w = feature.shape
    out, _ = orc._styled_conv(sd, prefix, feature, style.view(b, 1, -1, 1, 1), torch.zeros(b, 1, h, w), False, True)
    return out


def cluster_mapper_forward(state, x, feature_map, size, initial_state, initial_bias, layers, latent_dim=512,
                           attention_layer=11, cluster_layer=11, channel_multiplier=1, attention_text=None):
    """forward(x, feature_map, size, attention_text) (:755-893).  x: list of [B,1,latent_dim + dim_c] tensors.
    Returns (out list of [B,1,dim_c,1,1], final_attention_map [B,1,size,size], [loss_delta, loss_reg, loss_tv],
    extras dict with the per-pixel attention and the cluster ids)."""
    dim = mapper_dims(channel_multiplier)
    mapper_layer = STYLE_LAYERS[attention_layer]                                                   # :712
    clusters = initial_state.shape[0]
    batch = x[0].shape[0]
    x_text = x[0][:, 0, :latent_dim]
    if attention_text is None:
        attention_text = x_text
    ids, _ = region_oracle.assign_clusters(feature_map[cluster_layer - 1].numpy(), initial_state.numpy(), size, clusters)

    out = []
    feature = feature_map[-1]
    x_text_ca = _linear(state, "attention_textca_first", attention_text, latent_dim, dim[0])
    feature_res = _head(state, "attention_first", feature, dim[0], 32, x_text_ca)
    attention_feature = [F.interpolate(feature_res, size)]
    loss_delta = 0
    for c in range(len(x)):
        x_c = x[c][:, :, latent_dim:]
        if c < mapper_layer:
            h = _linear(state, f"mapper_text_{c}.0", x_text, latent_dim, (latent_dim + 512) // 2, "fused_lrelu")
            x_text_hidden = _linear(state, f"mapper_text_{c}.1", h, (latent_dim + 512) // 2, 512, "fused_lrelu").unsqueeze(1)
            x_c_hidden = _linear(state, f"mapper_{c}", x_c, dim[c], dim[c])
            mixed = _linear(state, f"mapper_all_{c}", torch.cat([x_c_hidden, x_text_hidden], dim=-1), dim[c] + 512, dim[c])
            x_c_new = x_c + 0.1 * (mixed - x_c)
            loss_delta = loss_delta + torch.mean(torch.norm(x_c_new - x_c, dim=-1)) / float(mapper_layer)
            out.append(x_c_new.unsqueeze(3).unsqueeze(3))
        else:
            out.append(x_c.unsqueeze(3).unsqueeze(3))
        if c in LAYER_NUM:
            x_text_ca = _linear(state, f"attention_textca_{c}", attention_text, latent_dim, dim[c + 1])
            feature_res = _head(state, f"attention_{c}", feature_map[c], dim[c + 1], 32, x_text_ca)
            attention_feature.append(F.interpolate(feature_res, size))
    each = torch.cat(attention_feature, dim=1)
    x_text_ca = _linear(state, "attention_textca_last", attention_text, latent_dim, 32 * layers)
    each = _head(state, "attention_last", each, 32 * layers, 1, x_text_ca)
    each = torch.sigmoid(each + initial_bias).view(batch, size, size)
    final, same, loss_reg, loss_tv = region_oracle.region_attention(each.numpy(), ids, clusters)
    return out, torch.from_numpy(final), [loss_delta, torch.from_numpy(loss_reg), torch.tensor(loss_tv)], \
        {"each": each, "ids": ids, "same": same}
