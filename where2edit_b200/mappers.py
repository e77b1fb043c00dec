"""Latent mappers of the StyleCLIP-style edit loop: the caller-side MLP stacks
whose output `w_hat = w + 0.1 * mapper(w)` feeds Generator.forward (mapper/styleclip_mapper.py, BASELINE cfg3).

Mirrors mapper/latent_mappers.py:10-82 — same class names, constructor (`opts` with no_coarse_mapper /
no_medium_mapper / no_fine_mapper), attribute names (including the reference's `course_mapping` spelling, so
its checkpoints load) and forward semantics — on this package's EqualLinear / PixelNorm, whose fused
bias + leaky-ReLU runs in libw2e (w2e_bias_act_fwd/bwd); the 512x512 products are plain library GEMMs.
The reference's own file also works unmodified after `install_as_reference()`, since it only imports
`EqualLinear, PixelNorm` from models.stylegan2.model (:5).
"""
import torch
from torch import nn

from .model import EqualLinear, PixelNorm


class Mapper(nn.Module):
    """mapper/latent_mappers.py:10-29: PixelNorm (over dim 1) + 4 x EqualLinear(lr_mul=0.01, fused_lrelu)."""

    def __init__(self, opts, latent_dim=512):
        super().__init__()
        self.opts = opts
        layers = [PixelNorm()]
        for _ in range(4):
            layers.append(EqualLinear(latent_dim, latent_dim, lr_mul=0.01, activation="fused_lrelu"))
        self.mapping = nn.Sequential(*layers)

    def forward(self, x):
        return self.mapping(x)


class SingleMapper(nn.Module):
    """mapper/latent_mappers.py:32-44."""

    def __init__(self, opts):
        super().__init__()
        self.opts = opts
        self.mapping = Mapper(opts)

    def forward(self, x):
        return self.mapping(x)


class LevelsMapper(nn.Module):
    """mapper/latent_mappers.py:47-82: separate mappers for W+ rows [0,4), [4,8), [8,...); disabled levels
    contribute zeros."""

    def __init__(self, opts):
        super().__init__()
        self.opts = opts
        if not opts.no_coarse_mapper:
            self.course_mapping = Mapper(opts)
        if not opts.no_medium_mapper:
            self.medium_mapping = Mapper(opts)
        if not opts.no_fine_mapper:
            self.fine_mapping = Mapper(opts)

    def forward(self, x):
        x_coarse, x_medium, x_fine = x[:, :4, :], x[:, 4:8, :], x[:, 8:, :]
        x_coarse = torch.zeros_like(x_coarse) if self.opts.no_coarse_mapper else self.course_mapping(x_coarse)
        x_medium = torch.zeros_like(x_medium) if self.opts.no_medium_mapper else self.medium_mapping(x_medium)
        x_fine = torch.zeros_like(x_fine) if self.opts.no_fine_mapper else self.fine_mapping(x_fine)
        return torch.cat([x_coarse, x_medium, x_fine], dim=1)


class _TextCA(nn.Module):
    """Parameter container of the reference's CA_NET (utils.py:199-223): the cluster-style mapper constructs one
    per mapped layer (run_attention.py:717) but never calls it; kept so that its checkpoints load strictly."""

    def __init__(self, t_dim, c_dim):
        super().__init__()
        self.fc = nn.Linear(t_dim, c_dim * 4, bias=True)


class ClusterStyleMapper(nn.Module):
    """`FullSpaceMapperFEATClusterLinStyle_Net` (attention/run_attention.py:703-893) on this package's modules:
    per-layer style mappers conditioned on the text feature, 1x1 `StyledConv` attention heads over the captured
    generator features (stylespace inputs from `attention_textca_*`), and the region-mask block through
    `region.assign_clusters` / `region.region_attention` instead of the inline code at :775-794 / :852-884.
    Same constructor arguments, attribute / state-dict names and return value
    `(styles, final_attention_map, [loss_delta, loss_reg, loss_tv])`.

    End-to-end parity against the reference class: tests/test_variants_gpu.py (golden written by the
    unmodified reference).  The reference trains the attention heads, so their convolution-weight gradients are
    switched on here (the frozen generator's stay off)."""

    fused_heads = True    # A/B switch: False = one StyledConv module call per attention head (_attend)
    sample_first = True   # A/B switch of _attend: False = the heads run on the full-resolution feature maps (as the reference)
    LAYER_NUM = (0, 2, 3, 5, 6, 8, 9, 11, 12, 14, 15, 17, 18, 20, 21, 23, 24)
    STYLE_LAYERS = (0, 2, 2, 3, 5, 5, 6, 8, 8, 9, 11, 11, 12, 14, 14, 15, 17, 17, 18, 20, 20, 21, 23, 23, 24, 26, 26)

    def __init__(self, layers, in_dim=512, latent_dim=512, attention_layer=11, cluster_layer=11, channel_multiplier=1,
                 clusters=10, cluster_dim=512):
        super().__init__()
        from .model import StyledConv
        cm = channel_multiplier
        dim = [512] * 12 + [256 * cm] * 3 + [128 * cm] * 3 + [64 * cm] * 3 + [32 * cm] * 3 + [16 * cm] * 3
        self.dim = dim
        self.mapper_layer = self.STYLE_LAYERS[attention_layer]
        for c in range(layers + int((layers - 2) * 0.5)):
            if c < self.mapper_layer:
                setattr(self, f"mapper_{c}", EqualLinear(dim[c], dim[c], bias_init=1))
                setattr(self, f"mapper_textca_{c}", _TextCA(latent_dim, latent_dim))
                hidden = (latent_dim + 512) // 2
                setattr(self, f"mapper_text_{c}", nn.Sequential(
                    EqualLinear(latent_dim, hidden, lr_mul=1, activation="fused_lrelu"),
                    EqualLinear(hidden, 512, lr_mul=1, activation="fused_lrelu")))
                setattr(self, f"mapper_all_{c}", EqualLinear(dim[c] + 512, dim[c], bias_init=1))
            if c in self.LAYER_NUM:
                setattr(self, f"attention_textca_{c}", EqualLinear(latent_dim, dim[c + 1], bias_init=1))
                setattr(self, f"attention_{c}", StyledConv(dim[c + 1], 32, 1, dim[c + 1]))
        self.attention_textca_first = EqualLinear(latent_dim, dim[0], bias_init=1)
        self.attention_first = StyledConv(dim[0], 32, 1, dim[0])
        self.attention_textca_last = EqualLinear(latent_dim, 32 * layers, bias_init=1)
        self.attention_last = StyledConv(32 * layers, 1, 1, 32 * layers)
        self.initial_bias = nn.Parameter(torch.full((1,), 5.0))
        self.latent_dim = latent_dim
        self.register_buffer("initial_state", torch.randn(clusters, cluster_dim))
        self.cluster_layer = cluster_layer
        self.clusters = clusters
        from .model import ModulatedConv2d
        for m in self.modules():   # trainable 1x1 StyledConv heads: weight, noise weight and bias all get gradients
            if isinstance(m, ModulatedConv2d):
                m.weight_grad = True

    def store_clusters(self, initial_state):
        if tuple(initial_state.shape) != tuple(self.initial_state.shape):
            raise ValueError(f"cluster centres must be {tuple(self.initial_state.shape)}, got {tuple(initial_state.shape)}")
        self.initial_state = initial_state.to(self.initial_bias.device)

    def _attend(self, name, feature, text, size):
        """One attention head: 1x1 StyledConv over a captured feature map, nearest-resized to `size`
        (run_attention.py:803-806, 829-839).  A 1x1 modulated convolution, its noise / bias / leaky-ReLU and a
        nearest-neighbour DOWN-sampling commute exactly (every output pixel depends on one input pixel), so a feature
        map larger than `size` is sampled FIRST: the head then runs on size^2 pixels instead of up to 1024^2 (the
        reference evaluates all of them and keeps one in 256).  Fresh noise (noise=None, model.py:286-288) is drawn for
        the pixels that are kept.  Feature maps smaller than `size` go through the head first, as in the reference."""
        style = getattr(self, f"attention_textca_{name}")(text)
        if self.sample_first and size is not None and feature.shape[-2] > size and feature.shape[-1] > size:
            feature = torch.nn.functional.interpolate(feature, size)
        res, _ = getattr(self, f"attention_{name}")(feature, style.view(style.shape[0], 1, -1, 1, 1), input_is_stylespace=True)
        if size is None or tuple(res.shape[-2:]) == (size, size):
            return res
        return torch.nn.functional.interpolate(res, size)

    def _attend_all(self, names, features, text, size):
        """Every attention head in ONE launch (region.attention_heads -> w2e_attn_heads_fwd / _bwd): the 18 stylespace
        inputs from one GEMM over the concatenated `attention_textca_*` weights (they all read the same text feature),
        the 1x1 StyledConvs + nearest resize + concatenation from one kernel that only evaluates the surviving pixels."""
        from . import region
        lins = [getattr(self, f"attention_textca_{n}") for n in names]
        w = torch.cat([m.weight * m.scale for m in lins])
        bias = torch.cat([m.bias * m.lr_mul for m in lins])
        styles = torch.nn.functional.linear(text, w, bias).split([m.weight.shape[0] for m in lins], dim=1)
        heads = [getattr(self, f"attention_{n}") for n in names]
        weights = [h.conv.weight[0, :, :, 0, 0] * h.conv.scale for h in heads]
        return region.attention_heads(features, weights, list(styles), [h.activate.bias for h in heads],
                                      [h.noise.weight for h in heads], size)

    def _text_branches(self, x_text, n):
        """`mapper_text_{c}(x_text).unsqueeze(1)` for c < n (run_attention.py:816-817): every branch reads the same text
        feature, so the n first linears are ONE GEMM over the concatenated weights and the n second ones one batched
        GEMM, each followed by one fused bias + leaky-ReLU launch over all branches."""
        if n == 0:
            return []
        seqs = [getattr(self, f"mapper_text_{c}") for c in range(n)]
        if not (self.fused_heads and x_text.is_cuda):
            return [s(x_text).unsqueeze(1) for s in seqs]
        from .op.fused_act import fused_leaky_relu
        b = x_text.shape[0]
        hid = seqs[0][0].weight.shape[0]
        w1 = torch.cat([s[0].weight * s[0].scale for s in seqs])
        h1 = fused_leaky_relu(torch.nn.functional.linear(x_text, w1), torch.cat([s[0].bias * s[0].lr_mul for s in seqs]))
        w2 = torch.stack([s[1].weight * s[1].scale for s in seqs])                       # [n, 512, hid]
        o2 = torch.bmm(h1.view(b, n, hid).transpose(0, 1), w2.transpose(1, 2))           # [n, B, 512]
        o2 = fused_leaky_relu(o2.transpose(0, 1).reshape(b, -1), torch.cat([s[1].bias * s[1].lr_mul for s in seqs]))
        return list(o2.view(b, n, 1, -1).unbind(1))

    def forward(self, x, feature_map, size, attention_text=None):
        from . import region
        batch = x[0].shape[0]
        x_text = x[0][:, 0, :self.latent_dim]
        if attention_text is None:
            attention_text = x_text
        choice_cluster = region.assign_clusters(feature_map[self.cluster_layer - 1], self.initial_state, size, self.clusters)
        names = ["first"] + [c for c in range(len(x)) if c in self.LAYER_NUM]
        head_feats = [feature_map[-1]] + [feature_map[c] for c in names[1:]]
        fused = (self.fused_heads and len(names) <= 24 and
                 all(f.is_cuda and not f.requires_grad and f.ndim == 4 and f.shape[2] == f.shape[3] for f in head_feats))
        maps = [] if fused else [self._attend("first", feature_map[-1], attention_text, size)]
        text_hiddens = self._text_branches(x_text, min(len(x), self.mapper_layer))
        styles, loss_delta = [], 0
        for c in range(len(x)):
            x_c = x[c][:, :, self.latent_dim:]
            if c < self.mapper_layer:
                text_hidden = text_hiddens[c]
                mixed = getattr(self, f"mapper_all_{c}")(torch.cat([getattr(self, f"mapper_{c}")(x_c), text_hidden], dim=-1))
                x_new = x_c + 0.1 * (mixed - x_c)
                loss_delta = loss_delta + torch.mean(torch.norm(x_new - x_c, dim=-1)) / float(self.mapper_layer)
                x_c = x_new
            styles.append(x_c.unsqueeze(3).unsqueeze(3))
            if c in self.LAYER_NUM and not fused:
                maps.append(self._attend(c, feature_map[c], attention_text, size))
        all_maps = self._attend_all(names, head_feats, attention_text, size) if fused else torch.cat(maps, dim=1)
        logits = self._attend("last", all_maps, attention_text, None)
        each = torch.sigmoid(logits + self.initial_bias).view(batch, size, size)
        final_map, _, loss_reg, loss_tv = region.region_attention(each, choice_cluster, self.clusters)
        return styles, final_map, [loss_delta, loss_reg, loss_tv]


__all__ = ["Mapper", "SingleMapper", "LevelsMapper", "ClusterStyleMapper"]
