"""Region-mask construction of the cluster-style mapper: the step immediately before the blended
`Generator.forward`.

The reference has no function for it: the code is inline in
`FullSpaceMapperFEATClusterLinStyle_Net.forward` / `FullSpaceMapperFEATClusterLin_Net.forward`
(attention/run_attention.py:775-794 and :852-884; the older ...ClusterLin_Net has the same assignment without
the resize at :508-527 and the same mask block with margin 0.8, training mode only, at :560-587).  The two calls below
replace those two blocks (INTEGRATION.md section 6 shows the edit):

    choice_cluster = assign_clusters(feature_map[self.cluster_layer - 1], self.initial_state, size, self.clusters)
    final_attention_map, same_attention_map, loss_reg, loss_tv = region_attention(
        each_attention_map, choice_cluster, self.clusters)

Both run hand-written CUDA through libw2e's C ABI (include/w2e.h); there is no CPU path.
"""

import torch

# autocast safety (the reference's --amp wraps mapper + generator in torch.cuda.amp.autocast, run_attention.py:1231):
# the kernels take fp32 (or bf16) pointers, so half-precision tensors handed over by autocast-ed linears are cast to
# fp32 at every custom Function and autocast is off inside it
_amp_fwd = torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_amp_bwd = torch.amp.custom_bwd(device_type="cuda")

from . import _native as N

GAUSSIAN_KSIZE = 5


def gaussian_taps(ksize=GAUSSIAN_KSIZE):
    """torchvision.transforms.functional.gaussian_blur's default kernel for `ksize` (sigma = 0.15*k + 0.35),
    as a row-major [ksize*ksize] list of fp32 values (computed like torchvision: fp32 linspace/exp/normalise)."""
    sigma = ksize * 0.15 + 0.35
    half = (ksize - 1) * 0.5
    x = torch.linspace(-half, half, steps=ksize, dtype=torch.float32)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    k1 = pdf / pdf.sum()
    return (k1[:, None] * k1[None, :]).reshape(-1).tolist()


_TAPS = None


def _taps():
    global _TAPS
    if _TAPS is None:
        _TAPS = N.host_floats(gaussian_taps())
    return _TAPS


def assign_clusters(blend_feature, initial_state, size, clusters=None):
    """attention/run_attention.py:775-794.  blend_feature [B,C,h,h]; initial_state [K, C + 2*(C//16)] cluster
    centres over (feature channels, C//16 copies of the x position, C//16 copies of the y position).
    Returns int64 [B,size,size]: `b * K + argmin_k ||concat_feature - centre_k||^2`, nearest-resized to `size`.
    The [B*h*h, K, D] difference tensor of `pairwise_distance` (utils.py:244-263) is never materialised."""
    N.require_cuda(blend_feature, initial_state)
    if blend_feature.ndim != 4 or blend_feature.shape[2] != blend_feature.shape[3]:
        raise ValueError(f"blend_feature must be [B,C,h,h], got {tuple(blend_feature.shape)}")
    b, c, h, _ = blend_feature.shape
    pc = c // 16
    k = initial_state.shape[0]
    if clusters is not None and clusters != k:
        raise ValueError(f"clusters={clusters} but initial_state holds {k} centres")
    if initial_state.ndim != 2 or initial_state.shape[1] != c + 2 * pc:
        raise ValueError(f"initial_state must be [K,{c + 2 * pc}] for a {c}-channel feature, got {tuple(initial_state.shape)}")
    if h < 2:
        raise ValueError("cluster assignment needs a feature map of at least 2x2 (positions divide by h-1)")
    feat = blend_feature.detach().to(torch.float32).contiguous()
    ctr = initial_state.detach().to(torch.float32).contiguous()
    size = int(size)
    ids = torch.empty((b, size, size), device=feat.device, dtype=torch.int64)
    low = torch.empty((b, h, h), device=feat.device, dtype=torch.int32)
    N.check(N.load().w2e_cluster_assign(N.ptr(feat), N.ptr(ctr), N.ptr(low), N.ptr(ids), b, c, h, k, pc, size,
                                        N.stream_ptr()), "cluster_assign")
    return ids


class _RegionAttention(torch.autograd.Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, each, ids, clusters, threshold, margin):
        b, s, _ = each.shape
        dev = each.device
        final = torch.empty((b, 1, s, s), device=dev, dtype=torch.float32)
        same = torch.empty((b, s, s), device=dev, dtype=torch.float32)
        stats = torch.empty((b, clusters, 2), device=dev, dtype=torch.float32)
        parts = torch.empty((b, 2), device=dev, dtype=torch.float32)
        losses = torch.empty(2, device=dev, dtype=torch.float32)
        N.check(N.load().w2e_region_mask_fwd(N.ptr(each), N.ptr(ids), _taps(), N.ptr(final), N.ptr(same), N.ptr(stats),
                                             N.ptr(parts), N.ptr(losses), b, s, clusters, threshold, margin,
                                             N.stream_ptr()), "region_mask_fwd")
        ctx.save_for_backward(each, ids, same, stats)
        ctx.cfg = (clusters, margin)
        ctx.mark_non_differentiable(same)
        return final, same, losses[0:1].clone(), losses[1].clone()

    @staticmethod
    @torch.autograd.function.once_differentiable
    @_amp_bwd
    def backward(ctx, g_final, g_same, g_reg, g_tv):
        each, ids, same, stats = ctx.saved_tensors
        clusters, margin = ctx.cfg
        b, s, _ = each.shape
        g_final = None if g_final is None else g_final.to(torch.float32).contiguous()
        g_losses = None
        if g_reg is not None or g_tv is not None:
            zero = torch.zeros(1, device=each.device, dtype=torch.float32)
            g_losses = torch.cat([zero if g_reg is None else g_reg.reshape(1).float(),
                                  zero if g_tv is None else g_tv.reshape(1).float()])
        g_each = torch.empty_like(each)
        N.check(N.load().w2e_region_mask_bwd(N.ptr(g_final), N.ptr(g_losses), N.ptr(each), N.ptr(ids), N.ptr(same),
                                             N.ptr(stats), _taps(), N.ptr(g_each), b, s, clusters, margin,
                                             N.stream_ptr()), "region_mask_bwd")
        return g_each, None, None, None, None


def region_attention(each_attention_map, choice_cluster, clusters, threshold=0.8, margin=0.7, validate=False):
    """attention/run_attention.py:852-884.  each_attention_map [B,S,S] (sigmoid output), choice_cluster int64
    [B,S,S] as returned by `assign_clusters` (ids of sample b lie in [b*K, (b+1)*K)).  Returns

      final_attention_map [B,1,S,S]  gaussian_blur_5(where(same < threshold, same - same.detach(), same))
      same_attention_map  [B,S,S]    every pixel = mean attention of its cluster (detached, as in loss_tv)
      loss_reg            [1]        sum_b sum_k relu(mean_bk - margin) / B  (empty clusters skipped)
      loss_tv             []         mse(each_attention_map, same_attention_map.detach())

    differentiable w.r.t. each_attention_map.  The reference's B*K boolean-mask passes (one device sync each)
    become one launch; reductions run in a fixed order, so results do not depend on the batch composition.
    Pixels whose id is outside the sample's own range keep the reference's initial value 1.0 when the id
    matches no cluster at all; ids that point into ANOTHER sample's range are not supported (the reference
    never produces them) -- `validate=True` checks this (one device sync)."""
    N.require_cuda(each_attention_map, choice_cluster)
    if each_attention_map.ndim != 3 or each_attention_map.shape[1] != each_attention_map.shape[2]:
        raise ValueError(f"each_attention_map must be [B,S,S], got {tuple(each_attention_map.shape)}")
    if choice_cluster.shape != each_attention_map.shape or choice_cluster.dtype != torch.int64:
        raise ValueError("choice_cluster must be an int64 tensor of each_attention_map's shape")
    b, s, _ = each_attention_map.shape
    if s < 3:
        raise ValueError("the 5x5 reflect-padded blur needs S >= 3")
    clusters = int(clusters)
    if clusters < 1:
        raise ValueError("clusters must be positive")
    if b == 0:   # empty batch: nothing to launch (the reference would divide by zero in loss_reg)
        e = each_attention_map.to(torch.float32)
        return e.new_empty((0, 1, s, s)), e.new_empty((0, s, s)), e.new_zeros(1), e.new_zeros(())
    if validate:
        own = torch.arange(b, device=choice_cluster.device).view(b, 1, 1) * clusters
        foreign = (choice_cluster >= 0) & (choice_cluster < b * clusters) & \
                  ((choice_cluster < own) | (choice_cluster >= own + clusters))
        if bool(foreign.any()):
            raise ValueError("choice_cluster holds ids of another sample's clusters")
    each = each_attention_map.to(torch.float32).contiguous()
    return _RegionAttention.apply(each, choice_cluster.contiguous(), clusters, float(threshold), float(margin))


__all__ = ["assign_clusters", "region_attention", "gaussian_taps", "attention_heads"]


class _AttentionHeads(torch.autograd.Function):
    """All region-attention heads of the cluster-style mapper in one launch (csrc/attn_heads.cu): forward w2e_attn_heads_fwd,
    backward w2e_attn_heads_bwd (gradients to every head's weight, style, bias and noise weight; the feature maps come
    from the frozen generator and get none)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, size, feats, noises, *params):
        n = len(feats)
        weights, styles, biases, noise_ws = (list(params[i * n:(i + 1) * n]) for i in range(4))
        weights = [w.contiguous() for w in weights]
        styles = [s.contiguous() for s in styles]
        biases = [b.contiguous() for b in biases]
        noise_ws = [w.reshape(1).contiguous() for w in noise_ws]
        b = feats[0].shape[0]
        dev = feats[0].device
        out = torch.empty((b, 32 * n, size, size), device=dev, dtype=torch.float32)
        demod = torch.empty((n, b, 32), device=dev, dtype=torch.float32)
        chans = [int(f.shape[1]) for f in feats]
        res = [int(f.shape[2]) for f in feats]
        N.check(N.load().w2e_attn_heads_fwd(
            n, N.host_ptrs(feats), N.host_ints(chans), N.host_ints(res), N.host_ptrs(weights), N.host_ptrs(styles),
            N.host_ptrs(biases), N.host_ptrs(noise_ws), N.host_ptrs(noises), N.ptr(out), N.ptr(demod), b, size,
            N.stream_ptr()), "attn_heads_fwd")
        ctx.heads = (size, feats, noises, weights, styles, biases, noise_ws, chans, res)
        ctx.save_for_backward(out, demod)
        return out

    @staticmethod
    @_amp_bwd
    def backward(ctx, gout):
        size, feats, noises, weights, styles, biases, noise_ws, chans, res = ctx.heads
        out, demod = ctx.saved_tensors
        n = len(feats)
        b = out.shape[0]
        gout = gout.to(torch.float32).contiguous()
        gacc = torch.empty_like(out)
        red = torch.empty((n, b, 3, 32), device=out.device, dtype=torch.float32)
        dws = [torch.empty_like(w) for w in weights]
        dss = [torch.empty_like(s) for s in styles]
        N.check(N.load().w2e_attn_heads_bwd(
            n, N.host_ptrs(feats), N.host_ints(chans), N.host_ints(res), N.host_ptrs(weights), N.host_ptrs(styles),
            N.host_ptrs(biases), N.host_ptrs(noise_ws), N.host_ptrs(noises), N.ptr(gout), N.ptr(out), N.ptr(demod),
            N.ptr(gacc), N.ptr(red), N.host_ptrs(dws), N.host_ptrs(dss), b, size, N.stream_ptr()), "attn_heads_bwd")
        per_head = red.sum(1)                                   # [n, 3, 32]: the per-sample partials, summed over the batch
        dbias = [per_head[h, 0] for h in range(n)]
        dnw = [per_head[h, 2, 0:1] if noises[h] is not None else None for h in range(n)]
        return (None, None, None, *dws, *dss, *dbias, *dnw)


def attention_heads(feats, weights, styles, biases, noise_weights, size, noises=None):
    """Grouped launch of the 1x1 StyledConv attention heads + nearest resize + concatenation (attention/run_attention.py:
    803-806, 829-841).  feats[h]: [B,C_h,H_h,H_h] fp32 captured feature maps (no gradient: they come from the frozen
    generator under no_grad, :1196-1203); weights[h]: [32,C_h] with the equalised-lr scale folded in; styles[h]: [B,C_h];
    biases[h]: [32]; noise_weights[h]: scalar tensors; noises[h]: [B,r,r] with r = min(H_h, size), or None = fresh normal
    noise at the head's own resolution (model.py:286-288).  Returns [B, 32 * len(feats), size, size]."""
    N.require_cuda(*feats)
    size = int(size)
    fs = []
    for f in feats:
        if f.ndim != 4 or f.shape[2] != f.shape[3]:
            raise ValueError(f"attention_heads: feature maps must be [B,C,H,H], got {tuple(f.shape)}")
        if f.requires_grad:
            raise ValueError("attention_heads: feature maps must not require grad (use the per-head modules instead)")
        fs.append(f.detach().to(torch.float32).contiguous())
    b = fs[0].shape[0]
    if noises is None:
        noises = [f.new_empty(b, 1, min(f.shape[2], size), min(f.shape[2], size)).normal_() for f in fs]   # (as NoiseInjection)
    noises = [None if z is None else z.detach().to(torch.float32).contiguous() for z in noises]
    return _AttentionHeads.apply(size, fs, noises, *weights, *styles, *biases, *noise_weights)
