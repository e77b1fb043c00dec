"""CPU restatement (numpy, fp32) of the reference's region-mask construction — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
path (where2edit_b200/region.py -> libw2e.so) never does.

Follows, statement by statement:
  * `pairwise_distance`                      utils.py:244-263
  * cluster assignment                       attention/run_attention.py:775-794
  * per-cluster mean attention + losses      attention/run_attention.py:852-876
  * threshold with straight-through grad     attention/run_attention.py:882-883
  * `torchvision.transforms.functional.gaussian_blur(x, 5)` (third-party: torchvision, reference pin 0.8.2,
    `requirements.txt:199`; container 0.26 computes the same thing): sigma = 0.15*k + 0.35 = 1.1, taps
    exp(-x^2 / (2 sigma^2)) on linspace(-2, 2, 5) normalised, outer product, reflect padding 2, true 2-D conv.
Parity pinned by tests/golden/region.npz (oracle/make_region_golden.py runs the unmodified reference forward).
"""
import numpy as np

F32 = np.float32


def pairwise_distance(data1, data2):
    """utils.py:244-263: dis[n, k] = sum_m (data1[n, m] - data2[k, m])**2 (fp32)."""
    a = np.asarray(data1, F32)[:, None, :]
    b = np.asarray(data2, F32)[None, :, :]
    return ((a - b) ** F32(2.0)).sum(axis=-1, dtype=F32)


def nearest_index(out_size, in_size):
    """Source index of F.interpolate(mode='nearest'): floor(dst * (in / out)) in fp32, clamped."""
    scale = F32(in_size) / F32(out_size)
    idx = np.floor(np.arange(out_size, dtype=F32) * scale).astype(np.int64)
    return np.minimum(idx, in_size - 1)


def cluster_features(feature):
    """run_attention.py:776-790: [B, C, h, h] -> [B*h*h, C + 2*(C//16)] (feature ++ x positions ++ y positions)."""
    feature = np.asarray(feature, F32)
    b, c, h, _ = feature.shape
    pc = c // 16
    pos = np.arange(h, dtype=F32) * F32(2) / F32(h - 1) - F32(1)
    xpos = np.broadcast_to(pos[None, None, None, :], (b, pc, h, h))
    ypos = np.broadcast_to(pos[None, None, :, None], (b, pc, h, h))
    cat = np.concatenate([feature, xpos, ypos], axis=1)
    return np.ascontiguousarray(cat.transpose(0, 2, 3, 1)).reshape(-1, c + 2 * pc)


def assign_clusters(feature, centres, size, clusters):
    """run_attention.py:775-794: int64 [B, size, size] = b * clusters + argmin_k distance, nearest-resized."""
    b, _, h, _ = feature.shape
    dis = pairwise_distance(cluster_features(feature), centres)
    low = dis.argmin(axis=1).reshape(b, h, h) + np.arange(b)[:, None, None] * clusters
    idx = nearest_index(size, h)
    return low[:, idx][:, :, idx].astype(np.int64), dis


def gaussian_kernel2d(ksize=5):
    sigma = ksize * 0.15 + 0.35
    half = (ksize - 1) * 0.5
    x = np.linspace(-half, half, ksize, dtype=F32)
    pdf = np.exp(F32(-0.5) * (x / F32(sigma)) ** 2).astype(F32)
    k1 = (pdf / pdf.sum(dtype=F32)).astype(F32)
    return (k1[:, None] * k1[None, :]).astype(F32)


def gaussian_blur(x, ksize=5):
    """[B, S, S] -> [B, S, S]; reflect padding ksize//2, cross-correlation with the (symmetric) 2-D kernel."""
    k2 = gaussian_kernel2d(ksize)
    p = ksize // 2
    xp = np.pad(np.asarray(x, F32), ((0, 0), (p, p), (p, p)), mode="reflect")
    s = x.shape[-1]
    out = np.zeros_like(x, dtype=F32)
    for dy in range(ksize):
        for dx in range(ksize):
            out += k2[dy, dx] * xp[:, dy:dy + s, dx:dx + s]
    return out


def region_attention(each, ids, clusters, threshold=0.8, margin=0.7):
    """run_attention.py:852-883.  each [B, S, S] in (0, 1), ids [B, S, S] int64.
    Returns final_attention_map [B, 1, S, S], same_attention_map [B, S, S], loss_reg [1], loss_tv []."""
    each = np.asarray(each, F32)
    batch = each.shape[0]
    same = np.ones_like(each)
    cluster_attention = F32(0)
    batch_attention = F32(0)
    for i in range(batch * clusters):
        sel = ids == i
        if sel.any():
            m = each[sel].mean(dtype=F32)
            same[sel] = m
            cluster_attention = F32(cluster_attention + max(F32(m - F32(margin)), F32(0)))   # relu(mean - 0.7)
        if (i + 1) % clusters == 0:
            batch_attention = F32(batch_attention + cluster_attention)
            cluster_attention = F32(0)
    loss_reg = np.array([batch_attention / F32(batch)], F32)
    loss_tv = np.mean((each - same) ** 2, dtype=F32)
    thr = np.where(same < F32(threshold), F32(0), same)      # a - a.detach() == 0 below the threshold
    final = gaussian_blur(thr)[:, None]
    return final, same, loss_reg, F32(loss_tv)


def region_attention_backward(g_final, g_reg, g_tv, each, ids, clusters, margin=0.7):
    """Gradient w.r.t. `each` of  sum(final * g_final) + g_reg * loss_reg + g_tv * loss_tv  (what autograd
    computes through run_attention.py:852-883: the threshold passes the gradient unchanged, `same` is
    detached inside loss_tv)."""
    each = np.asarray(each, np.float64)
    g_final = np.asarray(g_final, np.float64).reshape(each.shape)
    batch, s, _ = each.shape
    k2 = gaussian_kernel2d().astype(np.float64)
    # adjoint of reflect-pad + correlation: correlate the zero-padded gradient with the flipped kernel on the
    # padded grid, then fold the 2-pixel borders back onto their mirror sources
    gp = np.zeros((batch, s + 4, s + 4))
    gz = np.pad(g_final, ((0, 0), (4, 4), (4, 4)))
    for dy in range(5):
        for dx in range(5):
            gp += k2[dy, dx] * gz[:, 4 - dy:4 - dy + s + 4, 4 - dx:4 - dx + s + 4]
    src = np.pad(np.arange(s), (2, 2), mode="reflect")
    g_same = np.zeros((batch, s, s))
    for u in range(s + 4):
        for v in range(s + 4):
            g_same[:, src[u], src[v]] += gp[:, u, v]
    g_each = np.zeros_like(each)
    n = each.size
    same = np.ones_like(each)
    for i in range(batch * clusters):
        sel = ids == i
        cnt = int(sel.sum())
        if cnt == 0:
            continue
        m = each[sel].mean()
        same[sel] = m
        g_mean = g_same[sel].sum() + (g_reg / batch if m > margin else 0.0)
        g_each[sel] += g_mean / cnt
    g_each += g_tv * 2.0 * (each - same) / n
    return g_each.astype(F32)
