"""Timing of the SURVEY.md section 8f rows built so far (region-mask construction, CLIP pre-resample) on one
B200, CUDA events on the current stream, with the reference's formulation (plain torch on the same GPU) and the
CPU oracle beside them.  Prints one JSON object; the committed copy is profiles/r01_next_rows.json.

    python tools/next_rows_bench.py [--iters 20]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from where2edit_b200 import region, resample  # noqa: E402


def gpu_ms(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured hbm_gbs"
    except Exception:
        return 6650.0, "fallback"


def measure(iters=20):
    """All timings as one dict (also called by bench.py, outside its timed region)."""
    args = argparse.Namespace(iters=iters)
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(0)
    peak, peak_src = hbm_peak()
    out = {"hbm_peak_gbs": peak, "peak_source": peak_src}

    # ---- region mask at the cfg3 shape: B=64, 512-channel 64x64 feature, 20 clusters, 64x64 attention
    b, c, h, k = 64, 512, 64, 20
    feat = torch.randn(b, c, h, h, device=dev)
    centres = torch.randn(k, c + 2 * (c // 16), device=dev) * 0.2
    each = torch.rand(b, h, h, device=dev) * 0.5 + 0.5
    ids = region.assign_clusters(feat, centres, h, k)
    ms = gpu_ms(lambda: region.assign_clusters(feat, centres, h, k), args.iters)
    byts = feat.numel() * 4 + ids.numel() * 8
    out["cluster_assign"] = {"shape": f"B={b} C={c} h={h} K={k}", "ms": ms, "algorithmic_bytes": byts,
                             "gbs": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / peak,
                             "distance_terms_per_s": b * h * h * k * (c + 2 * (c // 16)) / ms * 1e3}
    out["region_attention_fwd"] = {"ms": gpu_ms(lambda: region.region_attention(each, ids, k), args.iters)}
    e = each.clone().requires_grad_(True)
    head = torch.randn(b, 1, h, h, device=dev)

    def fwd_bwd():
        e.grad = None
        f, _, r, tv = region.region_attention(e, ids, k)
        ((f * head).sum() + r.sum() + tv).backward()
    out["region_attention_fwd_bwd"] = {"ms": gpu_ms(fwd_bwd, args.iters), "note": "through autograd (host-bound)"}
    # the backward kernel alone, through the C ABI
    from where2edit_b200 import _native as N
    f_, same_, _, _ = region.region_attention(each, ids, k)
    stats_ = torch.empty(b, k, 2, device=dev)
    parts_ = torch.empty(b, 2, device=dev)
    losses_ = torch.empty(2, device=dev)
    lib = N.load()
    taps = region._taps()
    lib.w2e_region_mask_fwd(N.ptr(each), N.ptr(ids), taps, N.ptr(f_), N.ptr(same_), N.ptr(stats_), N.ptr(parts_),
                            N.ptr(losses_), b, h, k, 0.8, 0.7, N.stream_ptr())
    g_each = torch.empty_like(each)
    g_l = torch.ones(2, device=dev)
    out["region_mask_bwd_kernel"] = {"ms": gpu_ms(lambda: lib.w2e_region_mask_bwd(
        N.ptr(head), N.ptr(g_l), N.ptr(each), N.ptr(ids), N.ptr(same_), N.ptr(stats_), taps, N.ptr(g_each), b, h, k, 0.7,
        N.stream_ptr()), args.iters)}

    # the reference's formulation on the same GPU (utils.py:244-263, run_attention.py:775-794, 852-884)
    import torchvision

    def ref_assign():
        pc = c // 16
        xp = torch.arange(h, device=dev).float().unsqueeze(0).repeat(h, 1) * 2 / float(h - 1) - 1
        yp = torch.arange(h, device=dev).float().unsqueeze(1).repeat(1, h) * 2 / float(h - 1) - 1
        cat = torch.cat([feat, xp[None, None].repeat(b, pc, 1, 1), yp[None, None].repeat(b, pc, 1, 1)], 1)
        cat = cat.permute(0, 2, 3, 1).contiguous().view(-1, cat.shape[1])
        dis = ((cat.unsqueeze(1) - centres.unsqueeze(0)) ** 2.0).sum(-1)
        return torch.arange(b, device=dev).view(b, 1, 1) * k + dis.argmin(1).view(b, h, h)

    def ref_region(ids_):
        same = torch.ones((b, h, h), device=dev)
        reg = torch.zeros(1, device=dev)
        for i in range(b * k):
            m = torch.mean(each[ids_ == i])
            same[ids_ == i] = m
            if not torch.isnan(m):
                reg += torch.relu(m - 0.7)
        a = same.unsqueeze(1)
        fin = a.clone()
        fin[a < 0.8] = a[a < 0.8] - a[a < 0.8].detach()
        return torchvision.transforms.functional.gaussian_blur(fin, 5), reg / b

    try:
        ids_ref = ref_assign()
        out["cluster_assign"]["torch_formulation_ms"] = gpu_ms(ref_assign, 3, warmup=1)
        out["cluster_assign"]["ids_equal_to_torch_formulation"] = float((ids_ref == ids).float().mean())
        fin_ref, _ = ref_region(ids)
        fin, _, _, _ = region.region_attention(each, ids, k)
        out["region_attention_fwd"]["max_abs_diff_vs_torch_formulation"] = float((fin - fin_ref).abs().max())
        t0 = time.time()
        ref_region(ids)
        torch.cuda.synchronize()
        out["region_attention_fwd"]["torch_formulation_ms"] = (time.time() - t0) * 1e3
    except Exception as exc:  # memory: the broadcast difference tensor is 12 GB at this shape
        out["torch_formulation_error"] = repr(exc)[:200]
    del feat

    # ---- CLIP pre-resample: B=32 images of 1024^2 -> 224^2
    bb = 32
    img = torch.randn(bb, 3, 1024, 1024, device=dev)
    y = resample.clip_resample(img, 7, 32)
    ms = gpu_ms(lambda: resample.clip_resample(img, 7, 32), args.iters)
    byts = (img.numel() + y.numel()) * 4
    out["clip_resample_fwd"] = {"shape": "B=32 [3,1024,1024] -> [3,224,224]", "ms": ms, "algorithmic_bytes": byts,
                                "gbs": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / peak}
    gy = torch.randn_like(y)
    xg = img.requires_grad_(True)

    def rs_bwd():
        xg.grad = None
        resample.clip_resample(xg, 7, 32).backward(gy)
    ms2 = gpu_ms(rs_bwd, args.iters)
    out["clip_resample_fwd_bwd"] = {"ms": ms2, "algorithmic_bytes": 2 * byts, "gbs": 2 * byts / ms2 / 1e6,
                                    "frac_hbm": 2 * byts / ms2 / 1e6 / peak}
    gx = torch.empty(bb, 3, 1024, 1024, device=dev)
    ms3 = gpu_ms(lambda: lib.w2e_box_resample_bwd(N.ptr(gy), N.ptr(gx), bb * 3, 1024, 1024, 7, 32, N.stream_ptr()), args.iters)
    out["clip_resample_bwd_kernel"] = {"ms": ms3, "algorithmic_bytes": byts, "gbs": byts / ms3 / 1e6,
                                       "frac_hbm": byts / ms3 / 1e6 / peak}
    out["clip_resample_fwd_bwd"]["note"] = "through autograd (host-bound)"
    small = img[:4].detach()
    f = torch.nn.functional
    out["clip_resample_fwd"]["torch_formulation_ms_per_32"] = 8 * gpu_ms(
        lambda: f.avg_pool2d(f.interpolate(small, scale_factor=7), 32), 3, warmup=1)

    # ---- CPU oracle, bounded samples
    from oracle import region_oracle as ro
    from oracle import resample_oracle as rso
    e_np, i_np = each[:4].cpu().numpy(), (ids[:4]).cpu().numpy()
    t0 = time.time()
    ro.region_attention(e_np, i_np, k)
    # ---- cluster-style mapper (run_attention.py:703-893) over the features of a 1024^2 generator: attention heads on the
    # pixels that survive the nearest resize to 64^2 against heads on the full-resolution maps (the reference's order)
    try:
        import where2edit_b200 as w2e
        from where2edit_b200 import mappers
        mb = 4
        gen = w2e.Generator(1024, 512, 8, channel_multiplier=2, precision="bf16").to(dev).eval()
        with torch.no_grad():
            _, _, styles, feats = gen([torch.randn(mb, gen.n_latent, 512, device=dev)], input_is_latent=True,
                                      randomize_noise=False, return_features=True)
            feats = list(feats) + [gen.input.input.repeat(mb, 1, 1, 1)]
        cl = next(i for i, f in enumerate(feats) if f.shape[1] >= 256 and f.shape[-1] == 64) + 1   # a 64^2 convolution feature
        cdim = feats[cl - 1].shape[1] + 2 * (feats[cl - 1].shape[1] // 16)
        cm = mappers.ClusterStyleMapper(gen.n_latent, 1024, 512, attention_layer=11, cluster_layer=cl, channel_multiplier=2,
                                        clusters=10, cluster_dim=cdim).to(dev).eval()
        text = torch.randn(mb, 512, device=dev)
        xin = [torch.cat([text.unsqueeze(1), s_[:, :, :, 0, 0]], dim=-1) for s_ in styles]
        res = {}
        with torch.no_grad():
            for fused, sample in ((True, True), (False, True), (False, False)):
                cm.fused_heads, cm.sample_first = fused, sample
                res[(fused, sample)] = gpu_ms(lambda: cm(xin, feats, 64), 3, warmup=1)
        cm.fused_heads, cm.sample_first = True, True
        graphed_ms = None
        try:   # the same forward as one CUDA-graph launch (no host synchronisation anywhere on the path)
            with torch.no_grad():
                fast = w2e.GraphedStep(lambda xs_, fs_: cm(xs_, fs_, 64), [xin, feats])
                graphed_ms = gpu_ms(lambda: fast.graph.replay(), 5, warmup=2)
            del fast
        except Exception as exc:
            graphed_ms = repr(exc)[:200]
        out["cluster_style_mapper_fwd"] = {"shape": f"B={mb}, 26 feature maps of the 1024^2 generator, 64^2 attention map",
                                           "ms": res[(True, True)], "cuda_graph_ms": graphed_ms,
                                           "per_head_modules_on_sampled_pixels_ms": res[(False, True)],
                                           "per_head_modules_on_full_resolution_ms": res[(False, False)],
                                           "note": "grouped launch of all attention heads (w2e_attn_heads_fwd) vs one "
                                                   "StyledConv module call per head"}
        del gen, cm, feats, styles
        torch.cuda.empty_cache()
    except Exception as exc:
        out["cluster_style_mapper_fwd"] = {"error": repr(exc)[:200]}
    out["region_attention_fwd"]["cpu_oracle_ms_per_64"] = (time.time() - t0) * 1e3 * 16
    x1 = small[:1].cpu().numpy()
    t0 = time.time()
    rso.clip_resample(x1, 7, 32)
    out["clip_resample_fwd"]["cpu_oracle_ms_per_32"] = (time.time() - t0) * 1e3 * 32
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    print(json.dumps(measure(ap.parse_args().iters)))


if __name__ == "__main__":
    main()
