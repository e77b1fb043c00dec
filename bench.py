#!/usr/bin/env python
"""Benchmark of the Where2edit StyleGAN2 synthesis hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one `Generator.forward` over a batch of synthetic W+ latents (configs[1] of
BASELINE.json: FFHQ-1024 generator, batch 32 per GPU, bf16 tensor-core mode, fixed noise buffers).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.

Keys beyond the contract: `e2e_fp32_images` / `e2e_u8_images` (the end-to-end number with the reference's fp32 images and
with uint8 images quantised by the last layer's epilogue), `e2e_d2h_ceiling` (the same device-to-host copies with no
kernel running: what the host link allows), `gather_ok` (N > 1: the gathered batch checked slot by slot), `kernels`
(per-kind device time of a traced step), `pipelines` (edit pipeline of config 3 at batch 64, training step of config 4
eager and as one CUDA graph, tf32 forward, batch-1 latency), `next_rows` (the callers either side of the path).
`--workload train` is the contract line for config 4 (key `cuda_graph`: the same step as one graph launch).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "1024x1024 edited images/sec"
UNIT = "images/s"
# algorithmic work per image of the 1024^2 generator (SURVEY.md appendix A / BASELINE.md section 4)
MODCONV_GFLOP_PER_IMAGE = 148.13


def workload_config(size, batch, precision, world, gather="NCCL all-gather", workload="synthesis", image_dtype="bf16"):
    if workload == "train":
        what = (f"CLIP-loss latent optimisation step (BASELINE config 4): LevelsMapper edit w + 0.1*mapper(w) -> "
                f"StyleGAN2 FFHQ-{size} generator forward + backward to the mapper parameters through modconv / "
                f"upfirdn2d / fused_act (random init, channel_multiplier 2), batch {batch} per GPU, seeded synthetic "
                f"dL/dimage, {precision} mode")
        par = (f"batch sharded over {world} GPU(s)"
               + (", bucketed all-reduce of the shared mapper gradients overlapped with the backward (NCCL)" if world > 1 else ""))
    else:
        what = (f"StyleGAN2 FFHQ-{size} generator forward (random init, channel_multiplier 2), "
                f"batch {batch} per GPU from W+ latents, fixed noise buffers, {precision} mode")
        par = (f"batch sharded over {world} GPU(s)"
               + (f", per-step all-gather of the {image_dtype} images ({batch} per GPU per step: steady state of config 5's "
                  f"256 per GPU in micro-batches) overlapped on a side stream ({gather})" if world > 1 else ""))
    return {"workload": what, "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": par,
            "image_dtype": image_dtype if workload != "train" else "f32",
            "l2": "inputs larger than L2: every step streams multi-GB activations (no flush needed)"}


def ncu_traffic(kernel_substr):
    """dram bytes (read+write) per step of a kernel, from the committed `ncu --set full` capture of all
    its launches in one step (profiles/*_traffic.json, key '<round>_step_<kernel>'); None if absent."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json"))):
        with open(path) as fh:
            data = json.load(fh)
        for name, launches in data.items():
            if "_step_" in name and kernel_substr in name:
                best = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches)
    return best


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.25] or [l for (_, l) in self.lines]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_generator(size, precision, device):
    """Random-init generator of the reference architecture (no checkpoints exist offline); the
    zero-initialised noise weights / biases are perturbed so those code paths do real work."""
    import where2edit_b200 as w2e
    torch.manual_seed(0)
    gen = w2e.Generator(size, 512, 8, channel_multiplier=2, precision=precision)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in gen.named_parameters():
            if name.endswith("noise.weight") or name.endswith("activate.bias") or (name.endswith(".bias") and p.ndim == 4):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    return gen.to(device).eval()


def cpu_forward_images_per_s(state_dict, size, n_latent, batch, threads):
    """The reference's CPU algorithm (oracle port of models/stylegan2/model.py + op/) timed on the
    host cores: one warm-up image, then `batch` images."""
    from oracle import stylegan2_oracle as orc
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        orc.generator_forward_ref(state_dict, [torch.randn(1, n_latent, 512, generator=g)], size, input_is_latent=True)
        w = torch.randn(batch, n_latent, 512, generator=g)
        t0 = time.perf_counter()
        orc.generator_forward_ref(state_dict, [w], size, input_is_latent=True)
        dt = time.perf_counter() - t0
    return batch / dt, dt


def cpu_op_baselines(threads):
    """BASELINE.md section 5 (a, b) and config 1 on the host cores, through the oracle port of the reference's
    native ops (op/upfirdn2d.py:19-60, op/fused_act.py:23-39) and of its 256^2 generator: best of 3 each."""
    from oracle import stylegan2_oracle as orc
    from oracle import synth
    torch.set_num_threads(threads)

    def best(fn, reps=3):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return min(ts)

    out = {"cores": threads, "kind": "port"}
    with torch.no_grad():
        k4 = synth.blur_kernel_2d(gain=4.0)
        x = torch.randn(4, 128, 257, 257)
        dt = best(lambda: orc.upfirdn2d_native_ref(x, k4, 1, 1, 1, 1, 1, 1, 1, 1))
        byts = 4 * 128 * (257 * 257 + 256 * 256) * 4
        out["upfirdn2d_native_blur"] = {"shape": "[4,128,257,257] -> [4,128,256,256] fp32, pad (1,1)", "ms": dt * 1e3,
                                        "gbs_algorithmic": byts / dt / 1e9}
        xs = torch.randn(4, 3, 128, 128)
        dt = best(lambda: orc.upfirdn2d_native_ref(xs, k4, 2, 2, 1, 1, 2, 1, 2, 1))
        byts = 4 * 3 * (128 * 128 + 256 * 256) * 4
        out["upfirdn2d_native_up2"] = {"shape": "[4,3,128,128] -> [4,3,256,256] fp32, up 2, pad (2,1)", "ms": dt * 1e3,
                                       "gbs_algorithmic": byts / dt / 1e9}
        y = torch.randn(4, 128, 256, 256)
        b = torch.randn(128)
        dt = best(lambda: orc.fused_leaky_relu_ref(y, b))
        out["fused_leaky_relu"] = {"shape": "[4,128,256,256] fp32", "ms": dt * 1e3, "gbs_algorithmic": 2 * y.numel() * 4 / dt / 1e9}
        sd = synth.make_state_dict(256, seed=0, channel_multiplier=2)
        w = synth.make_wplus(4, 14, seed=2)
        dt = best(lambda: orc.generator_forward_ref(sd, [w], 256, input_is_latent=True))
        out["config1_generator256_batch4"] = {"ms": dt * 1e3, "images_per_s": 4 / dt}
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (it ships no native
    code and its Python cannot travel to the GPU box, so the pinned oracle port is timed) on all
    host cores; every step is a bounded sample of the workload (one 1024^2 image)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import stylegan2_oracle as orc
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    import math
    from oracle import synth
    sd = synth.make_state_dict(args.size, seed=0, channel_multiplier=2)   # nothing of the product package on this arm
    n_latent = 2 * int(math.log2(args.size)) - 2
    g = torch.Generator().manual_seed(2)
    sample_batch = 1
    with torch.no_grad():
        for _ in range(args.warmup):
            orc.generator_forward_ref(sd, [torch.randn(sample_batch, n_latent, 512, generator=g)], args.size,
                                      input_is_latent=True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            orc.generator_forward_ref(sd, [torch.randn(sample_batch, n_latent, 512, generator=g)], args.size,
                                      input_is_latent=True)
        dt = time.perf_counter() - t0
    value = sample_batch * args.steps / dt
    sample = (f"{sample_batch} image(s) of the {args.size}^2 generator per step (a bounded sample of the batch-{args.batch} "
              f"workload; the CPU path's images/s does not grow with the batch), fp32, torch CPU ops through the oracle "
              f"port of the reference (the reference is pure Python and does not travel to the GPU box), {threads} threads")
    cfg = workload_config(args.size, args.batch, args.precision, 1)
    cfg["sample_batch"] = sample_batch
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def summarise_trace(trace, peaks):
    """Per-kernel-kind totals from the traced step: algorithmic work / CUDA-event time."""
    kinds = {}
    layers = []
    for name, note, e0, e1 in trace:
        ms = e0.elapsed_time(e1)
        kind = (note or {}).get("kind", name)
        k = kinds.setdefault(kind, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        k["ms"] += ms
        k["launches"] += 1
        k["flops"] += (note or {}).get("flops", 0.0)
        k["bytes"] += (note or {}).get("bytes", 0.0)
        layers.append({"call": name, "tag": (note or {}).get("tag"), "ms": ms, "flops": (note or {}).get("flops"),
                       "bytes": (note or {}).get("bytes")})
    return kinds, layers


def gpu_ms(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def roofline_from_trace(trace, peaks, B, size, with_traffic):
    kinds, layers = summarise_trace(trace, peaks)
    traced_ms = sum(k["ms"] for k in kinds.values())
    roofline = roofline_up = None
    if "modconv" in kinds and kinds["modconv"]["ms"] > 0:
        k = kinds["modconv"]
        ach = k["flops"] / (k["ms"] * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "modconv_tc2_kernel (all 3x3 modulated-conv launches of a step)",
                    "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_tflops_sustained"],
                    "traffic": ncu_traffic("modconv") if with_traffic else None,
                    "peak_source": peaks["source"] + " bf16_tflops_sustained", "algorithmic_flops": k["flops"],
                    "launches": k["launches"], "share_of_step": k["ms"] / traced_ms}
    if "upfirdn2d" in kinds and kinds["upfirdn2d"]["ms"] > 0:
        k = kinds["upfirdn2d"]
        ach = k["bytes"] / (k["ms"] * 1e-3) / 1e9
        roofline_up = {"bound": "hbm", "kernel": "blur_act_nhwc_kernel (all blur launches of a step)",
                       "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                       "traffic": ncu_traffic("blur") if with_traffic else None,
                       "algorithmic_bytes": k["bytes"], "peak_source": peaks["source"] + " hbm_gbs", "launches": k["launches"],
                       "share_of_step": k["ms"] / traced_ms}
    return roofline, roofline_up, kinds, layers


def make_levels_mapper(dev):
    """LevelsMapper of the StyleCLIP-style edit (mapper/latent_mappers.py:47-82), seeded like the parity tests."""
    import types
    from where2edit_b200 import mappers
    torch.manual_seed(5)
    opts = types.SimpleNamespace(no_coarse_mapper=False, no_medium_mapper=False, no_fine_mapper=False)
    return mappers.LevelsMapper(opts).to(dev)


def extra_pipelines(gen, dev, size, n_lat):
    """Numbers the metric's name asks for but the headline workload does not contain (outside the timed region):
    cfg3 = mapper edit + original forward with feature capture + blended edited forward (edited images/s);
    cfg4 = forward + backward of the latent-optimisation step; the B=1 latency eager vs CUDA graph."""
    import where2edit_b200 as w2e
    out = {}
    mapper = make_levels_mapper(dev).eval()
    for p in mapper.parameters():
        p.requires_grad_(False)
    eb = 64   # BASELINE.json configs[2]: "batch 64 edit pipeline"
    g = torch.Generator().manual_seed(11)
    w = torch.randn(eb, n_lat, 512, generator=g).to(dev)
    mask = torch.rand(eb, 1, 64, 64, generator=g).to(dev)

    def edit():
        # the W+ flow of attention/run_attention.py:1196-1245: original forward with its features, mapper edit of the
        # latent, edited forward blended with the original's features through the attention mask
        with torch.no_grad():
            _, _, _, feats = gen([w], input_is_latent=True, randomize_noise=False, return_features=True)
            w_hat = w + 0.1 * mapper(w)
            img, _ = gen([w_hat], input_is_latent=True, randomize_noise=False, attention_layer=13,
                         attention_map=mask, feature_map=feats)
        return img

    def edit_only():   # the per-step part of the optimisation loop: features of the original are computed once (:1196-1203)
        with torch.no_grad():
            w_hat = w + 0.1 * mapper(w)
            img, _ = gen([w_hat], input_is_latent=True, randomize_noise=False, attention_layer=13,
                         attention_map=mask, feature_map=feats_once)
        return img
    def edit_one_map():   # the same pipeline bringing back only the feature maps the blend reads (capture_layers)
        with torch.no_grad():
            _, _, _, feats = gen([w], input_is_latent=True, randomize_noise=False, return_features=True,
                                 capture_layers=gen.blend_feature_layers(13))
            w_hat = w + 0.1 * mapper(w)
            img, _ = gen([w_hat], input_is_latent=True, randomize_noise=False, attention_layer=13,
                         attention_map=mask, feature_map=feats)
        return img
    try:
        ms = gpu_ms(edit, 3, warm=2)
        ms_one = gpu_ms(edit_one_map, 3, warm=2)
        with torch.no_grad():
            _, _, _, feats_once = gen([w], input_is_latent=True, randomize_noise=False, return_features=True)
        ms_edit = gpu_ms(edit_only, 3, warm=2)
        del feats_once
        out["edit_pipeline_cfg3"] = {
            "what": "original forward with all 26 features captured (fp32 NCHW, as the reference returns them) + "
                    "LevelsMapper edit of the W+ code + edited forward blended at layer 13 through a 64^2 mask "
                    "(run_attention.py:1196-1245), bf16 engine",
            "batch": eb, "ms": ms, "edited_images_per_s": eb / ms * 1e3,
            "capturing_only_the_blend_layer_ms": ms_one, "capturing_only_the_blend_layer_images_per_s": eb / ms_one * 1e3,
            "edited_forward_only_ms": ms_edit, "edited_forward_only_images_per_s": eb / ms_edit * 1e3}
    except Exception as exc:
        out["edit_pipeline_cfg3"] = {"error": repr(exc)[:300]}
    torch.cuda.empty_cache()
    # cfg4: forward + backward of the latent-optimisation step (python bench.py --workload train is the full contract line)
    try:
        tb = 16
        train_mapper = make_levels_mapper(dev).train()
        wt = torch.randn(tb, n_lat, 512, generator=g).to(dev)
        gimg = (torch.randn(tb, 3, size, size, generator=g) / (3 * size * size)).to(dev)
        frozen = [p.requires_grad for p in gen.parameters()]
        for p in gen.parameters():
            p.requires_grad_(False)

        def train_step():
            for p in train_mapper.parameters():
                p.grad = None
            img, _ = gen([wt + 0.1 * train_mapper(wt)], input_is_latent=True, randomize_noise=False)
            (img * gimg).sum().backward()
        ms = gpu_ms(train_step, 3, warm=2)
        with torch.no_grad():
            fwd = gpu_ms(lambda: gen([wt], input_is_latent=True, randomize_noise=False), 3, warm=1)
        graphed = None
        try:   # the same step as one CUDA-graph launch (where2edit_b200.GraphedStep)
            def graph_step(w_in):
                img, _ = gen([w_in + 0.1 * train_mapper(w_in)], input_is_latent=True, randomize_noise=False)
                (img * gimg).sum().backward()
            fast_step = w2e.GraphedStep(graph_step, [wt], params=train_mapper.parameters())
            graphed = gpu_ms(lambda: fast_step(wt), 3, warm=2)
            del fast_step
        except Exception as exc:
            graphed = repr(exc)[:200]
        for p, r in zip(gen.parameters(), frozen):
            p.requires_grad_(r)
        out["train_step_cfg4"] = {
            "what": "LevelsMapper edit -> 1024^2 forward + backward to the mapper parameters (channels-last bf16 engine), "
                    "seeded synthetic dL/dimage", "batch": tb, "ms": ms, "images_per_s": tb / ms * 1e3,
            "forward_only_ms": fwd, "ratio_to_forward": ms / fwd, "cuda_graph_ms": graphed}
        del train_mapper, wt, gimg
    except Exception as exc:
        out["train_step_cfg4"] = {"error": repr(exc)[:300]}
    torch.cuda.empty_cache()
    # tf32 mode (BASELINE config 2 names "bf16/tf32"): the fp32 module path with the 3x3 convolutions on tcgen05 kind::tf32
    try:
        tb32 = 8
        w32 = w[:tb32].contiguous()
        gen.set_precision("tf32")
        with torch.no_grad():
            ms32 = gpu_ms(lambda: gen([w32], input_is_latent=True, randomize_noise=False), 3, warm=2)
        out["tf32_forward"] = {"batch": tb32, "ms": ms32, "images_per_s": tb32 / ms32 * 1e3,
                               "note": "an accuracy mode (1.4e-3 of the reference at 1024^2 against 1.2e-2 in bf16): fp32 "
                                       "tensors, layout passes and fp32 elementwise kernels around every convolution"}
    except Exception as exc:
        out["tf32_forward"] = {"error": repr(exc)[:300]}
    finally:
        gen.set_precision("bf16")
    torch.cuda.empty_cache()
    # CUDA graph: B = 1 latency
    try:
        w1 = w[:1].contiguous()
        with torch.no_grad():
            eager = gpu_ms(lambda: gen([w1], input_is_latent=True, randomize_noise=False), 20, warm=3)
        fast = w2e.GraphedGenerator(gen, [w1], input_is_latent=True)
        graphed = gpu_ms(lambda: fast([w1]), 20, warm=3)
        out["latency_batch1"] = {"eager_ms": eager, "cuda_graph_ms": graphed}
        del fast
    except Exception as exc:
        out["latency_batch1"] = {"error": repr(exc)[:300]}
    torch.cuda.empty_cache()
    return out


def run_train(args, world, rank, dev, dist, peaks):
    """--workload train (BASELINE config 4): one step = LevelsMapper edit of a W+ batch -> generator forward ->
    seeded synthetic dL/dimage -> backward to the mapper parameters (+ at N > 1 the bucketed all-reduce of those
    gradients, overlapped with the backward).  Same contract line; `value` with the latents resident, `e2e` with the
    latents copied from pinned host memory and the loss read back every step."""
    from where2edit_b200 import _native as N
    from where2edit_b200 import parallel
    gen = make_generator(args.size, args.precision, dev)
    for p in gen.parameters():
        p.requires_grad_(False)
    mapper = make_levels_mapper(dev).train()
    reducer = parallel.GradBucketReducer(mapper.parameters())
    B, K_, W_ = args.batch, args.steps, args.warmup
    g = torch.Generator().manual_seed(2 + rank)
    host_w = [torch.randn(B, gen.n_latent, 512, generator=g).pin_memory() for _ in range(2)]
    dev_w = [w.to(dev) for w in host_w]
    gimg = (torch.randn(B, 3, args.size, args.size, generator=g) / (3 * args.size * args.size)).to(dev)
    host_loss = torch.zeros(1).pin_memory()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step(w):
        reducer.zero_grad()
        w_hat = w + 0.1 * mapper(w)
        img, _ = gen([w_hat], input_is_latent=True, randomize_noise=False)
        loss = (img * gimg).sum()
        loss.backward()
        reducer.finish()
        return loss

    for i in range(W_):
        step(dev_w[i % 2])
    barrier()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    N.STATS.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for i in range(K_):
        step(dev_w[i % 2])
    e1.record()
    barrier()
    t1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    launches = N.STATS.total()
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    gen.assert_ok()
    value = world * B * K_ / (total_ms / 1e3)

    def e2e_step(i):
        w = host_w[i % 2].to(dev, non_blocking=True)
        loss = step(w)
        host_loss.copy_(loss.detach().reshape(1), non_blocking=True)
    for i in range(2):
        e2e_step(i)
    barrier()
    e0.record()
    for i in range(K_):
        e2e_step(i)
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K_ / (float(ms2.item()) / 1e3)

    roofline = None
    kinds = {}
    fwd_ms = None
    if rank == 0:
        torch.cuda.synchronize()
        N.STATS.trace = []
        reducer.world, saved_world = 1, reducer.world    # rank-0 only: no collective in the traced step
        step(dev_w[0])
        reducer.world = saved_world
        torch.cuda.synchronize()
        trace, N.STATS.trace = N.STATS.trace, None
        roofline, _, kinds, layers = roofline_from_trace(trace, peaks, B, args.size, False)
        if roofline is not None:
            roofline["kernel"] = "modconv_tc2_kernel (forward and dgrad launches of one forward+backward step)"
        if args.layers_out:
            with open(args.layers_out, "w") as fh:
                json.dump({"batch": B, "size": args.size, "workload": "train", "launches": layers, "kinds": kinds}, fh, indent=1)
        with torch.no_grad():   # the forward alone (same precision, module path) for the fwd+bwd : fwd ratio
            fwd_ms = gpu_ms(lambda: gen([dev_w[0]], input_is_latent=True, randomize_noise=False), 3, warm=1)
    # the same step as ONE CUDA-graph launch (where2edit_b200.GraphedStep) on a second, hook-free copy of the mapper; at
    # N > 1 the refreshed gradients are all-reduced after the replay (one flat 12.6 MB buffer: ~0.1 ms, no overlap needed).
    # Every rank takes part (collectives), a rank whose capture failed makes all of them skip the measurement.
    graph_info = None
    if args.precision == "bf16" and os.environ.get("W2E_BENCH_GRAPH", "1") == "1":
        fast = err = None
        try:
            import where2edit_b200 as w2e
            mapper2 = make_levels_mapper(dev).train()

            def plain(w):
                img, _ = gen([w + 0.1 * mapper2(w)], input_is_latent=True, randomize_noise=False)
                loss = (img * gimg).sum()
                loss.backward()
                return loss
            fast = w2e.GraphedStep(plain, [dev_w[0]], params=mapper2.parameters())
        except Exception as exc:   # an auxiliary number must never cost the contract line
            err = repr(exc)[:300]
        ok_flag = torch.tensor([1 if fast is not None else 0], device=dev)
        if dist is not None:
            dist.all_reduce(ok_flag, op=dist.ReduceOp.MIN)
        if int(ok_flag.item()) == 1:
            params2 = list(mapper2.parameters())
            sizes = [p.numel() for p in params2]

            def graphed_step(w):
                loss = fast(w)
                if dist is not None:   # average the mapper gradients over the ranks (DDP semantics, run_attention.py:1022-1030)
                    grads = [p.grad for p in params2]
                    flat = torch.cat([g.reshape(-1) for g in grads])
                    dist.all_reduce(flat)
                    flat.div_(world)
                    torch._foreach_copy_(grads, [c.view_as(g) for c, g in zip(flat.split(sizes), grads)])
                return loss
            for i in range(3):
                graphed_step(dev_w[i % 2])
            barrier()
            e0.record()
            for i in range(K_):
                graphed_step(dev_w[i % 2])
            e1.record()
            barrier()
            gt = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(gt, op=dist.ReduceOp.MAX)
            gms = float(gt.item()) / K_
            fast(dev_w[0])
            torch.cuda.synchronize()
            got = [p.grad.detach().clone() for p in params2]
            for p in params2:
                p.grad = None
            plain(dev_w[0])
            torch.cuda.synchronize()
            same = all(torch.equal(a, p.grad) for a, p in zip(got, params2))
            graph_info = {"ms_per_step": gms, "images_per_s": world * B / gms * 1e3,
                          "gradients_bit_identical_to_eager": bool(same),
                          "what": "the same step (mapper forward, generator forward, loss, backward) captured once and "
                                  "replayed as one CUDA-graph launch per step, latents copied into the graph's input buffer"
                                  + ("; the mapper gradients all-reduced (one flat buffer) after every replay" if dist is not None else "")}
        else:
            graph_info = {"error": err or "capture failed on another rank"}
        del fast
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K_, "warmup": W_,
            "ms_per_step": total_ms / K_, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args.size, B, args.precision, world, workload="train"),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": host_w[0].numel() * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "roofline": roofline,
            "kernels": {k: {"ms": round(v["ms"], 3), "launches": v["launches"]} for k, v in kinds.items()},
            "forward_only_ms_same_path": fwd_ms,
            "cuda_graph": graph_info,
            "mapper_gradient_floats": int(reducer.flat.numel()), "gradient_buckets": len(reducer.buckets),
            "cpu_baseline": None,
        }))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="synthesis", choices=["synthesis", "train"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (32; 16 for --workload train)")
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--image-dtype", default="bf16", choices=["bf16", "fp32", "u8"],
                    help="dtype of the images the last layer writes / the caller receives (u8: the save_image quantisation)")
    ap.add_argument("--cpu-sample", type=int, default=12, help="images of the CPU-baseline sample (0 = skip)")
    ap.add_argument("--layers-out", default=None, help="write the per-launch trace of one step to this JSON file")
    ap.add_argument("--verify", action="store_true", help="N > 1: check the gathered images slot by slot (default on)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.batch is None:
        args.batch = 16 if args.workload == "train" else 32

    if args.impl == "reference":
        run_reference(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU implementation")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))

    peaks = measured_peaks()
    if args.workload == "train":
        run_train(args, world, rank, dev, dist, peaks)
        return

    from where2edit_b200 import _native as N
    from where2edit_b200 import parallel
    gen = make_generator(args.size, args.precision, dev)
    img_dtype = torch.float32
    if args.precision == "bf16" and args.image_dtype != "fp32":
        img_dtype = torch.bfloat16 if args.image_dtype == "bf16" else torch.uint8
    img_name = {torch.bfloat16: "bf16", torch.float32: "f32", torch.uint8: "u8"}[img_dtype]
    gen.set_image_output(img_dtype)
    # Multi-GPU: the persistent convolution kernels normally occupy every SM, so the NCCL all-gather kernels of
    # the previous step could only run in the gaps between them; leaving a few SMs free lets the collective
    # overlap the kernels (W2E_RESERVE_SMS; default 0: measured at 4 GPUs 14 649 / 14 362 / 13 335 images/s with 0 / 8 / 16 SMs reserved).
    reserve = int(os.environ.get("W2E_RESERVE_SMS", "0"))
    if reserve > 0 and args.precision == "bf16":
        from where2edit_b200 import engine as _engine
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        gen._engine = _engine.SynthesisEngine(gen)
        gen._engine.tc2_cfg = N.tc2_config(max_ctas=max(2, sms - reserve))   # per-engine, per-call: no library global
    B, K_, W_ = args.batch, args.steps, args.warmup

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    g = torch.Generator().manual_seed(2 + rank)
    n_lat = gen.n_latent
    host_w = [torch.randn(B, n_lat, 512, generator=g).pin_memory() for _ in range(2)]
    dev_w = [w.to(dev) for w in host_w]
    comm_stream = torch.cuda.Stream() if dist is not None else None
    gathered = peer = None
    gather_kind = "none"
    SLOTS = 2
    if dist is not None:
        # default: copy-engine pushes into peer-mapped buffers (no SM use), each step closed by two device-side
        # cross-rank barriers so that every step's gathered batch is consumable; W2E_GATHER=nccl: the NCCL all-gather
        if os.environ.get("W2E_GATHER", "p2p") == "p2p":
            try:
                peer = parallel.PeerGather((B, 3, args.size, args.size), img_dtype, dev, slots=SLOTS)
                gather_kind = "copy-engine peer pushes into symmetric memory, barrier before and after every step's pushes"
            except Exception as exc:   # no symmetric memory in this build / topology: the NCCL collective
                if rank == 0:
                    print(f"bench.py: PeerGather unavailable ({exc!r}); using NCCL", file=sys.stderr)
        if peer is None:
            gathered = [torch.empty((world * B, 3, args.size, args.size), device=dev, dtype=img_dtype)
                        for _ in range(SLOTS)]
            gather_kind = "NCCL all-gather"
    pushed = [None] * SLOTS   # event on the side stream after the pushes of the step that last used a slot

    def step(i, w, gather=True):
        """One pass of the hot path; at N>1 the images are all-gathered on a side stream so the collective of
        step i overlaps the kernels of step i+1.  The last layer writes its image straight into this rank's shard of
        the gather buffer, in the gathered dtype (no conversion pass, no self-copy)."""
        slot = i % SLOTS
        if peer is not None and gather:
            # the compute stream must not overwrite the shard while step i-SLOTS is still being pushed from it
            if pushed[slot] is not None:
                torch.cuda.current_stream().wait_event(pushed[slot])
            gen.set_image_output(img_dtype, peer.own(slot))
        else:
            gen.set_image_output(img_dtype)
        with torch.no_grad():
            img, _ = gen([w], input_is_latent=True, randomize_noise=False)
        if dist is not None and gather:
            comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm_stream):
                if peer is not None:
                    peer.push(None, slot)
                    pushed[slot] = torch.cuda.Event()
                    pushed[slot].record(comm_stream)
                else:
                    parallel.gather_images(img, out=gathered[slot])
            img.record_stream(comm_stream)
        return img

    # ---------------- device-resident throughput (`value`)
    for i in range(W_):
        step(i, dev_w[i % 2])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    N.STATS.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for i in range(K_):
        step(i, dev_w[i % 2])
    if comm_stream is not None:
        torch.cuda.current_stream().wait_stream(comm_stream)
    e1.record()
    barrier()
    t1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    launches = N.STATS.total()
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    gen.assert_ok()
    value = world * B * K_ / (total_ms / 1e3)

    # ---------------- N > 1: the gathered batch of one more step, checked slot by slot against local recomputation
    gather_ok = None
    if dist is not None:
        step(0, dev_w[0])
        torch.cuda.current_stream().wait_stream(comm_stream)
        torch.cuda.synchronize()
        full = peer.result(0) if peer is not None else gathered[0]
        ok = True
        gen.set_image_output(img_dtype)
        for q in range(world):   # every kernel is deterministic and batch-invariant: rank q's images can be recomputed here
            gq = torch.Generator().manual_seed(2 + q)
            wq = torch.randn(B, n_lat, 512, generator=gq).to(dev)
            with torch.no_grad():
                want, _ = gen([wq], input_is_latent=True, randomize_noise=False)
            ok = ok and bool(torch.equal(full[q * B:(q + 1) * B], want))
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_ok = bool(flag.item())
        barrier()

    # ---------------- end to end through the public API with host buffers (`e2e`)
    copy_stream = torch.cuda.Stream()

    def measure_e2e(dtype, steps):
        host_img = [torch.empty((B, 3, args.size, args.size), dtype=dtype).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        keep_dtype = dtype

        def e2e_step(i):
            done[i % 2].synchronize()                               # host buffer (and gather shard) i%2 free again
            w = host_w[i % 2].to(dev, non_blocking=True)            # H2D of this step's latents (pinned)
            if dist is None or peer is None:
                gen.set_image_output(keep_dtype)
                with torch.no_grad():
                    img, _ = gen([w], input_is_latent=True, randomize_noise=False)
            else:
                img = step(i, w)
            copy_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(copy_stream):
                host_img[i % 2].copy_(img, non_blocking=True)       # D2H of the step's images
                done[i % 2].record(copy_stream)
            img.record_stream(copy_stream)

        for i in range(2):
            e2e_step(i)
        barrier()
        e0.record()
        for i in range(steps):
            e2e_step(i)
        torch.cuda.current_stream().wait_stream(copy_stream)
        if comm_stream is not None:
            torch.cuda.current_stream().wait_stream(comm_stream)
        e1.record()
        barrier()
        ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        return world * B * steps / (float(ms2.item()) / 1e3), host_img[0].numel() * host_img[0].element_size()

    e2e_value, d2h = measure_e2e(img_dtype, K_)
    h2d = host_w[0].numel() * 4
    e2e_f32 = e2e_u8 = None
    if dist is None and img_dtype != torch.float32:
        v32, d32 = measure_e2e(torch.float32, max(5, K_ // 3))
        e2e_f32 = {"value": v32, "unit": UNIT, "d2h_bytes_per_step": d32,
                   "note": "same, with fp32 images (the reference's output dtype) copied out: PCIe-bound"}
    if img_dtype != torch.uint8 and args.precision == "bf16" and (dist is None or peer is None):
        v8, d8 = measure_e2e(torch.uint8, max(5, K_ // 3))
        e2e_u8 = {"value": v8, "unit": UNIT, "d2h_bytes_per_step": d8,
                  "note": "same, with uint8 images (the save_image quantisation the reference applies to its results, "
                          "done by the last layer's epilogue) copied out"}
    gen.set_image_output(img_dtype)

    # ---------------- the host side's ceiling for `e2e`: the same device-to-host copies with NO kernels running, all ranks
    # at once (pinned buffers of one step's images, as in the e2e loop).  e2e cannot exceed images / this time.
    def d2h_ceiling(dtype, steps):
        src = torch.zeros((B, 3, args.size, args.size), device=dev, dtype=dtype)
        dst = [torch.empty(src.shape, dtype=dtype).pin_memory() for _ in range(2)]
        for i in range(2):
            dst[i].copy_(src, non_blocking=True)
        barrier()
        e0.record()
        for i in range(steps):
            dst[i % 2].copy_(src, non_blocking=True)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item()) / 1e3
        nbytes = src.numel() * src.element_size()
        return {"aggregate_gb_per_s": world * nbytes * steps / sec / 1e9, "images_per_s_ceiling": world * B * steps / sec,
                "image_dtype": {torch.bfloat16: "bf16", torch.float32: "f32", torch.uint8: "u8"}[dtype]}

    host_ceiling = d2h_ceiling(img_dtype, max(5, K_ // 3))

    # ---------------- per-kernel roofline from one traced step (CUDA events around every launch)
    roofline = roofline_up = None
    kinds = {}
    TRACED = 5
    if rank == 0:
        step(0, dev_w[0], gather=False)   # rank-0 only: must not enter a collective
        torch.cuda.synchronize()
        N.STATS.trace = []
        for i in range(TRACED):           # the AVERAGE launch duration over several steps, not one sample
            step(i, dev_w[i % 2], gather=False)
        torch.cuda.synchronize()
        trace, N.STATS.trace = N.STATS.trace, None
        roofline, roofline_up, kinds, layers = roofline_from_trace(trace, peaks, B, args.size, B == 32 and args.size == 1024)
        for k in kinds.values():          # per step
            k["ms"] /= TRACED
            k["launches"] //= TRACED
            k["flops"] /= TRACED
            k["bytes"] /= TRACED
        for r in (roofline, roofline_up):
            if r is not None:
                r["launches"] //= TRACED
                for key in ("algorithmic_flops", "algorithmic_bytes"):
                    if key in r:
                        r[key] /= TRACED
                r["traced_steps"] = TRACED
        if args.layers_out:
            with open(args.layers_out, "w") as fh:
                json.dump({"batch": B, "size": args.size, "precision": args.precision, "launches": layers,
                           "kinds": kinds}, fh, indent=1)

    # ---------------- CPU baseline (rank 0, N == 1 only): the reference algorithm on the host cores
    cpu = None
    if rank == 0 and world == 1 and args.cpu_sample > 0:
        threads = os.cpu_count() or 1
        sd = {k: v.detach().cpu().float() for k, v in gen.state_dict().items()}
        v, dt = cpu_forward_images_per_s(sd, args.size, n_lat, args.cpu_sample, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.cpu_sample} images of the same {args.size}^2 generator forward in one batch "
                         f"({dt:.1f} s), fp32 torch CPU ops via the oracle port of the reference (pure Python: it cannot "
                         f"travel to the GPU box), after a 1-image warm-up",
               "ops": cpu_op_baselines(threads)}

    # ---------------- configs 3 / 4 and the callers either side of the path, outside the timed region
    extras = next_rows = None
    if rank == 0 and world == 1 and args.precision == "bf16" and args.size == 1024 \
            and os.environ.get("W2E_BENCH_EXTRAS", "1") == "1":
        gen.set_image_output(torch.float32)
        try:
            extras = extra_pipelines(gen, dev, args.size, n_lat)
        except Exception as exc:
            extras = {"error": repr(exc)[:300]}
        gen.set_image_output(img_dtype)
    if rank == 0 and world == 1 and args.precision == "bf16" and os.environ.get("W2E_BENCH_NEXT_ROWS", "1") == "1":
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("next_rows_bench", os.path.join(ROOT, "tools", "next_rows_bench.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            torch.cuda.empty_cache()
            next_rows = mod.measure(iters=10)
        except Exception as exc:   # auxiliary numbers must never cost the headline line
            next_rows = {"error": repr(exc)[:300]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K_, "warmup": W_,
            "ms_per_step": total_ms / K_, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args.size, B, args.precision, world, gather_kind, image_dtype=img_name),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "image_dtype": img_name},
            "e2e_fp32_images": e2e_f32,
            "e2e_u8_images": e2e_u8,
            "e2e_d2h_ceiling": host_ceiling,
            "gpu_launches": launches,
            "gather_ok": gather_ok,
            "roofline": roofline, "roofline_upfirdn2d": roofline_up,
            "modconv_tflops_per_step_algorithmic": MODCONV_GFLOP_PER_IMAGE * B / 1e3 if args.size == 1024 else None,
            "kernels": {k: {"ms": round(v["ms"], 3), "launches": v["launches"]} for k, v in kinds.items()},
            "cpu_baseline": cpu,
            "pipelines": extras,
            "next_rows": next_rows,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
