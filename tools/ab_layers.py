"""A/B timing of the 1024^2 generator's modconv launches under w2e_modconv_tc2 knob settings, all in
ONE process on one GPU (box-to-box variance is several percent, so only same-run ratios count).
    python tools/ab_layers.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import where2edit_b200 as w2e  # noqa: E402
from where2edit_b200 import _native as N  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = "cuda:0"
    torch.manual_seed(0)
    gen = w2e.Generator(1024, 512, 8, channel_multiplier=2, precision="bf16").to(dev).eval()
    w = torch.randn(batch, gen.n_latent, 512, device=dev)
    from where2edit_b200 import engine as E
    gen._engine = E.SynthesisEngine(gen)
    eng = gen._engine
    # (name, ts_mode, flags, cluster_log2, blur variant): per-call switches, no library-global state
    settings = [("default", 1, 0, 0, "auto"), ("no per-class hand-over", 1, 256, 0, "auto"),
                ("last layer not on pixel pairs", 1, 0, 0, "auto")]
    if len(sys.argv) > 2:
        settings = [("default", 1, 0, 0, "auto")] + [(a, *[(v if v == "auto" else int(v)) for v in a.split(",")])
                                                     for a in sys.argv[2:]]
    # Round-robin over the settings INSIDE every repetition: the board heats up and power-caps during the run (a column that
    # is measured later reads ~8 % slower), so every setting is sampled under the same thermal conditions (who goes first rotates); median over reps.
    def apply(name, ts, flags, clus, blur):
        eng.tc2_cfg = N.tc2_config(ts_mode=ts, flags=flags, cluster_log2=clus)
        eng.blur_variant = blur
        eng.pair_mode = "not on pixel pairs" not in name
    acc = {s_[0]: {} for s_ in settings}
    with torch.no_grad():
        for s_ in settings:                      # warm-up of every variant (layout caches, kernel attributes)
            apply(*s_)
            for _ in range(2):
                gen([w], input_is_latent=True, randomize_noise=False)
        torch.cuda.synchronize()
        for rep in range(2 * len(settings)):
            for s_ in settings[rep % len(settings):] + settings[:rep % len(settings)]:   # rotate who goes first
                apply(*s_)
                N.STATS.trace = []
                gen([w], input_is_latent=True, randomize_noise=False)
                torch.cuda.synchronize()
                for call, note, e0, e1 in N.STATS.trace:
                    tag = (note or {}).get("tag") or call
                    acc[s_[0]].setdefault(tag, []).append(e0.elapsed_time(e1))
                N.STATS.trace = None
    results = {name: {k: sorted(v)[len(v) // 2] for k, v in a.items()} for name, a in acc.items()}   # median
    eng.tc2_cfg = None
    eng.pair_mode = True
    base = results["default"]
    names = [s_[0] for s_ in settings]
    print(f"{'launch':30s} " + " ".join(f"{n:>22s}" for n in names))
    tot = {n: 0.0 for n in names}
    for tag, t in base.items():
        if t < 0.03:
            continue
        row = f"{tag:30s} "
        for n in names:
            v = results[n].get(tag, float('nan'))
            tot[n] += v
            row += f"{v:22.3f} "
        print(row)
    print(f"{'sum':30s} " + " ".join(f"{tot[n]:22.3f}" for n in names))


if __name__ == "__main__":
    main()
