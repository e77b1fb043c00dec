"""Where does the difference between a step and the sum of its kernels go?  Batch-32 1024^2 forward, eager launches
against ONE CUDA-graph launch per step (no host work between the kernels), 30 steps each, CUDA events around the loop;
the sum of the per-launch event durations of traced steps is printed next to them.
    python tools/step_gap.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import where2edit_b200 as w2e  # noqa: E402
from where2edit_b200 import _native as N  # noqa: E402


def loop_ms(fn, steps=30, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = "cuda:0"
    torch.manual_seed(0)
    gen = w2e.Generator(1024, 512, 8, channel_multiplier=2, precision="bf16").to(dev).eval()
    gen.set_image_output(torch.bfloat16)
    w = torch.randn(batch, gen.n_latent, 512, device=dev)
    with torch.no_grad():
        eager = loop_ms(lambda: gen([w], input_is_latent=True, randomize_noise=False))
        fast = w2e.GraphedGenerator(gen, [w], input_is_latent=True)
        graphed = loop_ms(lambda: fast.graph.replay())
        eager2 = loop_ms(lambda: gen([w], input_is_latent=True, randomize_noise=False))
        N.STATS.trace = []
        for _ in range(5):
            gen([w], input_is_latent=True, randomize_noise=False)
        torch.cuda.synchronize()
        trace, N.STATS.trace = N.STATS.trace, None
    ksum = sum(e0.elapsed_time(e1) for _, _, e0, e1 in trace) / 5
    print(f"batch {batch}: eager {eager:.3f} ms/step, CUDA graph {graphed:.3f}, eager again {eager2:.3f}, "
          f"sum of per-launch event durations (traced steps) {ksum:.3f}, launches per step {len(trace) // 5}")


if __name__ == "__main__":
    main()
