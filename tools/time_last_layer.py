"""Time the RGB-only 32 -> 32 channel last layer (w2e_modconv_tc2_rgb[_pair]) alone: batch 32, 1024^2, CUDA events around
each of 30 launches, min / median.  W2E_LIB_PATH selects another build of libw2e.so (kernel A/B of code versions in one
gpurun call), W2E_PAIR=0 the pixel kernel.
    python tools/time_last_layer.py [batch]"""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import where2edit_b200 as w2e  # noqa: E402
from where2edit_b200 import _native as N, engine as E, functional as K  # noqa: E402


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = "cuda:0"
    torch.manual_seed(0)
    gen = w2e.Generator(8, 512, 1).to(dev)
    eng = E.SynthesisEngine(gen)
    for c, h, last in ((32, 1024, True), (64, 512, False)):
        time_layer(eng, b, c, h, last, dev)


def time_layer(eng, b, c, h, last, dev):
    m = w2e.StyledConv(c, c, 3, 16).to(dev)
    rgbm = w2e.ToRGB(c, 16).to(dev)
    xs = torch.randn(b, h, h, c, device=dev).to(torch.bfloat16)
    s = (1 + 0.3 * torch.randn(b, c, device=dev)).contiguous()
    pw = eng._tc_weight(m.conv)
    d = K.demod_coefficients(s, pw.wsq)
    noise = torch.randn(1, 1, h, h, device=dev)
    srgb = (1 + 0.3 * torch.randn(b, c, device=dev)).contiguous()
    nxt = (1 + 0.3 * torch.randn(b, c, device=dev)).contiguous()
    skip = torch.randn(b, 3, h // 2, h // 2, device=dev)
    out = torch.empty(b, 3, h, h, device=dev, dtype=torch.bfloat16 if last else torch.float32)

    def run():   # the last layer writes only the image; the others the modulated activation for the next layer + the image
        return eng._conv2_rgb(xs, pw, d, noise, m.noise.weight.detach(), m.activate.bias.detach(), None if last else nxt,
                              False, not last, rgbm, srgb, skip, out.dtype, out)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    eng.assert_ok()
    print(f"lib {os.environ.get('W2E_LIB_PATH', 'default')} pair {os.environ.get('W2E_PAIR', '1')} {c}->{c}@{h}: "
          f"min {min(ts):.3f} ms, median {statistics.median(ts):.3f} ms")


if __name__ == "__main__":
    main()
