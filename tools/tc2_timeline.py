"""Timeline of CTA 0 of one w2e_modconv_tc2[_rgb] launch (clock64 stamps written by the kernel's
three roles, see include/w2e.h: w2e_tc2_config.timeline).  Prints per tile, relative to the first
stamp: producer {inputs slot free, A issued}, MMA {acc free, A landed, issued}, epilogue {inputs,
acc ready, done}; with tr=2 (fused up-conv + blur) the P columns hold {pre-blur tile written, FIR of chunk 0 done}.
    python tools/tc2_timeline.py [cin cout h batch rgb(0/1) tr(0/1/2)]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import where2edit_b200 as w2e  # noqa: E402
from where2edit_b200 import _native as N, engine as E, functional as K  # noqa: E402


def main():
    a = [int(v) for v in sys.argv[1:]] + [None] * 6
    cin, cout, h, b, rgb, tr = (a[0] or 32, a[1] or 32, a[2] or 1024, a[3] or 8, 1 if a[4] is None else a[4], a[5] or 0)
    dev = "cuda:0"
    torch.manual_seed(0)
    gen = w2e.Generator(8, 512, 1).to(dev)
    eng = E.SynthesisEngine(gen)
    m = w2e.StyledConv(cin, cout, 3, 16, upsample=bool(tr)).to(dev)
    x = torch.randn(b, cin, h, h, device=dev)
    s = (1 + 0.3 * torch.randn(b, cin, device=dev)).contiguous()
    pw = eng._tc_weight(m.conv)
    d = K.demod_coefficients(s, pw.wsq)
    xs = eng._to_nhwc(x, s, b)
    del x
    buf = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
    noise = torch.randn(1, 1, h, h, device=dev)
    nxt = (1 + 0.3 * torch.randn(b, cout, device=dev)).contiguous()

    def run():
        if tr == 2:   # fused up-convolution + blur
            return eng._upblur(xs, pw, d, m.conv.blur.kernel, m.conv.blur.pad, m.activate.bias.detach(),
                               torch.randn(1, 1, 2 * h, 2 * h, device=dev), m.noise.weight.detach(), nxt, False, True)
        if tr:
            return eng._conv2(xs, pw, d, None, None, None, None, True, False, True, N.ACT_NONE)
        if rgb:
            rgbm = w2e.ToRGB(cout, 16).to(dev)
            skip = torch.randn(b, 3, h // 2, h // 2, device=dev)
            return eng._conv2_rgb(xs, pw, d, noise, m.noise.weight.detach(), m.activate.bias.detach(), None, False, False,
                                  rgbm, nxt, skip)
        return eng._conv2(xs, pw, d, noise, m.noise.weight.detach(), m.activate.bias.detach(), nxt, False, True, False,
                          N.ACT_LRELU)

    run()
    torch.cuda.synchronize()
    eng.tc2_cfg = N.tc2_config(timeline=buf, flags=int(os.environ.get("W2E_TC2_FLAGS", "0")))
    run()
    torch.cuda.synchronize()
    eng.tc2_cfg = None
    t = buf.cpu().reshape(64, 8)
    t0 = int(t[t > 0].min())
    print("tile |  P:slot  P:Aiss | M:accfree M:Aland M:issued | E:inputs E:accrdy E:done   (cycles since start)")
    for i in range(40):
        r = [int(v) - t0 if v > 0 else -1 for v in t[i]]
        print(f"{i:4d} | {r[0]:7d} {r[1]:7d} | {r[2]:8d} {r[3]:7d} {r[4]:8d} | {r[5]:8d} {r[6]:8d} {r[7]:7d}")


if __name__ == "__main__":
    main()
