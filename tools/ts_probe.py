"""Bisection probe for the TS (TMA-store) epilogue of w2e_modconv_tc2[_rgb]: every case runs in its
own process (a CUDA fault poisons the context) and reports ok / max error against a torch fp32
evaluation of the same math on the GPU, or the CUDA error.

    python tools/ts_probe.py            # all cases
    python tools/ts_probe.py CASE_JSON  # one case (child mode)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # b, cin, cout, h, transposed, rgb, skip, want_out, want_mod
    dict(b=1, cin=64, cout=32, h=72, tr=0, rgb=1, skip=1, out=1, mod=1),
    dict(b=1, cin=64, cout=32, h=64, tr=0, rgb=1, skip=1, out=1, mod=1),
    dict(b=1, cin=64, cout=32, h=72, tr=0, rgb=1, skip=0, out=1, mod=1),
    dict(b=1, cin=64, cout=32, h=72, tr=0, rgb=1, skip=1, out=0, mod=1),
    dict(b=1, cin=32, cout=32, h=128, tr=0, rgb=1, skip=1, out=0, mod=0),
    dict(b=1, cin=64, cout=64, h=128, tr=0, rgb=1, skip=1, out=0, mod=1),
    dict(b=2, cin=128, cout=128, h=64, tr=0, rgb=1, skip=1, out=0, mod=1),
    dict(b=1, cin=128, cout=64, h=64, tr=1, rgb=0, skip=0, out=1, mod=0),
    dict(b=1, cin=64, cout=32, h=128, tr=1, rgb=0, skip=0, out=1, mod=0),
    dict(b=2, cin=256, cout=128, h=32, tr=1, rgb=0, skip=0, out=1, mod=0),
    dict(b=2, cin=128, cout=128, h=64, tr=0, rgb=0, skip=0, out=1, mod=1),
    dict(b=3, cin=512, cout=256, h=20, tr=1, rgb=0, skip=0, out=1, mod=0),
    dict(b=2, cin=512, cout=512, h=8, tr=1, rgb=0, skip=0, out=1, mod=0),
]


def child(case):
    import torch
    import torch.nn.functional as F
    import where2edit_b200 as w2e
    from where2edit_b200 import _native as N, engine as E, functional as K
    dev = "cuda:0"
    torch.manual_seed(1)
    b, cin, cout, h = case["b"], case["cin"], case["cout"], case["h"]
    gen = w2e.Generator(8, 512, 1).to(dev)
    eng = E.SynthesisEngine(gen)
    m = w2e.StyledConv(cin, cout, 3, 16, upsample=bool(case["tr"])).to(dev)
    with torch.no_grad():
        m.noise.weight.fill_(0.3)
        m.activate.bias.copy_(0.2 * torch.randn(cout))
    x = torch.randn(b, cin, h, h, device=dev)
    s = 1 + 0.3 * torch.randn(b, cin, device=dev)
    nxt = (1 + 0.3 * torch.randn(b, cout, device=dev)).contiguous()
    pw = eng._tc_weight(m.conv)
    d = K.demod_coefficients(s, pw.wsq)
    xs = eng._to_nhwc(x, s.contiguous(), b)
    xq = (x * s[:, :, None, None]).to(torch.bfloat16).float()
    wq = pw.tc.float()  # [9, cout, cin]
    wk = wq.reshape(3, 3, cout, cin).permute(2, 3, 0, 1).contiguous()
    res = {}
    if case["tr"]:
        z, _ = eng._conv2(xs, pw, d, None, None, None, None, True, False, True, N.ACT_NONE)
        torch.cuda.synchronize()
        eng.assert_ok()
        ref = F.conv_transpose2d(xq, wk.permute(1, 0, 2, 3), stride=2) * d[:, :, None, None]
        res["out"] = float((z.float().permute(0, 3, 1, 2) - ref).abs().max() / ref.abs().max())
    else:
        noise = torch.randn(1, 1, h, h, device=dev)
        bias = m.activate.bias.detach()
        ref = F.conv2d(xq, wk, padding=1) * d[:, :, None, None] + 0.3 * noise + bias[None, :, None, None]
        ref = F.leaky_relu(ref, 0.2) * 2 ** 0.5
        if case["rgb"]:
            rgbm = w2e.ToRGB(cout, 16).to(dev)
            with torch.no_grad():
                rgbm.bias.copy_(0.1 * torch.randn(1, 3, 1, 1))
            s_rgb = (1 + 0.3 * torch.randn(b, cout, device=dev)).contiguous()
            skip = torch.randn(b, 3, h // 2, h // 2, device=dev) if case["skip"] else None
            out, out_mod, rgb = eng._conv2_rgb(xs, pw, d, noise, m.noise.weight.detach(), bias,
                                               nxt if case["mod"] else None, bool(case["out"]), bool(case["mod"]), rgbm,
                                               s_rgb, skip)
            torch.cuda.synchronize()
            eng.assert_ok()
            wr = rgbm.conv.weight.detach()[0, :, :, 0, 0] / cout ** 0.5  # [3, cout]
            rr = torch.einsum("bchw,bc,oc->bohw", ref, s_rgb, wr) + rgbm.bias.detach()
            if skip is not None:
                from where2edit_b200.op import upfirdn2d
                rr = rr + upfirdn2d(skip, rgbm.upsample.kernel, up=2, pad=rgbm.upsample.pad)
            res["rgb"] = float((rgb - rr).abs().max() / rr.abs().max())
        else:
            out, out_mod = eng._conv2(xs, pw, d, noise, m.noise.weight.detach(), bias, nxt if case["mod"] else None,
                                      bool(case["out"]), bool(case["mod"]), False, N.ACT_LRELU)
            torch.cuda.synchronize()
            eng.assert_ok()
        if out is not None:
            res["out"] = float((out.float().permute(0, 3, 1, 2) - ref).abs().max() / ref.abs().max())
        if out_mod is not None:
            rm = ref * nxt[:, :, None, None]
            res["mod"] = float((out_mod.float().permute(0, 3, 1, 2) - rm).abs().max() / rm.abs().max())
    print("RESULT", json.dumps(res))


def main():
    if len(sys.argv) > 1 and sys.argv[1].startswith("{"):
        child(json.loads(sys.argv[1]))
        return
    sel = [int(a) for a in sys.argv[1:]] or range(len(CASES))
    for i in sel:
        case = CASES[i]
        p = subprocess.run([sys.executable, os.path.abspath(__file__), json.dumps(case)], capture_output=True, text=True,
                           timeout=300)
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
        err = [l for l in p.stderr.splitlines() if "Error" in l or "error" in l][-2:]
        print(i, case, "->", line[0] if line else f"FAILED rc={p.returncode} {err}", flush=True)


if __name__ == "__main__":
    main()
