"""Hardware probe for the shifted-descriptor scheme of modconv_tc2 (run on the B200 box):
for every knob combination (pitch of the haloed tile, base-offset field on/off) and a few layer
shapes, report the error of the v2 kernel against the already-validated v1 kernel."""
import itertools
import sys

import torch

sys.path.insert(0, ".")
import where2edit_b200 as w2e  # noqa: E402
from where2edit_b200 import _native as N  # noqa: E402
from where2edit_b200 import engine as E  # noqa: E402

DEV = "cuda:0"
torch.manual_seed(0)
gen = w2e.Generator(8, 512, 1).to(DEV)
eng = E.SynthesisEngine(gen)
lib = N.load()

SHAPES = [  # b, cin, cout, h, transposed
    (2, 64, 64, 32, False), (1, 32, 32, 40, False), (2, 512, 512, 16, False), (3, 128, 64, 8, False),
    (2, 64, 64, 16, True), (1, 32, 32, 24, True), (2, 512, 256, 8, True), (5, 64, 32, 4, False),
]
for (b, cin, cout, h, tr) in SHAPES:
    m = w2e.StyledConv(cin, cout, 3, 16, upsample=tr).to(DEV)
    with torch.no_grad():
        m.noise.weight.fill_(0.3)
        m.activate.bias.normal_(0, 0.2)
    pw = eng._tc_weight(m.conv)
    xs = torch.randn(b, h, h, cin, device=DEV).to(torch.bfloat16)
    d = 0.5 + torch.rand(b, cout, device=DEV)
    nxt = 0.5 + torch.rand(b, cout, device=DEV)
    oh = 2 * h + 1 if tr else h
    noise = torch.randn(1, 1, oh, oh, device=DEV)
    nw, bias = m.noise.weight.detach(), m.activate.bias.detach()
    if tr:
        ref = torch.full((b, oh, oh, cout), float("nan"), device=DEV, dtype=torch.bfloat16)
        for (py, px), taps in E._TAPS_UP.items():
            eng._conv(xs, pw, d, None, None, None, None, True, False, taps, (h, h), (oh, oh), (h + 1 - py, h + 1 - px),
                      2, py, px, N.ACT_NONE, out=ref)
        ref_mod = None
    else:
        ref, ref_mod = eng._conv(xs, pw, d, noise, nw, bias, nxt, True, True, E._TAPS_PLAIN, (h, h), (h, h), (h, h), 1,
                                 0, 0, N.ACT_LRELU)
    eng.assert_ok()
    for pitch, boff in itertools.product((10, 16), (0, 1)):
        lib.w2e_modconv_tc2_knobs(0)
        try:
            if tr:
                got, _ = eng._conv2(xs, pw, d, None, None, None, None, True, False, True, N.ACT_NONE)
                got_mod = None
            else:
                got, got_mod = eng._conv2(xs, pw, d, noise, nw, bias, nxt, True, True, False, N.ACT_LRELU)
            torch.cuda.synchronize()
            eng.assert_ok()
            err = (got.float() - ref.float()).abs().max().item()
            err2 = (got_mod.float() - ref_mod.float()).abs().max().item() if got_mod is not None else 0.0
            scale = ref.float().abs().max().item()
            nan = int(torch.isnan(got.float()).sum().item())
            print(f"shape {(b, cin, cout, h, tr)} pitch {pitch} base_off {boff}: max err {err:.4g} mod {err2:.4g} "
                  f"(scale {scale:.3g}) nan {nan}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"shape {(b, cin, cout, h, tr)} pitch {pitch} base_off {boff}: FAILED {e}", flush=True)
lib.w2e_modconv_tc2_knobs(0)
