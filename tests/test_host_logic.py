"""CPU-side checks: the C-ABI library loads and exports every symbol include/w2e.h declares, the
drop-in modules are state-dict compatible with the reference, and the host-side geometry/indexing
logic agrees with the oracle.  No GPU compute here."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

import where2edit_b200 as w2e
from where2edit_b200 import _native, functional as K
from oracle import stylegan2_oracle as orc
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "w2e.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(w2e_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_native.library_path())
    names = header_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"libw2e.so does not export {name}"
    assert set(names) == set(_native.PROTOTYPES), "ctypes prototypes out of sync with include/w2e.h"
    assert _native.load().w2e_version() >= 100


def test_no_cpu_path():
    g = w2e.Generator(16, 512, 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        g([torch.randn(1, 6, 512)], input_is_latent=True, randomize_noise=False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        w2e.upfirdn2d(torch.randn(1, 1, 4, 4), torch.ones(2, 2))


@pytest.mark.parametrize("size,cm", [(32, 2), (128, 1), (1024, 2)])
def test_state_dict_matches_reference_layout(size, cm):
    """synth.make_state_dict is strict-loaded into the reference Generator by oracle/make_golden.py,
    so equality with it is equality with the reference's keys and shapes."""
    if size == 1024:
        sd = {k: tuple(v.shape) for k, v in synth.make_state_dict(32, seed=0).items()}
        g = w2e.Generator(1024, 512, 8)
        ours = {k: tuple(v.shape) for k, v in g.state_dict().items()}
        assert len(ours) == 171
        assert ours["convs.15.conv.weight"] == (1, 32, 32, 3, 3)
        assert ours["to_rgbs.7.conv.weight"] == (1, 3, 32, 1, 1)
        assert ours["noises.noise_16"] == (1, 1, 1024, 1024)
        assert g.n_latent == 18 and g.num_layers == 17 and len(g.styled_layers()) == 26
        return
    sd = synth.make_state_dict(size, channel_multiplier=cm, seed=0)
    g = w2e.Generator(size, 512, 8, channel_multiplier=cm)
    ours = g.state_dict()
    assert sorted(ours.keys()) == sorted(sd.keys())
    for k in sd:
        assert tuple(ours[k].shape) == tuple(sd[k].shape), k
    g.load_state_dict(sd, strict=True)
    assert torch.equal(g.convs[0].conv.blur.kernel, sd["convs.0.conv.blur.kernel"])


def test_latent_row_schedule():
    g = w2e.Generator(1024, 512, 1)
    rows_w = g.latent_rows(False)
    rows_s = g.latent_rows(True)
    assert rows_s == list(range(26))
    assert rows_w[:5] == [0, 1, 1, 2, 3] and rows_w[-1] == 17 and max(rows_w) == g.n_latent - 1


def test_upfirdn2d_backward_geometry_matches_autograd():
    """w2e_upfirdn2d_bwd's geometry reproduces autograd of the oracle."""
    cases = [((1, 2, 9, 9), (4, 4), 1, 1, (1, 1)), ((1, 2, 8, 8), (4, 4), 2, 1, (2, 1)),
             ((1, 1, 12, 10), (3, 5), 1, 2, (2, 1)), ((1, 1, 7, 9), (4, 4), 2, 1, (3, 0)),
             ((1, 1, 9, 8), (2, 2), 3, 2, (1, 2))]
    for shape, kshape, up, down, pad in cases:
        x = synth.make_tensor(shape, 1).double().requires_grad_(True)
        k = synth.make_tensor(kshape, 2).double()
        y = orc.upfirdn2d_ref(x, k, up=up, down=down, pad=pad)
        gy = synth.make_tensor(tuple(y.shape), 3).double()
        y.backward(gy)
        kh, kw = kshape
        in_h, in_w = shape[2:]
        oh, ow = y.shape[2:]
        gpx0, gpy0 = kw - pad[0] - 1, kh - pad[0] - 1
        gpx1 = in_w * up - ow * down + pad[0] - up + 1
        gpy1 = in_h * up - oh * down + pad[0] - up + 1
        gx = orc.upfirdn2d_native_ref(gy, torch.flip(k, [0, 1]), down, down, up, up, gpx0, gpx1, gpy0, gpy1)
        assert gx.shape == x.shape
        assert (gx - x.grad).abs().max().item() < 1e-12


def test_separable_taps_factorisation():
    k2 = tuple(w2e.make_kernel([1, 3, 3, 1]).mul(4).reshape(-1).tolist())
    k1 = K.separable_taps(k2)
    np.testing.assert_allclose(k1, [0.25, 0.75, 0.75, 0.25], rtol=1e-6)
    assert K.separable_taps(tuple(synth.make_tensor((4, 4), 5).reshape(-1).tolist())) is None


def test_shared_weight_identity_against_oracle():
    """The reformulation used by the kernels (style on the activation, demod as an output scale,
    polyphase transposed conv) equals the reference formulation -- checked on CPU in fp64."""
    import torch.nn.functional as F
    for up in (False, True):
        cin, cout, h = 6, 5, 7
        weight = synth.make_tensor((1, cout, cin, 3, 3), 11).double()
        s = (1 + 0.3 * synth.make_tensor((2, cin), 12)).double()
        x = synth.make_tensor((2, cin, h, h), 13).double()
        blur = synth.blur_kernel_2d(gain=4.0).double()
        ref, _ = orc.modulated_conv2d_ref(x, s.reshape(2, 1, cin, 1, 1), weight, None, None, True, up, blur,
                                          input_is_stylespace=True)
        wt = weight[0] / (cin * 9) ** 0.5
        d = torch.rsqrt((s * s) @ wt.pow(2).sum((2, 3)).t() + 1e-8)
        xs = x * s[:, :, None, None]
        if not up:
            z = F.conv2d(xs, wt, padding=1)
        else:
            z = xs.new_zeros(2, cout, 2 * h + 1, 2 * h + 1)
            axis = {0: [(0, 0), (-1, 2)], 1: [(0, 1)]}
            xp = F.pad(xs, [1, 1, 1, 1])
            for py in (0, 1):
                for px in (0, 1):
                    gh, gw = h + 1 - py, h + 1 - px
                    acc = xs.new_zeros(2, cout, gh, gw)
                    for dy, ky in axis[py]:
                        for dx, kx in axis[px]:
                            patch = xp[:, :, 1 + dy:1 + dy + gh, 1 + dx:1 + dx + gw]
                            acc += torch.einsum("bchw,oc->bohw", patch, wt[:, :, ky, kx])
                    z[:, :, py::2, px::2] = acc
        y = z * d[:, :, None, None]
        if up:
            y = orc.upfirdn2d_ref(y, blur, pad=(1, 1))
        assert (y - ref).abs().max().item() < 1e-12


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) needs no GPU: one JSON line with the
    contract's keys, impl = reference, the oracle port as cpu_baseline and an e2e block repeating the value."""
    import json
    import subprocess
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "32", "--batch", "2",
                        "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["impl"] == "reference" and j["unit"] == "images/s" and j["value"] > 0 and j["warmup"] >= 3
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "model" not in j["config"]


def test_blend_feature_layers_names_the_maps_a_blended_forward_reads():
    """attention/attention_model.py:546-561: a blend at a convolution layer is followed by a blend of the next ToRGB image
    (`this_layer` carry), a blend at a ToRGB layer stands alone"""
    import where2edit_b200 as w2e
    gen = w2e.Generator(1024, 512, 8, channel_multiplier=2)
    kinds = [k for _, k in gen.styled_layers()]
    assert len(kinds) == 26 and kinds[:5] == ["conv", "rgb", "up", "conv", "rgb"]
    assert gen.blend_feature_layers(13) == [12, 13]      # the published configuration: conv at 64^2 + its ToRGB
    assert gen.blend_feature_layers(12) == [11, 13]      # an up-convolution: the next ToRGB comes after the plain conv
    assert gen.blend_feature_layers(14) == [13]          # a ToRGB layer
    assert gen.blend_feature_layers(26) == [25]
    with pytest.raises(ValueError):
        gen.blend_feature_layers(0)
    with pytest.raises(ValueError):
        gen.blend_feature_layers(27)
