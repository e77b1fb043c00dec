"""CPU checks of the drop-in surface: install_as_reference() next to the reference's own packages, and the
rosinality-format checkpoint contract (`ckpt["g_ema"]`, strict=False; attention/run_attention.py:982-986)."""
import os
import subprocess
import sys

import pytest
import torch

import where2edit_b200 as w2e
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("W2E_REFERENCE", "/root/reference")

SCRIPT = r"""
import sys
sys.path[:0] = [{root!r}, {ref!r}, {ref!r} + "/attention"]
import where2edit_b200
where2edit_b200.install_as_reference()          # the documented first line of a reference script (INTEGRATION.md)
# the reference's other packages must keep resolving (criteria/id_loss.py:4, models/psp.py:8, run_attention.py:29)
from models.facial_recognition.model_irse import Backbone
from models.encoders import psp_encoders
from models.stylegan2.model import Generator, EqualLinear, PixelNorm
from models.stylegan2.op import fused_leaky_relu, upfirdn2d, FusedLeakyReLU
import attention_model
from mapper import latent_mappers               # imports EqualLinear, PixelNorm from models.stylegan2.model
assert Generator is where2edit_b200.Generator and attention_model.Generator is where2edit_b200.Generator
assert latent_mappers.EqualLinear is where2edit_b200.EqualLinear
assert psp_encoders.EqualLinear is where2edit_b200.EqualLinear
assert Backbone.__module__ == "models.facial_recognition.model_irse"
print("DROPIN OK")
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_install_as_reference_keeps_the_reference_packages_importable():
    p = subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT, ref=REF)], capture_output=True, text=True,
                       timeout=300)
    assert p.returncode == 0 and "DROPIN OK" in p.stdout, p.stdout[-1500:] + p.stderr[-3000:]


def test_install_as_reference_without_the_reference_tree():
    code = ("import sys; sys.path.insert(0, %r); import where2edit_b200 as w; w.install_as_reference();"
            "from models.stylegan2.op import upfirdn2d; from attention.attention_model import Generator;"
            "assert Generator is w.Generator; print('STUB OK')") % ROOT
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd="/")
    assert p.returncode == 0 and "STUB OK" in p.stdout, p.stderr[-2000:]


def test_rosinality_checkpoint_layout_loads(tmp_path):
    """A rosinality-format file {"g_ema": state_dict, "g": ..., "latent_avg": ...} loads exactly as the reference
    does it (run_attention.py:982-986: torch.load -> g_ema.load_state_dict(ckpt["g_ema"], strict=False)), with no
    missing or unexpected key, and round-trips every tensor bit-exactly.  Older checkpoints lack the `noises.*`
    buffers: strict=False must then report ONLY those as missing."""
    sd = synth.make_state_dict(64, seed=11, perturbed=True)
    path = str(tmp_path / "stylegan2-ffhq-config-f.pt")
    torch.save({"g_ema": sd, "g": {}, "d": {}, "latent_avg": torch.zeros(512)}, path)
    ckpt = torch.load(path, map_location="cpu")
    g = w2e.Generator(64, 512, 8, channel_multiplier=2)
    res = g.load_state_dict(ckpt["g_ema"], strict=False)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in g.state_dict().items():
        assert torch.equal(v, sd[k]), k
    old = {k: v for k, v in sd.items() if not k.startswith("noises.")}
    g2 = w2e.Generator(64, 512, 8, channel_multiplier=2)
    res = g2.load_state_dict(old, strict=False)
    assert sorted(res.missing_keys) == sorted(k for k in sd if k.startswith("noises.")) and not res.unexpected_keys
