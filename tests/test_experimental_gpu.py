"""A/B checks of kernel variants that are in the tree but NOT yet measured on hardware and therefore off by
default.  Skipped unless W2E_TEST_EXPERIMENTAL=1 (run them first thing when a GPU is available):

    W2E_TEST_EXPERIMENTAL=1 python -m pytest tests/test_experimental_gpu.py -q -m gpu
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("W2E_TEST_EXPERIMENTAL") != "1", reason="experimental variants: opt-in")]

SCRIPT = r"""
import sys, torch
sys.path.insert(0, {root!r})
import where2edit_b200 as w2e
from oracle import synth
size, batch, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
gen = w2e.Generator(size, 512, 8, channel_multiplier=2, precision="bf16")
gen.load_state_dict(synth.make_state_dict(size, seed=0, perturbed=True), strict=True)
gen = gen.to("cuda:0").eval()
with torch.no_grad():
    img, _ = gen([synth.make_wplus(batch, gen.n_latent, seed=2).to("cuda:0")], input_is_latent=True, randomize_noise=False)
gen._engine.assert_ok()
torch.save(img.cpu(), out)
"""


@pytest.mark.parametrize("size,batch", [(256, 2), (1024, 1)])
def test_blur_v2_is_bit_identical_to_the_default_kernel(tmp_path, size, batch):
    """W2E_BLUR_V2=1 (templated channel tile, predicate-free interior body; csrc/nhwc_ops.cu) must reproduce the
    default blur kernel bit for bit through the whole bf16 engine (channel tiles 128 / 64 / 32, edge tiles)."""
    import torch
    outs = []
    for flag in ("0", "1"):
        path = str(tmp_path / f"img_{flag}.pt")
        env = dict(os.environ, W2E_BLUR_V2=flag)
        p = subprocess.run([sys.executable, "-c", SCRIPT.format(root=ROOT), str(size), str(batch), path], env=env,
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(torch.load(path))
    assert torch.isfinite(outs[0]).all() and torch.equal(outs[0], outs[1])
