"""A/B check of the second blur kernel (bit-identity with the first) and the whole cluster-style mapper
against the reference golden.  Both run by default."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

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
    """The two blur kernels of csrc/nhwc_ops.cu (W2E_BLUR_V2=0: run-time tile shape; W2E_BLUR_V2=1: templated
    channel tile with a predicate-free interior body) must agree bit for bit through the whole bf16 engine
    (channel tiles 128 / 64 / 32, edge tiles)."""
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


@pytest.mark.parametrize("name", ["same_res", "upsampled"])
def test_cluster_style_mapper_matches_reference_golden(name):
    """mappers.ClusterStyleMapper (the reference's cluster-style mapper on this package's modules and region
    kernels) end to end against tests/golden/cluster_mapper.npz (written by the unmodified reference class)."""
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import where2edit_b200 as w2e
    from oracle import cluster_mapper_oracle as cmo
    from oracle import synth
    from where2edit_b200 import mappers
    dev = "cuda:0"
    g = np.load(os.path.join(ROOT, "tests", "golden", "cluster_mapper.npz"))
    size, clusters, cluster_layer, attention_layer = (int(v) for v in g[f"{name}/cfg"])
    gen = w2e.Generator(32, 512, 8, channel_multiplier=2)
    gen.load_state_dict(synth.make_state_dict(32, seed=0, perturbed=True), strict=True)
    gen = gen.to(dev).eval()
    with torch.no_grad():
        _, _, styles, feats = gen([synth.make_wplus(2, gen.n_latent, seed=2).to(dev)], input_is_latent=True,
                                  randomize_noise=False, return_features=True)
        feats = list(feats) + [gen.input.input.repeat(2, 1, 1, 1)]
    m = mappers.ClusterStyleMapper(gen.n_latent, 1024, 512, attention_layer=attention_layer, cluster_layer=cluster_layer,
                                   clusters=clusters, cluster_dim=576)
    with torch.no_grad():
        for key, p in m.named_parameters():
            if key != "initial_bias":
                p.copy_(cmo.seeded_value(key, tuple(p.shape)))
        m.initial_bias.fill_(float(g[f"{name}/bias"]))
    m = m.to(dev).eval()
    m.store_clusters(torch.from_numpy(g[f"{name}/centres"]).to(dev))
    text = torch.from_numpy(g[f"{name}/text"]).to(dev)
    x = [torch.cat([text.unsqueeze(1), s[:, :, :, 0, 0]], dim=-1) for s in styles]
    with torch.no_grad():
        out, final, (loss_delta, loss_reg, loss_tv) = m(x, feats, size)
    got = torch.stack([s[:, 0, :, 0, 0] for s in out]).cpu().numpy()
    want = g[f"{name}/styles_out"]
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-4 * np.abs(want).max())
    np.testing.assert_allclose(final.cpu().numpy(), g[f"{name}/final"], rtol=0, atol=1e-4)
    np.testing.assert_allclose([float(loss_delta), float(loss_reg), float(loss_tv)], g[f"{name}/losses"], rtol=1e-3)


def test_cluster_style_mapper_trains_every_parameter():
    """Every parameter the reference trains (mapper linears, text branches, the 1x1 StyledConv attention heads'
    weight / noise weight / activation bias, run_attention.py:725-735) receives a gradient after
    one backward through styles, attention map and losses; the unused CA_NET containers (:717) do not."""
    import torch
    sys.path.insert(0, ROOT)
    import where2edit_b200 as w2e
    from oracle import synth
    from where2edit_b200 import mappers
    dev = "cuda:0"
    gen = w2e.Generator(32, 512, 8, channel_multiplier=2)
    gen.load_state_dict(synth.make_state_dict(32, seed=0, perturbed=True), strict=True)
    gen = gen.to(dev).eval()
    with torch.no_grad():
        _, _, styles, feats = gen([synth.make_wplus(2, gen.n_latent, seed=2).to(dev)], input_is_latent=True,
                                  randomize_noise=False, return_features=True)
        feats = list(feats) + [gen.input.input.repeat(2, 1, 1, 1)]
    m = mappers.ClusterStyleMapper(gen.n_latent, 1024, 512, attention_layer=7, cluster_layer=7, clusters=4,
                                   cluster_dim=feats[6].shape[1] + 64).to(dev).train()
    text = torch.randn(2, 512, device=dev)
    x = [torch.cat([text.unsqueeze(1), s[:, :, :, 0, 0]], dim=-1) for s in styles]
    out, final, (loss_delta, loss_reg, loss_tv) = m(x, feats, 16)
    loss = sum(s.square().mean() for s in out) + final.square().mean() + loss_delta + loss_reg + loss_tv
    loss.backward()
    # not trained by the reference either: the CA_NET containers (never called, :717) and the heads' own modulation
    # linears (the heads are driven with input_is_stylespace=True, :805/837/845, which skips them)
    missing = [n for n, p in m.named_parameters()
               if p.grad is None and not n.startswith("mapper_textca_") and ".conv.modulation." not in n]
    assert not missing, missing
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)


def _cluster_mapper_setup(noise_weight):
    import torch
    sys.path.insert(0, ROOT)
    import where2edit_b200 as w2e
    from oracle import synth
    from where2edit_b200 import mappers
    dev = "cuda:0"
    gen = w2e.Generator(64, 512, 8, channel_multiplier=2)
    gen.load_state_dict(synth.make_state_dict(64, seed=0, perturbed=True), strict=True)
    gen = gen.to(dev).eval()
    with torch.no_grad():
        _, _, styles, feats = gen([synth.make_wplus(3, gen.n_latent, seed=2).to(dev)], input_is_latent=True,
                                  randomize_noise=False, return_features=True)
        feats = list(feats) + [gen.input.input.repeat(3, 1, 1, 1)]
    torch.manual_seed(5)
    m = mappers.ClusterStyleMapper(gen.n_latent, 1024, 512, attention_layer=7, cluster_layer=7, clusters=4, channel_multiplier=2,
                                   cluster_dim=feats[6].shape[1] + 2 * (feats[6].shape[1] // 16)).to(dev).train()
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.endswith("noise.weight"):
                p.fill_(noise_weight)
            elif name.endswith("activate.bias"):
                p.normal_(0, 0.2)
    text = torch.randn(3, 512, device=dev)
    x = [torch.cat([text.unsqueeze(1), s[:, :, :, 0, 0]], dim=-1) for s in styles]
    return m, x, feats


@pytest.mark.parametrize("size,noise_weight", [(16, 0.0), (24, 0.3), (64, 0.3)])
def test_grouped_attention_heads_match_the_per_head_modules(size, noise_weight):
    """w2e_attn_heads_fwd / _bwd (one launch for all heads, only the pixels that survive the nearest resize) against one
    StyledConv module call per head on the full-resolution maps -- the reference's formulation, run_attention.py:803-841:
    outputs, and the gradient of every trained parameter.  Sizes below / between / at the feature resolutions
    (4^2..64^2, non-integer ratios included); the noise of both paths comes from the same generator state."""
    import torch
    m, x, feats = _cluster_mapper_setup(noise_weight)
    results = {}
    for fused in (True, False):
        m.fused_heads, m.sample_first = fused, fused
        m.zero_grad(set_to_none=True)
        torch.manual_seed(11)
        out, final, (loss_delta, loss_reg, loss_tv) = m(x, feats, size)
        loss = final.square().mean() + 0.1 * loss_reg + 0.1 * loss_tv + sum(s.square().mean() for s in out)
        loss.backward()
        results[fused] = (final.detach(), {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None})
    f1, g1 = results[True]
    f0, g0 = results[False]
    # (with noise the full-resolution modules draw it at the feature resolution and the resize keeps one value in k^2:
    #  another sample of the same distribution -- only the noise-free case is comparable value by value)
    if noise_weight == 0.0:
        assert float((f1 - f0).abs().max()) <= 1e-5
        assert set(g1) == set(g0)
        for n in g0:
            if n.endswith("noise.weight"):   # d/d(noise weight) = sum g * noise: other noise values at full resolution
                continue
            scale = float(g0[n].abs().max()) + 1e-12
            assert float((g1[n] - g0[n]).abs().max()) <= 2e-3 * scale + 1e-9, n
    else:
        # same noise values in both paths when nothing is down-sampled after the head: sample_first on the module path
        m.fused_heads, m.sample_first = False, True
        m.zero_grad(set_to_none=True)
        torch.manual_seed(11)
        out, final, (loss_delta, loss_reg, loss_tv) = m(x, feats, size)
        loss = final.square().mean() + 0.1 * loss_reg + 0.1 * loss_tv + sum(s.square().mean() for s in out)
        loss.backward()
        g2 = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
        assert float((f1 - final.detach()).abs().max()) <= 1e-5
        assert set(g1) == set(g2)
        for n in g2:
            scale = float(g2[n].abs().max()) + 1e-12
            assert float((g1[n] - g2[n]).abs().max()) <= 2e-3 * scale + 1e-9, n


@pytest.mark.parametrize("size", [12, 20, 32])
def test_grouped_attention_heads_match_the_oracle(size):
    """region.attention_heads (w2e_attn_heads_fwd) against the CPU oracle's restatement of one reference head
    (StyledConv(C, 32, 1) with a stylespace input, run_attention.py:805 / :837) + F.interpolate(nearest): feature maps
    below, at and above `size`, non-integer resize ratios, explicit noise"""
    import torch
    sys.path.insert(0, ROOT)
    from oracle import cluster_mapper_oracle as cmo
    from oracle import stylegan2_oracle as orc
    from where2edit_b200 import region
    dev = "cuda:0"
    g = torch.Generator().manual_seed(size)
    b = 2
    shapes = [(512, 4), (256, 16), (64, 32), (32, 48)]
    feats, weights, styles, biases, nws, noises, want = [], [], [], [], [], [], []
    for i, (c, h) in enumerate(shapes):
        f = torch.randn(b, c, h, h, generator=g)
        w = torch.randn(1, 32, c, 1, 1, generator=g)
        st = 1 + 0.3 * torch.randn(b, c, generator=g)
        bias = 0.2 * torch.randn(32, generator=g)
        nw = torch.tensor([0.25 * (i + 1)])
        r = min(h, size)
        nz = torch.randn(b, 1, r, r, generator=g)
        # oracle: the head at the feature's own resolution with the noise it would have drawn there, then the resize.  For a
        # map larger than `size` the kept pixels' noise is what matters: place the [r, r] noise at the kept positions.
        idx = torch.nn.functional.interpolate(torch.arange(h * h, dtype=torch.float32).view(1, 1, h, h), size).long().view(-1)
        full = torch.zeros(b, 1, h * h)
        if h > size:
            full[:, :, idx] = nz.view(b, 1, -1)
        else:
            full = nz.view(b, 1, -1).clone()
        sd = {"p.conv.weight": w, "p.conv.modulation.weight": None, "p.conv.modulation.bias": None,
              "p.noise.weight": nw, "p.activate.bias": bias}
        out, _ = orc._styled_conv({k: (v.double() if v is not None else None) for k, v in sd.items()}, "p", f.double(),
                                  st.double().view(b, 1, -1, 1, 1), full.view(b, 1, h, h).double(), False, True)
        want.append(torch.nn.functional.interpolate(out, size))
        feats.append(f.to(dev)); weights.append((w[0, :, :, 0, 0] / c ** 0.5).to(dev)); styles.append(st.to(dev))
        biases.append(bias.to(dev)); nws.append(nw.to(dev)); noises.append(nz.to(dev))
    got = region.attention_heads(feats, weights, styles, biases, nws, size, noises=noises).cpu().double()
    want = torch.cat(want, dim=1)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-5 * float(want.abs().max())


def test_cluster_style_mapper_forward_as_one_cuda_graph():
    """GraphedStep without parameters: the whole mapper forward (assign_clusters, grouped heads, text branches, region
    block) replayed as one graph launch gives the eager result, also after the inputs change"""
    import torch
    import where2edit_b200 as w2e
    m, x, feats = _cluster_mapper_setup(0.0)
    m.eval()
    with torch.no_grad():
        want_styles, want_map, want_losses = m(x, feats, 16)
        fast = w2e.GraphedStep(lambda xs, fs: m(xs, fs, 16), [x, feats])
        got_styles, got_map, got_losses = fast(x, feats)
        torch.cuda.synchronize()
        assert torch.equal(got_map, want_map)
        assert all(torch.equal(a, b) for a, b in zip(got_styles, want_styles))
        assert all(torch.equal(torch.as_tensor(a), torch.as_tensor(b)) for a, b in zip(got_losses, want_losses))
        x2 = [t * 1.1 for t in x]
        want2 = m(x2, feats, 16)[1].clone()
        got2 = fast(x2, feats)[1]
        torch.cuda.synchronize()
        assert torch.equal(got2, want2)
