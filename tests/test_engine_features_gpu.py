"""bf16 engine features around the kernels: CUDA-graph replay, the image written in the caller's dtype / buffer by the
last layer's epilogue, autocast safety of the public modules (the reference's --amp, run_attention.py:1068-1069),
a model on a non-current device, and the non-synchronising error check."""
import pytest
import torch

import where2edit_b200 as w2e
from conftest import max_abs
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build(size, precision="bf16", cm=2):
    gen = w2e.Generator(size, 512, 8, channel_multiplier=cm, precision=precision)
    gen.load_state_dict(synth.make_state_dict(size, seed=0, perturbed=True, channel_multiplier=cm), strict=True)
    gen = gen.to(DEV).eval()
    for p in gen.parameters():
        p.requires_grad_(False)
    return gen


@pytest.mark.parametrize("size,batch", [(64, 3), (1024, 1)])
def test_cuda_graph_replay_is_bit_identical_to_eager(size, batch):
    gen = build(size)
    w_a = synth.make_wplus(batch, gen.n_latent, seed=2).to(DEV)
    w_b = synth.make_wplus(batch, gen.n_latent, seed=3).to(DEV)
    with torch.no_grad():
        eager_a, _ = gen([w_a], input_is_latent=True, randomize_noise=False)
        eager_b, _ = gen([w_b], input_is_latent=True, randomize_noise=False)
    fast = w2e.GraphedGenerator(gen, [w_a], input_is_latent=True)
    img, none = fast([w_b])
    assert none is None and torch.equal(img, eager_b)
    img, _ = fast([w_a])
    assert torch.equal(img, eager_a)
    gen.assert_ok()
    with pytest.raises(ValueError):
        fast([w_a[:, :2]])


def test_cuda_graph_of_the_blended_edit_forward():
    gen = build(128)
    w = synth.make_wplus(2, gen.n_latent, seed=2).to(DEV)
    with torch.no_grad():
        _, _, styles, feats = gen([w], input_is_latent=True, randomize_noise=False, return_features=True)
        edited = [s * 1.05 for s in styles]
        edited2 = [s * 0.97 for s in styles]
        mask = synth.make_mask(2, 16, seed=3).to(DEV)
        mask2 = synth.make_mask(2, 16, seed=4, binary=True).to(DEV)
        kw = dict(input_is_stylespace=True, randomize_noise=False, attention_layer=7)
        want1, _ = gen([edited], attention_map=mask, feature_map=feats, **kw)
        want2, _ = gen([edited2], attention_map=mask2, feature_map=feats, **kw)
    fast = w2e.GraphedGenerator(gen, [edited], attention_map=mask, feature_map=feats, input_is_stylespace=True,
                                attention_layer=7)
    got2, _ = fast([edited2], attention_map=mask2)
    assert torch.equal(got2, want2)
    got1, _ = fast([edited], attention_map=mask)
    assert torch.equal(got1, want1)


def test_image_in_the_callers_dtype_and_buffer():
    """the last layer's epilogue writes the image as bf16 (== the fp32 image rounded once) and into a caller buffer"""
    gen = build(64)
    w = synth.make_wplus(2, gen.n_latent, seed=2).to(DEV)
    with torch.no_grad():
        img32, _ = gen([w], input_is_latent=True, randomize_noise=False)
        gen.set_image_output(torch.bfloat16)
        img16, _ = gen([w], input_is_latent=True, randomize_noise=False)
        assert img16.dtype == torch.bfloat16 and torch.equal(img16, img32.to(torch.bfloat16))
        buf = torch.full((2, 3, 64, 64), float("nan"), device=DEV, dtype=torch.bfloat16)
        gen.set_image_output(torch.bfloat16, buf)
        out, _ = gen([w], input_is_latent=True, randomize_noise=False)
        assert out.data_ptr() == buf.data_ptr() and torch.equal(buf, img16)
        # a blend at the last ToRGB takes the unfused path: same contract
        _, _, styles, feats = gen([w], input_is_latent=True, randomize_noise=False, return_features=True)
        gen.set_image_output(torch.float32)
        want, _ = gen([styles], input_is_stylespace=True, randomize_noise=False, attention_layer=len(feats),
                      attention_map=torch.ones(2, 1, 8, 8, device=DEV), feature_map=feats)
        gen.set_image_output(torch.bfloat16, buf)
        got, _ = gen([styles], input_is_stylespace=True, randomize_noise=False, attention_layer=len(feats),
                     attention_map=torch.ones(2, 1, 8, 8, device=DEV), feature_map=feats)
        assert got.data_ptr() == buf.data_ptr() and torch.equal(got, want.to(torch.bfloat16))
    gen.set_image_output(torch.float32)
    with pytest.raises(ValueError):
        gen.set_image_output(torch.float16)


def test_capture_layers_returns_only_the_requested_feature_maps():
    """Generator.forward(return_features=True, capture_layers=[...]): the requested maps are bit-identical to the full
    capture, the others are None, image and styles unchanged; the blended forward works from the one map it reads"""
    gen = build(128)
    w = synth.make_wplus(2, gen.n_latent, seed=6).to(DEV)
    mask = torch.rand(2, 1, 16, 16, device=DEV)
    with torch.no_grad():
        img_a, _, styles_a, feats_a = gen([w], input_is_latent=True, randomize_noise=False, return_features=True)
        layer = 9
        need = gen.blend_feature_layers(layer)
        assert need == [8, 10]                 # the up-convolution of layer 9 and the image of the next ToRGB (carry)
        img_b, _, styles_b, feats_b = gen([w], input_is_latent=True, randomize_noise=False, return_features=True,
                                          capture_layers=need + [2])
        assert torch.equal(img_a, img_b) and len(feats_b) == len(feats_a)
        assert all(torch.equal(a, b) for a, b in zip(styles_a, styles_b))
        for i, (a, b) in enumerate(zip(feats_a, feats_b)):
            assert (b is None) == (i not in need + [2])
            if b is not None:
                assert torch.equal(a, b)
        w2 = synth.make_wplus(2, gen.n_latent, seed=7).to(DEV)
        want, _ = gen([w2], input_is_latent=True, randomize_noise=False, attention_layer=layer, attention_map=mask,
                      feature_map=feats_a)
        got, _ = gen([w2], input_is_latent=True, randomize_noise=False, attention_layer=layer, attention_map=mask,
                     feature_map=feats_b)
        assert torch.equal(want, got)
    gen.set_precision("fp32")   # module path: same contract
    with torch.no_grad():
        _, _, _, feats_c = gen([w], input_is_latent=True, randomize_noise=False, return_features=True, capture_layers=[0])
    assert feats_c[0] is not None and all(f is None for f in feats_c[1:])


def test_cuda_graph_of_the_whole_optimisation_step():
    """GraphedStep: mapper forward -> generator forward -> loss -> backward as one graph launch; the gradients of the
    mapper parameters (and, second case, of a W+ latent optimised directly, run_attention.py:1233-1424) are bit-identical
    to the eager step, replay after replay, for new inputs"""
    import types
    from where2edit_b200 import mappers
    gen = build(64)
    torch.manual_seed(5)
    mapper = mappers.LevelsMapper(types.SimpleNamespace(no_coarse_mapper=False, no_medium_mapper=False,
                                                        no_fine_mapper=False)).to(DEV).train()
    ws = [synth.make_wplus(2, gen.n_latent, seed=s).to(DEV) for s in (3, 4, 5)]
    gimg = synth.make_tensor((2, 3, 64, 64), 9).to(DEV) / (3 * 64 * 64)

    def step(w):
        img, _ = gen([w + 0.1 * mapper(w)], input_is_latent=True, randomize_noise=False)
        loss = (img * gimg).sum()
        loss.backward()
        return loss

    def eager(w):
        for p in mapper.parameters():
            p.grad = None
        loss = step(w)
        return loss.detach().clone(), [p.grad.detach().clone() for p in mapper.parameters()]

    want = [eager(w) for w in ws]
    fast = w2e.GraphedStep(step, [ws[0]], params=mapper.parameters())
    for w, (loss_ref, grads_ref) in list(zip(ws, want)) + [(ws[0], want[0])]:
        loss = fast(w)
        torch.cuda.synchronize()
        assert torch.equal(loss.detach(), loss_ref)
        assert all(torch.equal(p.grad, g) for p, g in zip(mapper.parameters(), grads_ref))
    gen.assert_ok()
    # a latent optimised directly: the leaf is closed over, no inputs
    latent = ws[1].clone().requires_grad_(True)

    def latent_step():
        img, _ = gen([latent], input_is_latent=True, randomize_noise=False)
        (img * gimg).sum().backward()

    latent_step()
    ref = latent.grad.detach().clone()
    fast2 = w2e.GraphedStep(latent_step, [], params=[latent])
    fast2()
    torch.cuda.synchronize()
    assert torch.equal(latent.grad, ref)
    with torch.no_grad():
        latent.add_(0.01)                      # an optimiser step in place: the next replay sees the new latent
    fast2()
    g_graph = latent.grad.detach().clone()
    torch.cuda.synchronize()
    lat2 = latent.detach().clone().requires_grad_(True)
    img, _ = gen([lat2], input_is_latent=True, randomize_noise=False)
    (img * gimg).sum().backward()
    assert torch.equal(g_graph, lat2.grad)


@pytest.mark.parametrize("size", [64, 256, 1024])
def test_uint8_image_is_the_save_image_quantisation_of_the_fp32_image(size):
    """set_image_output(torch.uint8): the last layer's epilogue (64^2: the conversion after an fp32 ToRGB sum; 256^2: the
    staged epilogue; 1024^2: the pixel-pair kernel) quantises exactly like torchvision.utils.save_image(normalize=True,
    range=(-1, 1)) quantises the reference's results (run_attention.py:1470) -- bit-exact against that arithmetic applied
    to the fp32 image of the same kernels"""
    gen = build(size)
    w = synth.make_wplus(1, gen.n_latent, seed=4).to(DEV)
    with torch.no_grad():
        img, _ = gen([w], input_is_latent=True, randomize_noise=False)
        c = float(img.abs().max()) / 1.25          # the image is linear in the ToRGB weights: bring it to about [-1.25, 1.25]
        for m in [gen.to_rgb1] + list(gen.to_rgbs):
            m.conv.weight.div_(c)
            m.bias.div_(c)
        gen._engine = None
        img32, _ = gen([w], input_is_latent=True, randomize_noise=False)
        assert 1.0 < float(img32.abs().max()) < 1.6
        gen.set_image_output(torch.uint8)
        img8, _ = gen([w], input_is_latent=True, randomize_noise=False)
        gen.assert_ok()
    assert img8.dtype == torch.uint8 and img8.shape == img32.shape
    from oracle import save_image_oracle as so     # torchvision's make_grid(normalize) + save_image arithmetic, pinned on CPU
    want = torch.from_numpy(so.quantise_ref(img32.cpu().numpy())).to(DEV)
    assert torch.equal(img8, want)
    hist = torch.bincount(img8.flatten().long(), minlength=256)
    assert int((hist > 0).sum()) > 200               # the whole range is exercised, a clamped end included
    assert int(hist[0]) + int(hist[255]) > 0


@pytest.mark.parametrize("amp_dtype", [torch.float16, torch.bfloat16])
def test_public_modules_under_autocast(amp_dtype):
    """the reference's --amp wraps mapper + generator in torch.cuda.amp.autocast (run_attention.py:1231,1386): the
    drop-in modules must accept what autocast hands them (half-precision styles from the linears), return the
    reference's fp32 image, stay within the reduced-precision tolerance of the un-cast run, and backpropagate."""
    gen = build(32, precision="fp32")
    w = synth.make_wplus(2, gen.n_latent, seed=2).to(DEV)
    with torch.no_grad():
        want, _ = gen([w], input_is_latent=True, randomize_noise=False)
    wp = w.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=amp_dtype):
        got, _ = gen([wp], input_is_latent=True, randomize_noise=False)
        z = torch.randn(2, 512, device=DEV)
        img_z, _ = gen([z], randomize_noise=False)
        y = w2e.fused_leaky_relu(torch.randn(4, 8, device=DEV, dtype=amp_dtype), torch.zeros(8, device=DEV))
        u = w2e.upfirdn2d(torch.randn(1, 2, 8, 8, device=DEV, dtype=amp_dtype),
                          w2e.make_kernel([1, 3, 3, 1]).to(DEV), up=2, pad=(2, 1))
        loss = got.float().square().mean()
    loss.backward()
    c = float(want.abs().max())
    assert got.dtype == torch.float32 and max_abs(got.detach().cpu() / c, want.cpu() / c) <= 5e-2
    assert torch.isfinite(img_z).all() and torch.isfinite(y).all() and torch.isfinite(u).all()
    assert tuple(u.shape) == (1, 2, 16, 16)
    assert wp.grad is not None and torch.isfinite(wp.grad).all() and float(wp.grad.abs().max()) > 0


def test_model_on_a_non_current_device_and_error_poll():
    """the C-ABI calls launch on the tensors' device (device guard + per-device kernel attributes), and the
    pipeline-timeout flag of earlier calls is checked without a synchronisation (no false positive)"""
    gen = build(64)
    w = synth.make_wplus(1, gen.n_latent, seed=2).to(DEV)
    with torch.no_grad():
        a, _ = gen([w], input_is_latent=True, randomize_noise=False)
        for _ in range(3):   # polls the flag published by the previous call
            b, _ = gen([w], input_is_latent=True, randomize_noise=False)
    gen.assert_ok()
    assert torch.equal(a, b)
    if torch.cuda.device_count() > 1:
        gen1 = build(64).to("cuda:1")
        with torch.no_grad():          # cuda:0 stays the current device
            c, _ = gen1([w.to("cuda:1")], input_is_latent=True, randomize_noise=False)
        gen1.assert_ok()
        assert torch.equal(c.cpu(), a.cpu())
