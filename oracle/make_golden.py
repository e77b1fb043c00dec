"""Generate tests/golden/*.npz by executing the UNMODIFIED reference modules from /root/reference.

TEST INFRASTRUCTURE.  Run once in the build container (the reference tree does not exist on the
GPU box):  python oracle/make_golden.py

The reference has no tests or golden vectors of its own, so these files are the
pin for oracle/stylegan2_oracle.py and, through it, for the CUDA path.  The single shim applied
is `torch.Tensor.cuda = identity`, needed because models/stylegan2/op/fused_act.py:25 calls
input.cuda() unconditionally and this container has no GPU.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("W2E_REFERENCE", "/root/reference")

from oracle import synth  # noqa: E402


def import_reference():
    sys.path[:0] = [REF, os.path.join(REF, "attention")]
    torch.Tensor.cuda = lambda self, *a, **k: self
    import attention_model  # attention/attention_model.py (superset generator)
    import importlib
    up_mod = importlib.import_module("models.stylegan2.op.upfirdn2d")
    from models.stylegan2.op import fused_leaky_relu
    return attention_model, up_mod.upfirdn2d_native, fused_leaky_relu


def sd_checksum(sd):
    return np.array([float(sum(v.double().abs().sum() for v in sd.values())),
                     float(sum(v.double().sum() for v in sd.values()))])


def sub(t, step=97):
    return t.detach().reshape(-1)[::step].contiguous().numpy()


def stats(t):
    t = t.detach().double()
    return np.array([t.sum().item(), t.abs().sum().item(), t.abs().max().item()])


UPFIRDN_CASES = [
    # (shape, kernel spec, up_x, up_y, down_x, down_y, px0, px1, py0, py1)
    ((2, 3, 9, 9), "blur4", 1, 1, 1, 1, 1, 1, 1, 1),      # Blur after up-conv, model.py:206
    ((2, 3, 8, 8), "blur4", 2, 2, 1, 1, 2, 1, 2, 1),      # Upsample(skip), model.py:31-49
    ((1, 4, 16, 16), "blur1", 1, 1, 2, 2, 1, 1, 1, 1),    # Downsample, model.py:52-70
    ((1, 2, 10, 12), "asym35", 1, 1, 1, 1, 2, 2, 1, 1),
    ((1, 2, 10, 12), "asym35", 2, 3, 1, 2, 3, 0, 1, 2),
    ((1, 2, 12, 11), "asym35", 1, 1, 1, 1, -1, 2, 2, -2),  # negative pad = crop
    ((2, 2, 7, 5), "asym22", 3, 2, 2, 3, 1, 1, 0, 2),
    ((1, 1, 6, 6), "blur4", 1, 1, 1, 1, 2, 2, 2, 2),      # pad of the blur backward
]


def make_kernel_spec(name):
    if name == "blur4":
        return synth.blur_kernel_2d(gain=4.0)
    if name == "blur1":
        return synth.blur_kernel_2d(gain=1.0)
    if name == "asym35":
        return synth.make_tensor((3, 5), 91)
    if name == "asym22":
        return synth.make_tensor((2, 2), 92)
    raise KeyError(name)


def gen_ops(am, upfirdn2d_native, fused_leaky_relu, out):
    g = {}
    for i, (shape, kspec, *geom) in enumerate(UPFIRDN_CASES):
        x = synth.make_tensor(shape, 100 + i)
        k = make_kernel_spec(kspec)
        g[f"upfirdn_{i}"] = upfirdn2d_native(x, k, *geom).numpy()
    for i, (shape, cdim) in enumerate([((5, 7), 7), ((3, 4, 6), 6), ((2, 5, 4, 3), 5)]):
        x = synth.make_tensor(shape, 200 + i)
        b = synth.make_tensor((cdim,), 210 + i)
        g[f"flrelu_{i}"] = fused_leaky_relu(x, b).numpy()
        g[f"flrelu_{i}_args"] = fused_leaky_relu(x, b, 0.1, 1.5).numpy()
    # ModulatedConv2d / StyledConv / ToRGB on small channel counts
    for i, (cin, cout, k, up, demod, h) in enumerate([(8, 6, 3, False, True, 7), (8, 6, 3, True, True, 5),
                                                       (8, 3, 1, False, False, 6), (16, 16, 3, True, True, 4)]):
        m = am.ModulatedConv2d(cin, cout, k, 12, demodulate=demod, upsample=up)
        with torch.no_grad():
            m.weight.copy_(synth.make_tensor(tuple(m.weight.shape), 300 + i))
            m.modulation.weight.copy_(synth.make_tensor(tuple(m.modulation.weight.shape), 310 + i))
            m.modulation.bias.copy_(1 + synth.make_tensor((cin,), 320 + i, 0.1))
        x = synth.make_tensor((2, cin, h, h), 330 + i)
        w = synth.make_tensor((2, 12), 340 + i)
        y, s = m(x, w)
        g[f"modconv_{i}_y"] = y.detach().numpy()
        g[f"modconv_{i}_s"] = s.detach().numpy()
        y2, _ = m(x, s.detach() * 1.1, input_is_stylespace=True)
        g[f"modconv_{i}_y_ss"] = y2.detach().numpy()
    np.savez_compressed(os.path.join(out, "ops.npz"), **g)
    print("ops.npz:", len(g), "arrays")


def build_generator(am, size, cm, seed, perturbed):
    sd = synth.make_state_dict(size, channel_multiplier=cm, seed=seed, perturbed=perturbed)
    gen = am.Generator(size, 512, 8, channel_multiplier=cm)
    missing = gen.load_state_dict(sd, strict=True)   # proves key/shape compatibility of synth
    gen.eval()
    return gen, sd


def gen_generator32(am, out):
    size, batch = 32, 2
    gen, sd = build_generator(am, size, 2, seed=0, perturbed=True)
    n_latent = gen.n_latent
    wplus = synth.make_wplus(batch, n_latent, seed=2)
    g = {"sd_checksum": sd_checksum(sd)}
    with torch.no_grad():
        img, latent, styles, feats = gen([wplus], input_is_latent=True, randomize_noise=False,
                                         return_features=True)
        g["img_wplus"] = img.numpy()
        for i, f in enumerate(feats):
            g[f"feat_{i}_sub"] = sub(f)
            g[f"feat_{i}_stats"] = stats(f)
        for i, s in enumerate(styles):
            g[f"style_{i}"] = s.numpy()
        # stylespace round trip and an "edited" stylespace forward
        img_ss, _ = gen([styles], input_is_stylespace=True, randomize_noise=False)
        g["img_stylespace"] = img_ss.numpy()
        edited = [s * (1 + 0.05 * synth.make_tensor(tuple(s.shape), 500 + i)) for i, s in enumerate(styles)]
        img_ed, _, _, feats_ed = gen([edited], input_is_stylespace=True, randomize_noise=False,
                                     return_features=True)
        g["img_edited"] = img_ed.numpy()
        # blends (attention/attention_model.py:546-549 and siblings)
        for tag, layer, msize, binary in [("a", 7, 16, False), ("b", 6, 8, False), ("c", 9, 12, True),
                                          ("d", 1, 4, False)]:
            mask = synth.make_mask(batch, msize, seed=3 + layer, binary=binary)
            img_b, _, _, feats_b = gen([edited], input_is_stylespace=True, randomize_noise=False,
                                       return_features=True, attention_layer=layer,
                                       attention_map=mask, feature_map=feats)
            g[f"img_blend_{tag}"] = img_b.numpy()
            g[f"blend_{tag}_feat_stats"] = np.stack([stats(f) for f in feats_b])
        # the same blend driven from W+ (non-stylespace branch, :538-563 / :597-630)
        mask = synth.make_mask(batch, 16, seed=10)
        w_ed = wplus + 0.1 * synth.make_tensor(tuple(wplus.shape), 600)
        img_bw, _, _, _ = gen([w_ed], input_is_latent=True, randomize_noise=False, return_features=True,
                              attention_layer=7, attention_map=mask, feature_map=feats)
        g["img_blend_wplus"] = img_bw.numpy()
        # z input through the mapping network, with truncation and with style mixing
        z = synth.make_z(batch, seed=2)
        z2 = synth.make_z(batch, seed=3)
        mean_w = gen.style(synth.make_z(64, seed=9)).mean(0, keepdim=True)
        g["mean_w"] = mean_w.numpy()
        img_z, lat_z, _ = gen([z], truncation=0.7, truncation_latent=mean_w, randomize_noise=False,
                              return_latents=True)
        g["img_z_trunc"] = img_z.numpy()
        g["latent_z_trunc"] = lat_z.numpy()
        img_mix, _ = gen([z, z2], inject_index=3, randomize_noise=False)
        g["img_z_mix"] = img_mix.numpy()
        # explicit noise list
        noise = [synth.make_tensor((1, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)), 700 + i)
                 for i in range(gen.num_layers)]
        img_n, _ = gen([wplus], input_is_latent=True, noise=noise)
        g["img_noise_list"] = img_n.numpy()

    # gradients (consumer: run_attention.py:1419, coach.py:91) with a seeded upstream dL/dimage
    upstream = synth.make_tensor((batch, 3, size, size), 4) / (batch * 3 * size * size)
    wp = wplus.clone().requires_grad_(True)
    img, _ = gen([wp], input_is_latent=True, randomize_noise=False)
    (img * upstream).sum().backward()
    g["grad_wplus"] = wp.grad.numpy()
    st = [s.clone().requires_grad_(True) for s in edited]
    mask = synth.make_mask(batch, 16, seed=10).requires_grad_(True)
    img, _, _, _ = gen([st], input_is_stylespace=True, randomize_noise=False, return_features=True,
                       attention_layer=7, attention_map=mask, feature_map=feats)
    (img * upstream).sum().backward()
    for i, s in enumerate(st):
        g[f"grad_style_{i}"] = s.grad.numpy()
    g["grad_mask"] = mask.grad.numpy()
    np.savez_compressed(os.path.join(out, "generator32.npz"), **g)
    print("generator32.npz:", len(g), "arrays")


def gen_generator128(am, out):
    # channel-changing layers (512 -> 256 -> 128 with channel_multiplier=1 ... see model.py:392-402)
    size, batch = 128, 1
    gen, sd = build_generator(am, size, 1, seed=5, perturbed=True)
    wplus = synth.make_wplus(batch, gen.n_latent, seed=6)
    g = {"sd_checksum": sd_checksum(sd)}
    with torch.no_grad():
        img, _, styles, feats = gen([wplus], input_is_latent=True, randomize_noise=False,
                                    return_features=True)
    g["img_wplus"] = img.numpy()
    g["feat_stats"] = np.stack([stats(f) for f in feats])
    np.savez_compressed(os.path.join(out, "generator128.npz"), **g)
    print("generator128.npz:", len(g), "arrays")


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    am, upfirdn2d_native, fused_leaky_relu = import_reference()
    gen_ops(am, upfirdn2d_native, fused_leaky_relu, out)
    gen_generator32(am, out)
    gen_generator128(am, out)


if __name__ == "__main__":
    main()
