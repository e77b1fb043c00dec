"""Generate tests/golden/fullsize.npz by executing the UNMODIFIED reference at the sizes BASELINE.json quotes
its configurations on (TEST INFRASTRUCTURE; run once in the build container, ~3 min of CPU):

    python oracle/make_fullsize_golden.py

  cfg1  Generator(256, 512, 8), batch 4, W+ forward                        (models/stylegan2/model.py:473-574)
  cfg2  Generator(1024, 512, 8, channel_multiplier=2), batch 2, W+ forward (attention/attention_model.py:473-676)
  cfg3  LevelsMapper edit  w_hat = w + 0.1 * mapper(w)  (mapper/scripts/inference.py:98, latent_mappers.py:47-82)
        -> styles of w_hat -> stylespace forward blended at layer 13 with a 64^2 mask into the features of w
  cfg4  1024^2, batch 1: dL/dW+ and dL/d(styles, mask) of the blended forward for a seeded upstream dL/dimage,
        by the reference's autograd in fp32 AND in fp64 (the fp64 run is the yardstick of the gradient tests)

Full 1024^2 images are 12.6 MB each, so the fixture keeps a regular sub-grid of every image (every 16th / 4th
pixel in each direction) plus whole-image statistics; the GPU tests compare the same sub-grid and statistics with
these reference values and the FULL image with the live oracle (which tests/test_oracle_golden.py pins to the
same vectors).  The only shim is `torch.Tensor.cuda = identity` (op/fused_act.py:25 calls .cuda()).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import make_golden as mg  # noqa: E402
from oracle import mapper_oracle as mo  # noqa: E402
from oracle import synth  # noqa: E402

GRID = {256: 4, 1024: 16}   # sub-grid step per image size


def grid(img):
    st = GRID[img.shape[-1]]
    return img.detach()[:, :, ::st, ::st].contiguous().numpy()


def image_stats(img):
    """per-sample, per-channel (sum, sum |.|, max |.|) in fp64"""
    t = img.detach().double()
    return torch.stack([t.sum((2, 3)), t.abs().sum((2, 3)), t.abs().amax((2, 3))], -1).numpy()


def upstream_grad(shape, seed=4):
    n = 1
    for s in shape:
        n *= s
    return synth.make_tensor(shape, seed) / n


def edited_styles(styles):
    return [s * (1 + 0.05 * synth.make_tensor(tuple(s.shape), 500 + i)) for i, s in enumerate(styles)]


def main():
    torch.set_num_threads(os.cpu_count())
    am, _, _ = mg.import_reference()
    from mapper.latent_mappers import LevelsMapper
    g = {}

    # ---------------------------------------------------------------- cfg1
    gen, sd = mg.build_generator(am, 256, 2, seed=0, perturbed=True)
    wplus = synth.make_wplus(4, gen.n_latent, seed=2)
    with torch.no_grad():
        img, _ = gen([wplus], input_is_latent=True, randomize_noise=False)
    g["cfg1/grid"], g["cfg1/stats"] = grid(img), image_stats(img)
    print("cfg1", tuple(img.shape), float(img.abs().max()))

    # ---------------------------------------------------------------- cfg2
    gen, sd = mg.build_generator(am, 1024, 2, seed=0, perturbed=True)
    g["sd_checksum"] = mg.sd_checksum(sd)
    wplus = synth.make_wplus(2, gen.n_latent, seed=2)
    with torch.no_grad():
        img, _, styles, feats = gen([wplus], input_is_latent=True, randomize_noise=False, return_features=True)
    g["cfg2/grid"], g["cfg2/stats"] = grid(img), image_stats(img)
    g["cfg2/feat_stats"] = np.stack([mg.stats(f) for f in feats])
    print("cfg2", tuple(img.shape), float(img.abs().max()))

    # ---------------------------------------------------------------- cfg3
    opts = types.SimpleNamespace(no_coarse_mapper=False, no_medium_mapper=False, no_fine_mapper=False)
    mapper = LevelsMapper(opts)
    mapper.load_state_dict({k: torch.from_numpy(v) for k, v in mo.mapper_state().items()}, strict=True)
    mask = synth.make_mask(2, 64, seed=3)
    with torch.no_grad():
        w_hat = wplus + 0.1 * mapper(wplus)
        _, _, styles_hat = gen([w_hat], input_is_latent=True, randomize_noise=False, return_latents=True)
        img3, _, _, feats3 = gen([styles_hat], input_is_stylespace=True, randomize_noise=False, return_features=True,
                                 attention_layer=13, attention_map=mask, feature_map=feats)
    g["cfg3/w_hat"] = w_hat.numpy()
    g["cfg3/grid"], g["cfg3/stats"] = grid(img3), image_stats(img3)
    g["cfg3/feat_stats"] = np.stack([mg.stats(f) for f in feats3])
    print("cfg3", float(img3.abs().max()), float((img3 - img).abs().max()))

    # ---------------------------------------------------------------- cfg4 (batch 1)
    w1 = wplus[:1]
    up = upstream_grad((1, 3, 1024, 1024))
    feats1 = [f[:1] for f in feats]
    ed1 = [s[:1] for s in edited_styles(styles)]
    mask1 = mask[:1]
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        gen_t = gen if dt == torch.float32 else gen.double()
        wp = w1.detach().clone().to(dt).requires_grad_(True)
        out, _ = gen_t([wp], input_is_latent=True, randomize_noise=False)
        (out * up.to(dt)).sum().backward()
        g[f"cfg4/grad_wplus_{tag}"] = wp.grad.numpy()
        st = [s.detach().clone().to(dt).requires_grad_(True) for s in ed1]
        mk = mask1.detach().clone().to(dt).requires_grad_(True)
        out, _, _, _ = gen_t([st], input_is_stylespace=True, randomize_noise=False, return_features=True,
                             attention_layer=13, attention_map=mk, feature_map=[f.to(dt) for f in feats1])
        (out * up.to(dt)).sum().backward()
        g[f"cfg4/grad_styles_{tag}"] = np.concatenate([s.grad.reshape(-1).numpy() for s in st])
        g[f"cfg4/grad_mask_{tag}"] = mk.grad.numpy()
        print("cfg4", tag, float(wp.grad.abs().max()), float(mk.grad.abs().max()))
    path = os.path.join(ROOT, "tests", "golden", "fullsize.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
