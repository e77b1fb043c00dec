"""Generate tests/golden/cluster_mapper.npz: the UNMODIFIED reference cluster-style mapper
(`FullSpaceMapperFEATClusterLinStyle_Net`, attention/run_attention.py:703-893) on the captured features of the
reference generator, with the seeded parameters of oracle.cluster_mapper_oracle.seeded_value written into it.

TEST INFRASTRUCTURE; run once in the build container:  python oracle/make_cluster_mapper_golden.py
Shims as in oracle/make_region_golden.py (torch.Tensor.cuda = identity; stand-ins for the clip / torch_fidelity
imports, never executed on this path).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import cluster_mapper_oracle as cmo  # noqa: E402
from oracle import region_oracle, synth  # noqa: E402
from oracle.make_region_golden import import_reference  # noqa: E402

CASES = {
    # name: (size, clusters, cluster_layer, attention_layer, initial_bias)
    "same_res": (16, 6, 7, 7, 1.3),
    "upsampled": (32, 9, 7, 10, 1.0),
}


def main():
    am, ra = import_reference()
    gsize, batch = 32, 2
    gen = am.Generator(gsize, 512, 8, channel_multiplier=2).eval()
    gen.load_state_dict(synth.make_state_dict(gsize, seed=0, perturbed=True), strict=True)
    wplus = synth.make_wplus(batch, gen.n_latent, seed=2)
    with torch.no_grad():
        _, _, styles, features = gen([wplus], input_is_latent=True, randomize_noise=False, return_features=True)
        features = list(features)
        features.append(gen.input.input.repeat(batch, 1, 1, 1))                      # run_attention.py:1110
    out = {}
    for name, (size, clusters, cluster_layer, attention_layer, bias) in CASES.items():
        rng = np.random.default_rng(31 + size)
        blend = features[cluster_layer - 1].numpy()
        flat = region_oracle.cluster_features(blend)
        centres = (flat[rng.integers(0, flat.shape[0], clusters)]
                   + 0.05 * rng.standard_normal((clusters, flat.shape[1]))).astype(np.float32)
        mapper = ra.FullSpaceMapperFEATClusterLinStyle_Net(gen.n_latent, 1024, 512, attention_layer=attention_layer,
                                                           cluster_layer=cluster_layer, clusters=clusters,
                                                           cluster_dim=flat.shape[1])
        with torch.no_grad():
            for key, p in mapper.named_parameters():
                if key != "initial_bias":
                    p.copy_(cmo.seeded_value(key, tuple(p.shape)))
            mapper.initial_bias.fill_(bias)
        mapper.store_clusters(torch.from_numpy(centres))
        text = torch.from_numpy(rng.standard_normal((batch, 512)).astype(np.float32))
        x = [torch.cat([text.unsqueeze(1), s[:, :, :, 0, 0]], dim=-1) for s in styles]    # run_attention.py:1240
        with torch.no_grad():
            new_styles, final_map, (loss_delta, loss_reg, loss_tv) = mapper(x, features, size)
        out[f"{name}/text"], out[f"{name}/centres"] = text.numpy(), centres
        out[f"{name}/styles_out"] = torch.stack([s[:, 0, :, 0, 0] for s in new_styles]).numpy()
        out[f"{name}/final"] = final_map.numpy()
        out[f"{name}/losses"] = np.array([float(loss_delta), float(loss_reg), float(loss_tv)], np.float64)
        out[f"{name}/cfg"] = np.array([size, clusters, cluster_layer, attention_layer], np.int64)
        out[f"{name}/bias"] = np.float32(bias)
        out[f"{name}/keys"] = np.array(json.dumps({k: list(v.shape) for k, v in mapper.state_dict().items()}))
        print(name, out[f"{name}/styles_out"].shape, final_map.shape, out[f"{name}/losses"],
              "map above threshold:", float((final_map > 0).float().mean()))
    path = os.path.join(ROOT, "tests", "golden", "cluster_mapper.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
