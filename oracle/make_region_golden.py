"""Generate tests/golden/region.npz by executing the UNMODIFIED reference region-mask code:
`FullSpaceMapperFEATClusterLinStyle_Net.forward` (attention/run_attention.py:755-893) on the captured
features of the reference generator, and `pairwise_distance` (utils.py:244-263).

TEST INFRASTRUCTURE.  Run once in the build container (the reference tree does not exist on the GPU box):
    python oracle/make_region_golden.py

The region-mask construction is inline code of that forward (cluster assignment :775-794, per-cluster mean
attention + threshold + gaussian blur :847-884), so the pin is taken at its observable edges:
  * inputs : the feature the clusters are assigned on, the cluster centres, and the logits leaving
             `attention_last` (forward hook) together with `initial_bias`;
  * outputs: `pairwise_distance`'s result (wrapped, not modified), the returned `final_attention_map`,
             `loss_reg`, `loss_tv`, and autograd's gradient at the hooked logits for a seeded linear loss head.
Shims: `torch.Tensor.cuda = identity` (fused_act.py:25, no GPU here) and empty stand-ins for the `clip` and
`torch_fidelity` imports of utils.py / run_attention.py (not installed; never executed on this path).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("W2E_REFERENCE", "/root/reference")


class _Absent(types.ModuleType):
    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        return lambda *a, **k: None


def import_reference():
    sys.path[:0] = [REF, os.path.join(REF, "attention")]
    torch.Tensor.cuda = lambda self, *a, **k: self
    for name in ("clip", "torch_fidelity"):
        sys.modules.setdefault(name, _Absent(name))
    import attention_model
    import run_attention
    return attention_model, run_attention


def run_case(am, ra, size, clusters, cluster_layer, attention_layer, bias, seed):
    """One forward+backward of the reference mapper; returns the arrays described in the module docstring."""
    torch.manual_seed(seed)
    batch = 2
    g = am.Generator(32, 512, 2, channel_multiplier=2).eval()
    w = torch.randn(batch, g.n_latent, 512)
    with torch.no_grad():
        _, _, styles, features = g([w], input_is_latent=True, randomize_noise=False, return_features=True)
        features = list(features)
        features.append(g.input.input.repeat(batch, 1, 1, 1))          # run_attention.py:1110
    blend = features[cluster_layer - 1]
    cdim = blend.shape[1] + 2 * (blend.shape[1] // 16)
    mapper = ra.FullSpaceMapperFEATClusterLinStyle_Net(g.n_latent, 1024, 512, attention_layer=attention_layer,
                                                       cluster_layer=cluster_layer, clusters=clusters, cluster_dim=cdim)
    # cluster centres = features of random pixels (+ their positions) plus noise: every cluster gets members
    h = blend.shape[-1]
    pc = blend.shape[1] // 16
    pos = torch.arange(h).float() * 2 / float(h - 1) - 1
    rows = []
    for k in range(clusters):
        b, y, x = int(torch.randint(batch, ())), int(torch.randint(h, ())), int(torch.randint(h, ()))
        rows.append(torch.cat([blend[b, :, y, x], pos[x].repeat(pc), pos[y].repeat(pc)]))
    centres = torch.stack(rows) + 0.05 * torch.randn(clusters, cdim)
    mapper.store_clusters(centres)
    with torch.no_grad():
        mapper.initial_bias.fill_(bias)

    seen = {}
    real_pd = ra.pairwise_distance

    def spy(a, b=None):
        d = real_pd(a, b)
        seen["dis"] = d.detach().clone()
        return d
    ra.pairwise_distance = spy

    def hook(mod, inp, out):
        out[0].retain_grad()
        seen["logits"] = out[0]
    handle = mapper.attention_last.register_forward_hook(hook)

    text = torch.randn(batch, 512)
    x = [torch.cat([text.unsqueeze(1), s[:, :, :, 0, 0]], dim=-1) for s in styles]   # run_attention.py:1240
    _, final_map, (loss_delta, loss_reg, loss_tv) = mapper(x, features, size)
    handle.remove()
    ra.pairwise_distance = real_pd

    head = torch.randn_like(final_map)
    total = (final_map * head).sum() + 0.7 * loss_reg.sum() + 1.3 * loss_tv
    total.backward()
    logits = seen["logits"]
    each = torch.sigmoid(logits.detach() + bias).view(batch, size, size)
    ids = seen["dis"].argmin(dim=1).view(batch, h, h)
    return {
        "feature": blend.numpy(), "centres": centres.numpy(), "dis": seen["dis"].numpy(),
        "ids_lowres": ids.numpy().astype(np.int64),
        "logits": logits.detach().view(batch, size, size).numpy(), "bias": np.float32(bias),
        "each": each.numpy(), "final": final_map.detach().numpy(),
        "loss_reg": loss_reg.detach().numpy(), "loss_tv": loss_tv.detach().numpy(),
        "head": head.numpy(), "g_logits": logits.grad.view(batch, size, size).numpy(),
        "size": np.int64(size), "clusters": np.int64(clusters),
    }


CASES = {
    # name: (size, clusters, cluster_layer, attention_layer, initial_bias, seed)
    "same_res": (16, 6, 7, 7, 1.3, 11),     # clusters assigned at 16x16, attention at 16x16
    "upsampled": (32, 9, 7, 10, 1.0, 12),   # clusters at 16x16 nearest-resized to the 32x32 attention map
}


def main():
    am, ra = import_reference()
    out = {}
    for name, cfg in CASES.items():
        res = run_case(am, ra, *cfg)
        means_hi = float((res["final"] > 0).mean())
        print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in res.items()},
              "fraction of the map above the threshold:", means_hi, "loss_reg", res["loss_reg"], "loss_tv", res["loss_tv"])
        for k, v in res.items():
            out[f"{name}/{k}"] = v
    # function-level pin of pairwise_distance (utils.py:244-263)
    torch.manual_seed(5)
    a, b = torch.randn(37, 24), torch.randn(5, 24)
    out["pd/a"], out["pd/b"], out["pd/dis"] = a.numpy(), b.numpy(), ra.pairwise_distance(a, b).numpy()
    path = os.path.join(ROOT, "tests", "golden", "region.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
