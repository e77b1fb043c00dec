"""Generate tests/golden/mapper.npz with the UNMODIFIED reference LevelsMapper (mapper/latent_mappers.py:47-82)
loaded with oracle.mapper_oracle.mapper_state's seeded weights.  TEST INFRASTRUCTURE; run once in the build
container:  python oracle/make_mapper_golden.py     (shim: torch.Tensor.cuda = identity, fused_act.py:25)
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("W2E_REFERENCE", "/root/reference")

from oracle import mapper_oracle as mo  # noqa: E402


def main():
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self
    from mapper.latent_mappers import LevelsMapper
    out = {}
    torch.manual_seed(6)
    x = torch.randn(2, 18, 512)
    out["x"] = x.numpy()
    for name, flags in (("all", (False, False, False)), ("no_fine", (False, False, True))):
        opts = types.SimpleNamespace(no_coarse_mapper=flags[0], no_medium_mapper=flags[1], no_fine_mapper=flags[2])
        m = LevelsMapper(opts)
        levels = [lv for lv, off in zip(mo.LEVELS, flags) if not off]
        m.load_state_dict({k: torch.from_numpy(v) for k, v in mo.mapper_state(levels=levels).items()}, strict=True)
        xi = x.clone().requires_grad_(True)
        y = m(xi)
        head = torch.randn_like(y)
        (y * head).sum().backward()
        out[f"{name}/y"], out[f"{name}/head"], out[f"{name}/gx"] = y.detach().numpy(), head.numpy(), xi.grad.numpy()
        if name == "all":   # the mapper is what the edit loop trains: pin its parameter gradients (subsampled)
            for pn, p in m.named_parameters():
                out[f"{name}/grad/{pn}"] = p.grad.reshape(-1)[::997].numpy().copy()
                out[f"{name}/gradsum/{pn}"] = np.array([p.grad.double().sum().item(), p.grad.double().abs().sum().item()])
        print(name, tuple(y.shape), float(y.abs().max()), float(xi.grad.abs().max()))
    path = os.path.join(ROOT, "tests", "golden", "mapper.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
