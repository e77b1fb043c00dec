"""Generate tests/golden/resample.npz with the reference's own modules: `CLIPLoss(opts).upsample` and
`.avg_pool` (criteria/clip_loss.py:10-11), applied as in CLIPLoss.forward (:14), plus autograd's gradient.

TEST INFRASTRUCTURE; run once in the build container:  python oracle/make_resample_golden.py
Shim: the `clip` import is replaced by a stand-in whose `load` returns (None, None) — the CLIP model is not
installed and is not on this path.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("W2E_REFERENCE", "/root/reference")

# (image shape, stylegan_size) -> kernel_size = stylegan_size // 32
CASES = {
    "s64": ((1, 2, 64, 64), 64),       # pool 2 -> 224x224
    "s32": ((1, 1, 32, 32), 32),       # pool 1: pure nearest upsample -> 224x224
    "s160": ((1, 2, 40, 40), 160),     # pool 5 (not a power of two) -> 56x56
    "ragged": ((1, 1, 33, 21), 128),   # pool 4: 231x147 -> 57x36, AvgPool drops the remainder
}


def main():
    sys.path.insert(0, REF)
    fake = types.ModuleType("clip")
    fake.load = lambda *a, **k: (None, None)
    sys.modules["clip"] = fake
    from criteria.clip_loss import CLIPLoss
    out = {}
    torch.manual_seed(21)
    for name, (shape, size) in CASES.items():
        loss = CLIPLoss(types.SimpleNamespace(stylegan_size=size))
        x = torch.randn(*shape, requires_grad=True)
        y = loss.avg_pool(loss.upsample(x))
        gy = torch.randn_like(y)
        (y * gy).sum().backward()
        out[f"{name}/x"], out[f"{name}/y"] = x.detach().numpy(), y.detach().numpy()
        out[f"{name}/gy"], out[f"{name}/gx"] = gy.numpy(), x.grad.numpy()
        out[f"{name}/pool"] = np.int64(loss.avg_pool.kernel_size)
        out[f"{name}/scale"] = np.int64(loss.upsample.scale_factor)
        print(name, shape, "->", tuple(y.shape), "pool", loss.avg_pool.kernel_size)
    path = os.path.join(ROOT, "tests", "golden", "resample.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
