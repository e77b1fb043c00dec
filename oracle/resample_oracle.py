"""CPU restatement (numpy, fp32) of the reference's CLIP / VGG pre-resample — TEST INFRASTRUCTURE (only tests/,
smoke() and bench.py's cpu_baseline leg may import it).

Follows criteria/clip_loss.py:10-14 literally: `torch.nn.Upsample(scale_factor=U)` (nearest; source index
floor(dst * (1/U)) in fp32, clamped — third-party PyTorch, reference pin 1.7.1) materialised in full, then
`torch.nn.AvgPool2d(P)` (stride P, no padding, floor).  Pinned by tests/golden/resample.npz
(oracle/make_resample_golden.py runs the reference's own CLIPLoss.upsample / .avg_pool modules).
"""
import numpy as np

F32 = np.float32


def upsample_nearest(x, scale):
    h, w = x.shape[-2:]
    inv = F32(1.0 / scale)
    iy = np.minimum(np.floor(np.arange(h * scale, dtype=F32) * inv).astype(np.int64), h - 1)
    ix = np.minimum(np.floor(np.arange(w * scale, dtype=F32) * inv).astype(np.int64), w - 1)
    return x[..., iy, :][..., :, ix]


def avg_pool(x, p):
    h, w = x.shape[-2:]
    oh, ow = h // p, w // p
    v = x[..., :oh * p, :ow * p].reshape(x.shape[:-2] + (oh, p, ow, p))
    return (v.sum(axis=(-3, -1), dtype=np.float64) / (p * p)).astype(F32)


def clip_resample(image, scale=7, pool=32):
    return avg_pool(upsample_nearest(np.asarray(image, F32), scale), pool)


def clip_resample_backward(gy, in_hw, scale=7, pool=32):
    """Adjoint: spread gy / P^2 over each pooling window, then sum the U x U upsampled copies of a source pixel."""
    gy = np.asarray(gy, np.float64)
    h, w = in_hw
    oh, ow = gy.shape[-2:]
    gup = np.zeros(gy.shape[:-2] + (h * scale, w * scale))
    gup[..., :oh * pool, :ow * pool] = np.repeat(np.repeat(gy, pool, axis=-2), pool, axis=-1) / (pool * pool)
    return gup.reshape(gy.shape[:-2] + (h, scale, w, scale)).sum(axis=(-3, -1)).astype(F32)
