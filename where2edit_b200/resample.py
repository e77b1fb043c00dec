"""CLIP / VGG pre-resample: the step immediately after `Generator.forward` in every loss of the reference
(SURVEY.md section 8f rank 2).

    image = self.avg_pool(self.upsample(image))        # criteria/clip_loss.py:14, perceptual_loss.py:16-17,
                                                       # attention/run_attention.py:1163,1259
with `upsample = torch.nn.Upsample(scale_factor=7)` and `avg_pool = torch.nn.AvgPool2d(stylegan_size // 32)`
materialises a [B,3,7168,7168] tensor (617 MB fp32 per image) to produce 224x224.  `clip_resample` computes the
same result in one pass (image read once, fp32), differentiable; `ClipResample` is the module form holding the
reference's two hyper-parameters.  CUDA only (libw2e.so, include/w2e.h: w2e_box_resample_{fwd,bwd}).
"""
import torch

# autocast safety (the reference's --amp wraps mapper + generator in torch.cuda.amp.autocast, run_attention.py:1231):
# the kernels take fp32 (or bf16) pointers, so half-precision tensors handed over by autocast-ed linears are cast to
# fp32 at every custom Function and autocast is off inside it
_amp_fwd = torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_amp_bwd = torch.amp.custom_bwd(device_type="cuda")

from . import _native as N


class _BoxResample(torch.autograd.Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, x, up, pool):
        b, c, h, w = x.shape
        y = torch.empty((b, c, h * up // pool, w * up // pool), device=x.device, dtype=torch.float32)
        if y.numel():
            N.check(N.load().w2e_box_resample_fwd(N.ptr(x), N.ptr(y), b * c, h, w, up, pool, N.stream_ptr()), "box_resample_fwd")
        ctx.cfg = (b, c, h, w, up, pool)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    @_amp_bwd
    def backward(ctx, gy):
        b, c, h, w, up, pool = ctx.cfg
        gy = gy.to(torch.float32).contiguous()
        gx = torch.empty((b, c, h, w), device=gy.device, dtype=torch.float32)
        if gx.numel():
            N.check(N.load().w2e_box_resample_bwd(N.ptr(gy), N.ptr(gx), b * c, h, w, up, pool, N.stream_ptr()), "box_resample_bwd")
        return gx, None, None


def clip_resample(image, scale_factor=7, kernel_size=32):
    """AvgPool2d(kernel_size)(Upsample(scale_factor=scale_factor)(image)) for an integer nearest-neighbour scale
    factor and a square pooling window with stride = kernel_size (the reference's only configuration).
    image [B,C,H,W] -> [B,C, H*scale // kernel, W*scale // kernel] fp32."""
    N.require_cuda(image)
    if image.ndim != 4:
        raise ValueError(f"image must be [B,C,H,W], got {tuple(image.shape)}")
    if int(scale_factor) != scale_factor or scale_factor < 1:
        raise ValueError("clip_resample supports integer nearest-neighbour scale factors (the reference uses 7)")
    up, pool = int(scale_factor), int(kernel_size)
    if pool < 1 or image.shape[2] * up < pool or image.shape[3] * up < pool:
        raise ValueError(f"pooling window {pool} does not fit the {image.shape[2] * up}x{image.shape[3] * up} upsampled image")
    return _BoxResample.apply(image.to(torch.float32).contiguous(), up, pool)


class ClipResample(torch.nn.Module):
    """`CLIPLoss.upsample` + `CLIPLoss.avg_pool` (criteria/clip_loss.py:10-11) as one module."""

    def __init__(self, stylegan_size, scale_factor=7):
        super().__init__()
        self.scale_factor = scale_factor
        self.kernel_size = stylegan_size // 32

    def forward(self, image):
        return clip_resample(image, self.scale_factor, self.kernel_size)


__all__ = ["clip_resample", "ClipResample"]
