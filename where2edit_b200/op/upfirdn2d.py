"""upfirdn2d -- same signature and result as models/stylegan2/op/upfirdn2d.py:11-60, executed by
libw2e's CUDA kernels (csrc/upfirdn2d.cu).  There is no CPU implementation here."""
import weakref

import torch

# autocast safety (the reference's --amp wraps mapper + generator in torch.cuda.amp.autocast, run_attention.py:1231):
# the kernels take fp32 (or bf16) pointers, so half-precision tensors handed over by autocast-ed linears are cast to
# fp32 at every custom Function and autocast is off inside it
_amp_fwd = torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_amp_bwd = torch.amp.custom_bwd(device_type="cuda")

from .. import _native as N

_TAPS_CACHE = {}


def kernel_taps(kernel):
    """Host copy of the (tiny) FIR kernel as a flat tuple; cached per tensor object + version so the
    device->host read happens once (it would be illegal inside CUDA-graph capture)."""
    key = id(kernel)
    hit = _TAPS_CACHE.get(key)
    if hit is not None and hit[0]() is kernel and hit[1] == kernel._version:
        return hit[2]
    if kernel.ndim != 2:
        raise ValueError(f"upfirdn2d: kernel must be 2-D, got shape {tuple(kernel.shape)}")
    taps = tuple(kernel.detach().to("cpu", torch.float32).reshape(-1).tolist())
    if len(_TAPS_CACHE) > 512:
        for k in [k for k, v in _TAPS_CACHE.items() if v[0]() is None]:
            del _TAPS_CACHE[k]
    _TAPS_CACHE[key] = (weakref.ref(kernel), kernel._version, taps)
    return taps


def out_size(in_size, up, down, pad0, pad1, k):
    return (in_size * up + pad0 + pad1 - k) // down + 1


def _run(x, taps, kh, kw, geom, backward=False, in_hw=None):
    """x: [N,C,H,W] contiguous CUDA fp32/bf16.  geom = (up_x, up_y, down_x, down_y, px0, px1, py0, py1)."""
    up_x, up_y, down_x, down_y, px0, px1, py0, py1 = geom
    n, c = x.shape[0], x.shape[1]
    lib = N.load()
    host = N.host_floats(taps)
    if not backward:
        in_h, in_w = x.shape[2], x.shape[3]
        oh = out_size(in_h, up_y, down_y, py0, py1, kh)
        ow = out_size(in_w, up_x, down_x, px0, px1, kw)
        if oh <= 0 or ow <= 0:
            raise ValueError(f"upfirdn2d: empty output {oh}x{ow}")
        y = torch.empty((n, c, oh, ow), device=x.device, dtype=x.dtype)
        N.check(lib.w2e_upfirdn2d_fwd(N.ptr(x), N.ptr(y), host, n * c, in_h, in_w, kh, kw, up_x, up_y, down_x,
                                      down_y, px0, px1, py0, py1, N.dtype_code(x), N.stream_ptr()), "upfirdn2d_fwd")
        return y
    in_h, in_w = in_hw
    gx = torch.empty((n, c, in_h, in_w), device=x.device, dtype=x.dtype)
    N.check(lib.w2e_upfirdn2d_bwd(N.ptr(x), N.ptr(gx), host, n * c, in_h, in_w, kh, kw, up_x, up_y, down_x, down_y,
                                  px0, px1, py0, py1, N.dtype_code(x), N.stream_ptr()), "upfirdn2d_bwd")
    return gx


class _UpFirDn2dBackward(torch.autograd.Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, gy, taps, kh, kw, geom, in_hw):
        ctx.cfg = (taps, kh, kw, geom)
        return _run(gy.contiguous(), taps, kh, kw, geom, backward=True, in_hw=in_hw)

    @staticmethod
    @_amp_bwd
    def backward(ctx, ggx):
        taps, kh, kw, geom = ctx.cfg
        # the operator is linear: the gradient of its adjoint is the operator itself
        return _UpFirDn2d.apply(ggx, taps, kh, kw, geom), None, None, None, None, None


class _UpFirDn2d(torch.autograd.Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, x, taps, kh, kw, geom):
        ctx.cfg = (taps, kh, kw, geom, (x.shape[2], x.shape[3]))
        return _run(x, taps, kh, kw, geom)

    @staticmethod
    @_amp_bwd
    def backward(ctx, gy):
        taps, kh, kw, geom, in_hw = ctx.cfg
        return _UpFirDn2dBackward.apply(gy, taps, kh, kw, geom, in_hw), None, None, None, None


def _prepare(x):
    N.require_cuda(x)
    if x.ndim != 4:
        raise ValueError(f"upfirdn2d: input must be [N,C,H,W], got {tuple(x.shape)}")
    cast = None
    if x.dtype not in (torch.float32, torch.bfloat16):
        cast = x.dtype
        x = x.float()
    return x.contiguous(), cast


def upfirdn2d_native(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """models/stylegan2/op/upfirdn2d.py:19-60 (same ten geometry arguments)."""
    x, cast = _prepare(input)
    taps = kernel_taps(kernel)
    kh, kw = kernel.shape
    geom = (int(up_x), int(up_y), int(down_x), int(down_y), int(pad_x0), int(pad_x1), int(pad_y0), int(pad_y1))
    y = _UpFirDn2d.apply(x, taps, int(kh), int(kw), geom)
    return y if cast is None else y.to(cast)


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """models/stylegan2/op/upfirdn2d.py:11-16."""
    return upfirdn2d_native(input, kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
