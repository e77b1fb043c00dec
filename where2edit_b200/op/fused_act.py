"""fused_leaky_relu / FusedLeakyReLU -- same signature and result as
models/stylegan2/op/fused_act.py:11-39, executed by libw2e (csrc/bias_act.cu) in ONE pass instead
of the reference's add / leaky_relu / mul chain.  No CPU implementation."""
import torch

# autocast safety (the reference's --amp wraps mapper + generator in torch.cuda.amp.autocast, run_attention.py:1231):
# the kernels take fp32 (or bf16) pointers, so half-precision tensors handed over by autocast-ed linears are cast to
# fp32 at every custom Function and autocast is off inside it
_amp_fwd = torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_amp_bwd = torch.amp.custom_bwd(device_type="cuda")
from torch import nn

from .. import _native as N


def _layout(x, bias):
    """(outer, C, inner) view of the reference's broadcasting rule (fused_act.py:24-38): bias sits
    on the LAST dim for 3-D inputs and on dim 1 otherwise."""
    c = bias.shape[0]
    if x.ndim == 3:
        if x.shape[-1] != c:
            raise ValueError(f"fused_leaky_relu: bias of {c} does not match last dim of {tuple(x.shape)}")
        return x.shape[0] * x.shape[1], c, 1
    if x.ndim < 2 or x.shape[1] != c:
        raise ValueError(f"fused_leaky_relu: bias of {c} does not match dim 1 of {tuple(x.shape)}")
    inner = 1
    for d in x.shape[2:]:
        inner *= d
    return x.shape[0], c, inner


def bias_act_forward(x, bias, noise, noise_w, slope, scale):
    """Raw launcher.  noise: [1 or B, 1, H, W] float32 or None; noise_w: 1-element device tensor."""
    outer, c, inner = _layout(x, bias) if bias is not None else (x.shape[0], x.shape[1], x[0, 0].numel())
    y = torch.empty_like(x)
    nb = 0
    if noise is not None:
        nb = noise.shape[0]
        if noise.numel() != nb * inner:
            raise ValueError(f"noise shape {tuple(noise.shape)} does not match activation {tuple(x.shape)}")
    lib = N.load()
    N.check(lib.w2e_bias_act_fwd(N.ptr(x), N.ptr(bias), N.ptr(noise), N.ptr(noise_w), nb, N.ptr(y), outer, c, inner,
                                 float(slope), float(scale), N.dtype_code(x), N.stream_ptr()), "bias_act_fwd")
    return y


def bias_act_backward(gy, y, bias_shape_c, want_gbias, layout, slope, scale):
    outer, c, inner = layout
    gx = torch.empty_like(gy)
    lib = N.load()
    gbias = ws = None
    if want_gbias:
        gbias = torch.empty(c, device=gy.device, dtype=torch.float32)
        ws = torch.empty(max(int(lib.w2e_bias_act_bwd_workspace(outer, c, inner)) // 4, 1), device=gy.device,
                         dtype=torch.float32)
    N.check(lib.w2e_bias_act_bwd(N.ptr(gy), N.ptr(y), N.ptr(gx), N.ptr(gbias), N.ptr(ws), outer, c, inner,
                                 float(slope), float(scale), N.dtype_code(gy), N.stream_ptr()), "bias_act_bwd")
    return gx, gbias


class _BiasActBackward(torch.autograd.Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, gy, y, want_gbias, layout, slope, scale):
        ctx.save_for_backward(y)
        ctx.cfg = (layout, slope, scale)
        gx, gbias = bias_act_backward(gy.contiguous(), y, layout[1], want_gbias, layout, slope, scale)
        if gbias is None:
            gbias = gy.new_zeros(0)
        return gx, gbias

    @staticmethod
    @_amp_bwd
    def backward(ctx, ggx, ggbias):
        # second order: d(gx)/d(gy) is the same mask; the op is piecewise linear in y
        (y,) = ctx.saved_tensors
        layout, slope, scale = ctx.cfg
        gg = None
        if ggx is not None:
            gg, _ = bias_act_backward(ggx.contiguous(), y, layout[1], False, layout, slope, scale)
        return gg, None, None, None, None, None


class _BiasAct(torch.autograd.Function):
    """y = lrelu(x + bias + noise_w*noise) * scale; gradients to x and bias (noise terms are buffers)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, x, bias, noise, noise_w, slope, scale):
        y = bias_act_forward(x, bias, noise, noise_w, slope, scale)
        ctx.save_for_backward(y)
        ctx.layout = _layout(x, bias)
        ctx.cfg = (slope, scale)
        return y

    @staticmethod
    @_amp_bwd
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        slope, scale = ctx.cfg
        gx, gbias = _BiasActBackward.apply(gy, y, ctx.needs_input_grad[1], ctx.layout, slope, scale)
        return gx, (gbias if ctx.needs_input_grad[1] else None), None, None, None, None


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    """models/stylegan2/op/fused_act.py:23-39.  Like the reference (line 25) the input is moved to
    the CUDA device first; everything after that is one kernel."""
    if not input.is_cuda:
        input = input.cuda()
    N.require_cuda(bias)
    if negative_slope < 0 or scale <= 0:
        raise ValueError("fused_leaky_relu: needs negative_slope >= 0 and scale > 0")
    cast = None
    if input.dtype not in (torch.float32, torch.bfloat16):
        cast = input.dtype
        input = input.float()
    x = input.contiguous()
    b = bias.to(torch.float32).contiguous()
    y = _BiasAct.apply(x, b, None, None, float(negative_slope), float(scale))
    return y if cast is None else y.to(cast)


class FusedLeakyReLU(nn.Module):
    """models/stylegan2/op/fused_act.py:11-20 (parameter name `bias`, zeros at init)."""

    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)
