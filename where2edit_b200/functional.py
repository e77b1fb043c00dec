"""Differentiable building blocks of the fp32 ("exact") synthesis path, each a thin
torch.autograd.Function around libw2e kernels.  Math: SURVEY.md appendix C (shared-weight form of
models/stylegan2/model.py:239-274 and its backward).  No CPU path.
"""
import torch

# autocast safety (the reference's --amp wraps mapper + generator in torch.cuda.amp.autocast, run_attention.py:1231):
# the kernels take fp32 (or bf16) pointers, so half-precision tensors handed over by autocast-ed linears are cast to
# fp32 at every custom Function and autocast is off inside it
_amp_fwd = torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_amp_bwd = torch.amp.custom_bwd(device_type="cuda")

from . import _native as N
from .op.fused_act import _BiasAct
from .op.upfirdn2d import _UpFirDn2d, kernel_taps


# --------------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------------
def round_tf32(t):
    """fp32 -> nearest tf32-representable fp32 (ties away from zero, like cvt.rna.tf32.f32): the tf32 MMA truncates
    its operands, so they are rounded once when they are produced."""
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def quantize_u8(img):
    """fp32 image in [-1, 1] -> uint8 exactly as torchvision.utils.save_image(normalize=True, range=(-1, 1)) quantises the
    reference's results (run_attention.py:1470, 1535): clamp, (v + 1) / 2, * 255, + 0.5, clamp, truncate.  The last
    layer's epilogue does this in registers (W2E_U8); this is the same arithmetic for images that did not come out of it."""
    t = img.to(torch.float32).clamp(-1.0, 1.0).add(1.0).mul(0.5)
    return t.mul(255.0).add(0.5).clamp(0.0, 255.0).to(torch.uint8)


class PackedWeight:
    """Batch-shared layouts of one ModulatedConv2d weight [1,Cout,Cin,k,k], pre-multiplied by the
    equalised-lr scale 1/sqrt(Cin*k*k) (model.py:216-217):
        fwd  [k*k, Cin, Cout]  engine layout for the forward convolution
        dgr  [k*k, Cout, Cin]  engine layout for dgrad (roles of Cin/Cout swapped)
        wsq  [Cout, Cin]       sum over taps of (scale*W)^2 -> demodulation without per-sample weights
    """

    def __init__(self, weight, scale, key):
        self.key = key
        w = weight[0].to(torch.float32) * scale          # [Cout,Cin,k,k]
        cout, cin, k, _ = w.shape
        self.cin, self.cout, self.k = cin, cout, k
        self.fwd = w.permute(2, 3, 1, 0).reshape(k * k, cin, cout).contiguous()
        self.dgr = w.permute(2, 3, 0, 1).reshape(k * k, cout, cin).contiguous()
        self.wsq = w.pow(2).sum((2, 3)).contiguous()
        self.rgb = w.reshape(cout, cin).contiguous() if k == 1 else None
        self.tc = None  # bf16 tensor-core layout, filled lazily by the engine
        self._tc_dgrad = None

    def tc_fwd(self, dtype=torch.bfloat16):
        """[k*k][Cout][Cin] (equalised-lr scale folded in): operand of the tcgen05 forward convolution, bf16 or
        (tf32 mode) fp32 rounded to tf32."""
        if dtype == torch.float32:
            if getattr(self, "_tc32", None) is None:
                self._tc32 = round_tf32(self.dgr)
            return self._tc32
        if self.tc is None:
            self.tc = self.dgr.to(torch.bfloat16).contiguous()
        return self.tc

    @staticmethod
    def pair_operand(w):
        """[9][N][K] fp32 (tap = ky*3 + kx) -> bf16 [9][2N][2K]: the 3x3 kernel over horizontally adjacent PIXEL PAIRS --
        pair tap (ky, dj) connects input pixel b of pair i+dj to output pixel a of pair i through kx = 2*dj + b - a + 1."""
        n, k = w.shape[1], w.shape[2]
        wp = torch.zeros((9, 2 * n, 2 * k), device=w.device, dtype=torch.float32)
        for ky in range(3):
            for dj in (-1, 0, 1):
                for a in (0, 1):
                    for b in (0, 1):
                        kx = 2 * dj + b - a + 1
                        if 0 <= kx <= 2:
                            wp[ky * 3 + dj + 1, a * n:(a + 1) * n, b * k:(b + 1) * k] = w[ky * 3 + kx]
        return wp.to(torch.bfloat16).contiguous()

    def tc_pair(self):
        """Pair operand of the forward convolution (w2e_modconv_tc2_rgb_pair)."""
        if getattr(self, "_tc_pair", None) is None:
            self._tc_pair = self.pair_operand(self.dgr)    # [9][Cout][Cin]
        return self._tc_pair

    def tc_dgrad_pair(self):
        """Pair operand of the dgrad (flipped taps, channel roles swapped; w2e_modconv_tc2_pair)."""
        if getattr(self, "_tc_dgrad_pair", None) is None:
            self._tc_dgrad_pair = self.pair_operand(self.fwd.flip(0))   # [9][Cin][Cout]
        return self._tc_dgrad_pair

    def tc_dgrad(self, dtype=torch.bfloat16):
        """[k*k][Cin][Cout] with the taps flipped: the dgrad of a same-padded 3x3 convolution is the same
        convolution kernel run on the upstream gradient with input/output channels swapped."""
        if dtype == torch.float32:
            if getattr(self, "_tc_dgrad32", None) is None:
                self._tc_dgrad32 = round_tf32(self.fwd.flip(0))
            return self._tc_dgrad32
        if self._tc_dgrad is None:
            self._tc_dgrad = self.fwd.flip(0).to(torch.bfloat16).contiguous()
        return self._tc_dgrad

    def tc_dgrad_up_fused(self):
        """bf16 [9][Cin][Cout] (tap = ky*3 + kx of the forward kernel): operand of w2e_modconv_tc2_dgrad_up, the one-launch
        dgrad of the transposed (x2) convolution -- gx[j,i] = sum gz[2j+ky, 2i+kx] . W[ky,kx]^T."""
        if getattr(self, "_tc_dgrad_up_fused", None) is None:
            self._tc_dgrad_up_fused = self.fwd.to(torch.bfloat16).contiguous()
        return self._tc_dgrad_up_fused

    def tc_dgrad_up_k(self):
        """bf16 [9][Cin][4 * Cout]: the four per-class weights of `tc_dgrad_up` concatenated along K in the class order
        (0,0), (0,1), (1,0), (1,1) -- operand of w2e_modconv_tc2_dgrad_up_k (one accumulator, K loop over the classes)."""
        if getattr(self, "_tc_dgrad_up_k", None) is None:
            per = self.tc_dgrad_up()
            self._tc_dgrad_up_k = torch.cat([per[(0, 0)], per[(0, 1)], per[(1, 0)], per[(1, 1)]], dim=2).contiguous()
        return self._tc_dgrad_up_k

    def tc_dgrad_up(self, dtype=torch.bfloat16):
        """{(py, px): bf16 [9][Cin][Cout]} -- dgrad of the transposed (x2) convolution as four plain convolutions,
        one per output-parity class of the upstream gradient: class (py, px) holds gz[2j+py, 2i+px] and contributes
        gx[j, i] += gz_class[j+a, i+b] . W[ky, kx]^T with (a, ky) in {(0,0), (1,2)} for py = 0 and {(0,1)} for
        py = 1 (same for b, kx, px).  Offset (+a, +b) is tap (a+1, b+1) of the 3x3 kernel; the other taps are zero."""
        cache = "_tc_dgrad_up" + ("32" if dtype == torch.float32 else "")
        if getattr(self, cache, None) is None:
            axis = {0: [(0, 0), (1, 2)], 1: [(0, 1)]}
            out = {}
            for py in (0, 1):
                for px in (0, 1):
                    w = torch.zeros_like(self.fwd)
                    for a, ky in axis[py]:
                        for b, kx in axis[px]:
                            w[(a + 1) * 3 + (b + 1)] = self.fwd[ky * 3 + kx]
                    out[(py, px)] = round_tf32(w) if dtype == torch.float32 else w.to(dtype).contiguous()
            setattr(self, cache, out)
        return getattr(self, cache)


def demod_coefficients(s, wsq):
    """d[b,o] = rsqrt(sum_i s[b,i]^2 * wsq[o,i] + 1e-8) (model.py:242).  Uses libw2e's kernel when no
    gradient is needed, plain torch ops (tiny [B,C] algebra) when autograd must see it."""
    if torch.is_grad_enabled() and s.requires_grad:
        with torch.autocast("cuda", enabled=False):   # fp32 whatever the caller's autocast state
            s32 = s.to(torch.float32)
            return torch.rsqrt((s32 * s32) @ wsq.t() + 1e-8)
    s = s.to(torch.float32).contiguous()
    d = torch.empty((s.shape[0], wsq.shape[0]), device=s.device, dtype=torch.float32)
    N.check(N.load().w2e_style_demod(N.ptr(s), N.ptr(wsq), N.ptr(d), s.shape[0], s.shape[1], wsq.shape[0],
                                     N.stream_ptr()), "style_demod")
    return d


# --------------------------------------------------------------------------------------------
# convolution engine launches
# --------------------------------------------------------------------------------------------
def _taps_plain(k, flip=False):
    taps = []
    for ky in range(k):
        for kx in range(k):
            dy, dx = ky - k // 2, kx - k // 2
            taps += [(-dy if flip else dy), (-dx if flip else dx), ky * k + kx]
    return taps


# conv_transpose2d(stride 2, k=3): output parity class p uses taps ky in {0,2} (input offsets 0,-1)
# when p == 0 and ky == 1 (offset 0) when p == 1.
_UP_AXIS = {0: [(0, 0), (-1, 2)], 1: [(0, 1)]}


def _engine(x, w, in_scale, out_scale, y, cin, cout, in_hw, out_hw, grid_hw, in_stride, out_stride, py, px, taps,
            bias=None, noise=None, noise_w=None, act=N.ACT_NONE):
    nb = 0 if noise is None else noise.shape[0]
    N.check(N.load().w2e_conv_engine_f32(
        N.ptr(x), N.ptr(w), N.ptr(in_scale), N.ptr(out_scale), N.ptr(bias), N.ptr(noise), N.ptr(noise_w), nb,
        N.ptr(y), x.shape[0], cin, cout, in_hw[0], in_hw[1], out_hw[0], out_hw[1], grid_hw[0], grid_hw[1],
        in_stride, out_stride, py, px, N.host_ints(taps), len(taps) // 3, act, 0, N.stream_ptr()), "conv_engine")


def conv_forward(x, s, d, pw, k, upsample, bias=None, noise=None, noise_w=None, act=N.ACT_NONE):
    """y = d * conv(x*s, W) [ + fused noise/bias/act ].  Up: (2H+1)x(2W+1) pre-blur tensor."""
    b, cin, h, w = x.shape
    if not upsample:
        y = torch.empty((b, pw.cout, h, w), device=x.device, dtype=torch.float32)
        _engine(x, pw.fwd, s, d, y, cin, pw.cout, (h, w), (h, w), (h, w), 1, 1, 0, 0, _taps_plain(k), bias, noise,
                noise_w, act)
        return y
    oh, ow = 2 * h + 1, 2 * w + 1
    y = torch.empty((b, pw.cout, oh, ow), device=x.device, dtype=torch.float32)
    for py in (0, 1):
        for px in (0, 1):
            taps = []
            for dy, ky in _UP_AXIS[py]:
                for dx, kx in _UP_AXIS[px]:
                    taps += [dy, dx, ky * 3 + kx]
            _engine(x, pw.fwd, s, d, y, cin, pw.cout, (h, w), (oh, ow), (h + 1 - py, w + 1 - px), 1, 2, py, px, taps)
    return y


def conv_dgrad(gy, d, pw, k, upsample, in_hw):
    """gxs = conv^T(gy*d, W): gradient w.r.t. the MODULATED input x*s."""
    b = gy.shape[0]
    h, w = in_hw
    gxs = torch.empty((b, pw.cin, h, w), device=gy.device, dtype=torch.float32)
    if not upsample:
        _engine(gy, pw.dgr, d, None, gxs, pw.cout, pw.cin, (h, w), (h, w), (h, w), 1, 1, 0, 0, _taps_plain(k, flip=True))
    else:
        taps = []
        for ky in range(3):
            for kx in range(3):
                taps += [ky, kx, ky * 3 + kx]
        _engine(gy, pw.dgr, d, None, gxs, pw.cout, pw.cin, (gy.shape[2], gy.shape[3]), (h, w), (h, w), 2, 1, 0, 0, taps)
    return gxs


# --------------------------------------------------------------------------------------------
# tensor-core (bf16 operands, fp32 accumulate) variants of the two convolutions of the autograd path.
# Used when the generator's precision is "bf16" and gradients are required (CLIP/ID-loss latent edits,
# attention/run_attention.py:1419): the forward and the dgrad of every 3x3 ModulatedConv2d run on the
# tcgen05 kernel (csrc/modconv_tc2.cu) between two layout passes; everything else of the backward
# (activation, blur, ToRGB, style reductions) stays on the fp32 kernels.
# --------------------------------------------------------------------------------------------
# set by Generator.forward for the duration of a tensor-core-precision forward through the module path:
# False, "bf16" (bf16 operands) or "tf32" (fp32 tensors read as tf32 by the MMA; north star item 1)
TC_AUTOGRAD = False
_TC_ERR = {}


def _tc_error_flag(dev):
    f = _TC_ERR.get(dev)
    if f is None:
        f = _TC_ERR[dev] = N.ErrorFlag("libw2e modconv_tc2, autograd path")
    return f.tensor(dev)


def tc_poll():
    """Non-synchronising check (start of every bf16-precision autograd forward): raises if a tensor-core kernel
    of an EARLIER forward/backward timed out."""
    for f in _TC_ERR.values():
        f.poll()


def tc_publish():
    for f in _TC_ERR.values():
        f.publish()


def tc_assert_ok():
    """Synchronising check of the pipeline-timeout flag of the tensor-core kernels used by autograd."""
    for f in _TC_ERR.values():
        f.check()


def _tc_dtype(mode):
    return torch.float32 if mode == "tf32" else torch.bfloat16


def _nhwc_mod(x, scale, dtype=torch.bfloat16):
    """fp32 NCHW -> channels-last `dtype` (bf16, or fp32 in tf32 mode), times scale[b, c] (None = 1)."""
    b, c, h, w = x.shape
    y = torch.empty((b, h, w, c), device=x.device, dtype=dtype)
    N.check(N.load().w2e_nchw_to_nhwc_mod(N.ptr(x), N.ptr(scale), N.ptr(y), b, b, c, h * w, N.dtype_code(y),
                                          N.stream_ptr()), "nchw_to_nhwc_mod")
    return y


def _nchw_f32(x):
    b, h, w, c = x.shape
    y = torch.empty((b, c, h, w), device=x.device, dtype=torch.float32)
    N.check(N.load().w2e_nhwc_to_nchw_f32(N.ptr(x), N.ptr(y), b, c, h * w, N.dtype_code(x), N.stream_ptr()),
            "nhwc_to_nchw_f32")
    return y


def tc_supported(cin, cout, k, mode="bf16"):
    return (k == 3 and cin % (16 if mode == "tf32" else 32) == 0 and cout % 16 == 0
            and bool(N.load().w2e_modconv_tc_supported()))


def _tc_conv(xs, w, out_scale, cin, cout, transposed):
    """w2e_modconv_tc2 (bf16 tensors) or w2e_modconv_tc2_tf32 (fp32 tensors), by the dtype of xs."""
    b, h, wd, _ = xs.shape
    oh, ow = (2 * h + 1, 2 * wd + 1) if transposed else (h, wd)
    y = torch.empty((b, oh, ow, cout), device=xs.device, dtype=xs.dtype)
    tf32 = xs.dtype == torch.float32
    if w.dtype != xs.dtype:
        raise RuntimeError("where2edit_b200: tensor-core convolution operands must share one dtype")
    N.note(kind="modconv", flops=2.0 * 9 * cin * cout * b * h * wd,
           tag=f"{'up ' if transposed else ''}{cin}->{cout}@{h}x{wd} ({'tf32' if tf32 else 'bf16'}, module path)")
    fn = N.load().w2e_modconv_tc2_tf32 if tf32 else N.load().w2e_modconv_tc2
    N.check(fn(N.ptr(xs), N.ptr(w), N.ptr(out_scale), None, None, None, 0, None, N.ptr(y), None,
               N.ptr(_tc_error_flag(xs.device)), b, cin, cout, h, wd, int(transposed), N.ACT_NONE, None, N.stream_ptr()),
            "modconv_tc2_tf32" if tf32 else "modconv_tc2")
    return y


def conv_forward_tc(x, s, d, pw, upsample, mode="bf16"):
    """conv_forward on the tensor cores: y = d * conv(x*s, W) with bf16 (or tf32) operands, fp32 accumulate."""
    dt = _tc_dtype(mode)
    return _nchw_f32(_tc_conv(_nhwc_mod(x, s, dt), pw.tc_fwd(dt), d, pw.cin, pw.cout, upsample))


def conv_dgrad_tc(gy, d, pw, upsample=False, in_hw=None, mode="bf16"):
    """conv_dgrad on the tensor cores.  Plain (same-padded) 3x3 convolution: one launch with flipped taps.
    Transposed x2 convolution: its dgrad is a stride-2 convolution of the (2h+1)^2 upstream gradient, run as four
    plain convolutions over the gradient's output-parity classes (PackedWeight.tc_dgrad_up) and summed."""
    dt = _tc_dtype(mode)
    if not upsample:
        return _nchw_f32(_tc_conv(_nhwc_mod(gy, d, dt), pw.tc_dgrad(dt), None, pw.cout, pw.cin, False))
    h, w = in_hw
    b, cout, zh, zw = gy.shape
    lib = N.load()
    ys = []
    for py in (0, 1):
        for px in (0, 1):
            hc, wc = (zh - py + 1) // 2, (zw - px + 1) // 2
            g_c = torch.empty((b, hc, wc, cout), device=gy.device, dtype=dt)
            N.check(lib.w2e_nchw_class_to_nhwc_mod(N.ptr(gy), N.ptr(d), N.ptr(g_c), b, cout, zh, zw, py, px,
                                                   N.dtype_code(g_c), N.stream_ptr()), "nchw_class_to_nhwc_mod")
            ys.append(_tc_conv(g_c, pw.tc_dgrad_up(dt)[(py, px)], None, pw.cout, pw.cin, False))
    gxs = torch.empty((b, pw.cin, h, w), device=gy.device, dtype=torch.float32)
    N.check(lib.w2e_nhwc_sum4_to_nchw_f32(N.ptr(ys[0]), N.ptr(ys[1]), N.ptr(ys[2]), N.ptr(ys[3]), N.ptr(gxs), b, pw.cin,
                                          h, w, N.dtype_code(ys[0]), N.stream_ptr()), "nhwc_sum4_to_nchw_f32")
    return gxs


def _rowdot(a, b, scale=None, want_prod=False):
    rows = a.shape[0] * a.shape[1]
    inner = a[0, 0].numel()
    dot = torch.empty((a.shape[0], a.shape[1]), device=a.device, dtype=torch.float32)
    prod = torch.empty_like(a) if want_prod else None
    lib = N.load()
    nseg = int(lib.w2e_rowdot_segments(rows, inner)) if rows <= 65535 else 1
    if nseg > 1:   # long rows: segmented, deterministic two-pass reduction
        partial = torch.empty((rows, nseg), device=a.device, dtype=torch.float32)
        N.check(lib.w2e_rowdot_seg_f32(N.ptr(a), N.ptr(b), N.ptr(scale), N.ptr(prod), N.ptr(dot), N.ptr(partial), rows,
                                       inner, nseg, N.stream_ptr()), "rowdot_seg")
    else:
        N.check(lib.w2e_rowdot_f32(N.ptr(a), N.ptr(b), N.ptr(scale), N.ptr(prod), N.ptr(dot), rows, inner,
                                   N.stream_ptr()), "rowdot")
    return dot, prod


class _ModConv(torch.autograd.Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, x, s, d, pw, k, upsample):
        mode = TC_AUTOGRAD
        tc = bool(mode) and tc_supported(pw.cin, pw.cout, k, mode)
        y = conv_forward_tc(x, s, d, pw, upsample, mode) if tc else conv_forward(x, s, d, pw, k, upsample)
        ctx.save_for_backward(x, s, d if d is not None else x.new_zeros(0), y)
        ctx.cfg = (pw, k, upsample, d is not None, mode if (tc and tc_supported(pw.cout, pw.cin, k, mode)) else False)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    @_amp_bwd
    def backward(ctx, gy):
        x, s, d, y = ctx.saved_tensors
        pw, k, upsample, has_d, tc_dgrad = ctx.cfg
        gy = gy.contiguous()
        d_or_none = d if has_d else None
        gx = gs = gd = None
        if has_d and ctx.needs_input_grad[2]:
            dot, _ = _rowdot(gy, y)                 # sum_p gy*y = d * sum_p gy*z
            gd = dot / d
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            if tc_dgrad:
                gxs = conv_dgrad_tc(gy, d_or_none, pw, upsample, (x.shape[2], x.shape[3]), tc_dgrad)
            else:
                gxs = conv_dgrad(gy, d_or_none, pw, k, upsample, (x.shape[2], x.shape[3]))
            gs, gx = _rowdot(gxs, x, scale=s.contiguous(), want_prod=True)   # gs = sum_p gxs*x ; gx = gxs*s
        return gx, gs, gd, None, None, None


def modulated_conv2d(x, s, d, pw, k, upsample):
    """x [B,Cin,H,W] fp32, s [B,Cin], d [B,Cout] or None -> [B,Cout,H,W] (or (2H+1)^2 pre-blur)."""
    x = x.to(torch.float32).contiguous()
    s = s.to(torch.float32).contiguous()     # (autocast hands over half-precision styles)
    if d is not None:
        d = d.to(torch.float32).contiguous()
    if torch.is_grad_enabled() and (x.requires_grad or s.requires_grad or (d is not None and d.requires_grad)):
        return _ModConv.apply(x, s, d, pw, k, upsample)
    if TC_AUTOGRAD and tc_supported(pw.cin, pw.cout, k, TC_AUTOGRAD):
        return conv_forward_tc(x, s, d, pw, upsample, TC_AUTOGRAD)
    return conv_forward(x, s, d, pw, k, upsample)


# --------------------------------------------------------------------------------------------
# convolution-weight gradient (opt-in: trainable ModulatedConv2d outside the frozen generator)
# --------------------------------------------------------------------------------------------
def modconv_weight_grad(x, s, d, y, gy, weight, scale, k, upsample):
    """dL/dweight [1,Cout,Cin,k,k] of  y = modulated_conv2d(x, s, d, weight)  (models/stylegan2/model.py:240-274)
    for gy = dL/dy, y being the convolution output BEFORE the up path's Blur.  With W~ = scale * weight:

        y[b,o] = d[b,o] * conv(x[b] * s[b], W~)[o],      d[b,o] = rsqrt(sum_{c,t} (W~[o,c,t] s[b,c])^2 + eps)

      direct term   conv-weight-gradient of (x*s) against (gy*d), summed over the batch: ONE library
                    (cuDNN) weight-gradient call -- the reference's per-sample [B*Cout,Cin,k,k] weights never exist;
                    for the transposed x2 convolution the roles of input and output swap (stride-2 correlation of
                    gy with x*s);
      demod term    dL/dd[b,o] = sum_p gy*y / d   and   dd/dW~ = -d^3 s^2 W~   ->   -W~ * ((dL/dd * d^3)^T @ s^2).

    Plain torch ops (library GEMM / convolution): used only for modules whose weight is being trained -- the
    cluster-style mapper's 1x1 StyledConv attention heads (attention/run_attention.py:725-735); the synthesis
    path keeps the generator frozen and never calls it."""
    cout, cin = weight.shape[1], weight.shape[2]
    dt = torch.float64 if x.dtype == torch.float64 else torch.float32   # fp32 on the GPU path; fp64 for the CPU check
    w = weight[0].to(dt) * scale
    s = s.to(dt)
    xm = x.to(dt) * s[:, :, None, None]
    gy = gy.to(dt)
    gyd = gy if d is None else gy * d[:, :, None, None]
    if upsample:
        gw = torch.nn.grad.conv2d_weight(gyd, (cin, cout, k, k), xm, stride=2).transpose(0, 1)
    else:
        gw = torch.nn.grad.conv2d_weight(xm, (cout, cin, k, k), gyd, padding=k // 2)
    if d is not None:
        gd = (gy * y).sum((2, 3)) / d
        gw = gw - w * ((gd * d.pow(3)).t() @ s.square())[:, :, None, None]
    return (gw * scale).unsqueeze(0).to(weight.dtype)


class _WeightGradTap(torch.autograd.Function):
    """Identity on the convolution output that hands `modconv_weight_grad` to autograd as the gradient of
    `weight` (the convolution kernels themselves take the weight as a packed, detached constant)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, y, weight, x, s, d, scale, k, upsample):
        ctx.save_for_backward(y, weight, x, s, d if d is not None else y.new_zeros(0))
        ctx.cfg = (scale, k, upsample, d is not None)
        return y.view_as(y)

    @staticmethod
    @torch.autograd.function.once_differentiable
    @_amp_bwd
    def backward(ctx, gy):
        y, weight, x, s, d = ctx.saved_tensors
        scale, k, upsample, has_d = ctx.cfg
        gw = None
        if ctx.needs_input_grad[1]:
            gw = modconv_weight_grad(x, s, d if has_d else None, y, gy.contiguous(), weight, scale, k, upsample)
        return gy, gw, None, None, None, None, None, None


def weight_grad_tap(y, weight, x, s, d, scale, k, upsample):
    return _WeightGradTap.apply(y, weight, x.detach(), s.detach(), None if d is None else d.detach(), float(scale),
                                int(k), bool(upsample))


# --------------------------------------------------------------------------------------------
# noise + bias + activation
# --------------------------------------------------------------------------------------------
def noise_bias_act(x, bias, noise, noise_w, slope, scale):
    """NoiseInjection + FusedLeakyReLU in one pass (model.py:279-290 + op/fused_act.py:23-39)."""
    noise = noise.to(torch.float32).contiguous()
    return _BiasAct.apply(x.contiguous(), bias.detach().to(torch.float32).contiguous(), noise,
                          noise_w.detach().to(torch.float32).contiguous(), float(slope), float(scale))


# --------------------------------------------------------------------------------------------
# ToRGB
# --------------------------------------------------------------------------------------------
def separable_taps(kernel2d):
    """1-D factor of a separable 4x4 FIR kernel given as a flat tuple (row-major)."""
    n = int(round(len(kernel2d) ** 0.5))
    rows = [kernel2d[i * n:(i + 1) * n] for i in range(n)]
    total = sum(kernel2d)
    col = [sum(r) for r in rows]
    row = [sum(rows[i][j] for i in range(n)) for j in range(n)]
    # k = outer(col, row) / total  for a rank-1 kernel; split the gain evenly
    g = abs(total) ** 0.5
    kv = [c / g for c in col]
    kh = [r / g * (1 if total >= 0 else -1) for r in row]
    for i in range(n):
        for j in range(n):
            if abs(kv[i] * kh[j] - rows[i][j]) > 1e-6 * max(abs(v) for v in kernel2d):
                return None
    if any(abs(a - b) > 1e-7 for a, b in zip(kv, kh)):
        return None
    return kv


class _ToRGB(torch.autograd.Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, x, s, skip, w_rgb, bias, taps2d, taps1d):
        b, cin, h, w = x.shape
        rgb = torch.empty((b, 3, h, w), device=x.device, dtype=torch.float32)
        N.check(N.load().w2e_torgb_fwd(N.ptr(x), N.ptr(w_rgb), N.ptr(s), N.ptr(bias), N.ptr(skip),
                                       N.host_floats(taps1d) if skip is not None else None, N.ptr(rgb), b, cin, h,
                                       w, N.dtype_code(x), N.stream_ptr()), "torgb_fwd")
        ctx.save_for_backward(x, s)
        ctx.cfg = (w_rgb, taps2d, skip is not None)
        return rgb

    @staticmethod
    @torch.autograd.function.once_differentiable
    @_amp_bwd
    def backward(ctx, g):
        x, s = ctx.saved_tensors
        w_rgb, taps2d, has_skip = ctx.cfg
        g = g.contiguous()
        b, cin, h, w = x.shape
        gx = gs = gskip = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            gx = torch.empty_like(x, dtype=torch.float32)
            gs = torch.empty((b, cin), device=x.device, dtype=torch.float32)
            lib = N.load()
            nseg = int(lib.w2e_rowdot_segments(b * cin, h * w)) if x.dtype == torch.float32 else 1
            if nseg > 1:   # long planes: segmented, deterministic two-pass reduction of gstyle
                partial = torch.empty((b * cin, nseg), device=x.device, dtype=torch.float32)
                N.check(lib.w2e_torgb_bwd_seg(N.ptr(g), N.ptr(x), N.ptr(w_rgb), N.ptr(s), N.ptr(gx), N.ptr(gs), N.ptr(partial),
                                              b, cin, h, w, nseg, N.stream_ptr()), "torgb_bwd_seg")
            else:
                N.check(lib.w2e_torgb_bwd(N.ptr(g), N.ptr(x), N.ptr(w_rgb), N.ptr(s), N.ptr(gx), N.ptr(gs), b, cin,
                                          h, w, N.stream_ptr()), "torgb_bwd")
        if has_skip and ctx.needs_input_grad[2]:
            gskip = torch.empty((b, 3, h // 2, w // 2), device=x.device, dtype=torch.float32)
            N.check(N.load().w2e_upfirdn2d_bwd(N.ptr(g), N.ptr(gskip), N.host_floats(taps2d), b * 3, h // 2, w // 2,
                                               4, 4, 2, 2, 1, 1, 2, 1, 2, 1, N.F32, N.stream_ptr()), "upfirdn2d_bwd")
        return gx, gs, gskip, None, None, None, None


def to_rgb(x, s, pw, bias, skip, up_kernel):
    """rgb = conv1x1(x*s, W) + bias + Upsample(skip)  (model.py:353-362), fused."""
    x = x.contiguous()
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    s = s.to(torch.float32).contiguous()
    taps2d = taps1d = None
    if skip is not None:
        taps2d = kernel_taps(up_kernel)
        taps1d = separable_taps(taps2d) if len(taps2d) == 16 else None
        if taps1d is None:  # unusual (non-separable / non-4x4) filter: generic upfirdn2d + add
            rgb = _ToRGB.apply(x, s, None, pw.rgb, bias.detach().reshape(3).contiguous(), None, None)
            p = up_kernel.shape[0] - 2
            return rgb + _UpFirDn2d.apply(skip.contiguous(), taps2d, up_kernel.shape[0], up_kernel.shape[1],
                                          (2, 2, 1, 1, (p + 1) // 2 + 1, p // 2, (p + 1) // 2 + 1, p // 2))
        skip = skip.to(torch.float32).contiguous()
    return _ToRGB.apply(x, s, skip, pw.rgb, bias.detach().reshape(3).to(torch.float32).contiguous(), taps2d, taps1d)


# --------------------------------------------------------------------------------------------
# region-mask blend
# --------------------------------------------------------------------------------------------
class _MaskBlend(torch.autograd.Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, edited, orig, mask):
        b, c, h, w = edited.shape
        out = torch.empty_like(edited)
        N.check(N.load().w2e_mask_blend_fwd(N.ptr(edited), N.ptr(orig), N.ptr(mask), N.ptr(out), b, c, h, w,
                                            mask.shape[2], mask.shape[3], N.dtype_code(edited), N.stream_ptr()),
                "mask_blend_fwd")
        ctx.save_for_backward(edited, orig, mask)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    @_amp_bwd
    def backward(ctx, g):
        edited, orig, mask = ctx.saved_tensors
        b, c, h, w = edited.shape
        g = g.contiguous().float()
        ge = torch.empty_like(g)
        gm = torch.empty_like(mask)
        ws = torch.empty((b, h, w), device=g.device, dtype=torch.float32)
        N.check(N.load().w2e_mask_blend_bwd(N.ptr(g), N.ptr(edited), N.ptr(orig), N.ptr(mask), N.ptr(ge), N.ptr(gm),
                                            N.ptr(ws), b, c, h, w, mask.shape[2], mask.shape[3], N.stream_ptr()),
                "mask_blend_bwd")
        return ge, None, gm


def mask_blend(edited, orig, mask):
    """m*edited + (1-m)*orig with m = nearest-resized [B,1,h,w] mask broadcast over channels
    (attention/attention_model.py:548-549); the resized/repeated mask is never materialised."""
    N.require_cuda(edited, orig, mask)
    if mask.ndim != 4 or mask.shape[1] != 1 or mask.shape[0] != edited.shape[0]:
        raise ValueError(f"attention_map must be [B,1,h,w], got {tuple(mask.shape)}")
    if orig.shape != edited.shape:
        raise ValueError(f"feature_map entry {tuple(orig.shape)} does not match layer output {tuple(edited.shape)}")
    edited = edited.to(torch.float32).contiguous()
    orig = orig.detach().to(torch.float32).contiguous()
    return _MaskBlend.apply(edited, orig, mask.to(torch.float32).contiguous())
