"""CPU oracle for the Where2edit StyleGAN2 synthesis hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.  The product package (where2edit_b200/) never does: it has no CPU path and
raises if its CUDA library is missing.

This is a functional restatement (parameters come in as a plain dict keyed like the reference
state dict) of the algorithm in the reference's pure-PyTorch modules:

  upfirdn2d / upfirdn2d_native ........ models/stylegan2/op/upfirdn2d.py:11-60
  fused_leaky_relu .................... models/stylegan2/op/fused_act.py:23-39
  PixelNorm / EqualLinear ............. models/stylegan2/model.py:11-17, 130-159
  ModulatedConv2d.forward ............. models/stylegan2/model.py:234-276
  NoiseInjection / StyledConv / ToRGB . models/stylegan2/model.py:279-362
  Generator.forward (+ mask blend) .... attention/attention_model.py:473-676

The arithmetic itself lives in third-party PyTorch (F.conv2d, F.conv_transpose2d, F.linear,
F.leaky_relu, F.pad, F.interpolate; reference pin torch 1.7.1, this image torch 2.11) and the
reference holds no golden vectors or tests, so the oracle is pinned against
outputs of the reference modules themselves, executed in the build container and committed as
tests/golden/*.npz by oracle/make_golden.py (tests/test_oracle_golden.py re-checks them, and
re-runs the live reference when /root/reference is present).

All functions take/return CPU torch tensors and run in the dtype of their inputs (fp32 or fp64).
"""
import math

import torch
import torch.nn.functional as F

SQRT2 = 2 ** 0.5


# --------------------------------------------------------------------------------------------
# ops
# --------------------------------------------------------------------------------------------
def upfirdn2d_ref(x, kernel, up=1, down=1, pad=(0, 0)):
    """models/stylegan2/op/upfirdn2d.py:11-16 (square up/down, same pad on x and y)."""
    return upfirdn2d_native_ref(x, kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])


def upfirdn2d_native_ref(x, kernel, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    """Port of the reference's steps (models/stylegan2/op/upfirdn2d.py:19-60): zero-stuff by
    `up` with trailing zeros, pad (negative pad crops), correlate with the flipped kernel
    through a single-channel conv2d, keep every `down`-th sample."""
    n, c, in_h, in_w = x.shape
    kh, kw = kernel.shape
    planes = x.reshape(n * c, in_h, 1, in_w, 1)
    planes = F.pad(planes, [0, up_x - 1, 0, 0, 0, up_y - 1])           # :28-30
    planes = planes.reshape(n * c, in_h * up_y, in_w * up_x)
    planes = F.pad(planes, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])  # :32-34
    planes = planes[:, max(-py0, 0): planes.shape[1] - max(-py1, 0),
                    max(-px0, 0): planes.shape[2] - max(-px1, 0)]      # :35-40
    filt = torch.flip(kernel, [0, 1]).reshape(1, 1, kh, kw).to(x.dtype)  # :46
    out = F.conv2d(planes.unsqueeze(1), filt)                            # :47
    out = out[:, 0, ::down_y, ::down_x]                                  # :55
    out_h = (in_h * up_y + py0 + py1 - kh) // down_y + 1                 # :57
    out_w = (in_w * up_x + px0 + px1 - kw) // down_x + 1                 # :58
    return out.reshape(n, c, out_h, out_w)


def upfirdn2d_direct(x, kernel, up_x, up_y, down_x, down_y, px0, px1, py0, py1):
    """Independent direct-definition evaluation (no conv2d): explicit tap loop.
    out[oy,ox] = sum_{ky,kx} k[ky,kx] * U[oy*down_y + (kh-1-ky) - py0, ox*down_x + (kw-1-kx) - px0]
    where U is x zero-stuffed by `up` (U[iy*up_y, ix*up_x] = x[iy,ix]).  Used to cross-check
    upfirdn2d_native_ref in tests."""
    n, c, in_h, in_w = x.shape
    kh, kw = kernel.shape
    uh, uw = in_h * up_y, in_w * up_x
    u = x.new_zeros(n, c, uh, uw)
    u[:, :, ::up_y, ::up_x] = x
    ph, pw = uh + py0 + py1, uw + px0 + px1
    p = x.new_zeros(n, c, max(ph, 0), max(pw, 0))
    sy0, sx0 = max(-py0, 0), max(-px0, 0)
    sy1, sx1 = uh - max(-py1, 0), uw - max(-px1, 0)
    dy0, dx0 = max(py0, 0), max(px0, 0)
    p[:, :, dy0:dy0 + (sy1 - sy0), dx0:dx0 + (sx1 - sx0)] = u[:, :, sy0:sy1, sx0:sx1]
    fh, fw = ph - kh + 1, pw - kw + 1
    full = x.new_zeros(n, c, fh, fw)
    for ky in range(kh):
        for kx in range(kw):
            coeff = kernel[kh - 1 - ky, kw - 1 - kx].to(x.dtype)
            full += coeff * p[:, :, ky:ky + fh, kx:kx + fw]
    return full[:, :, ::down_y, ::down_x].contiguous()


def fused_leaky_relu_ref(x, bias, negative_slope=0.2, scale=SQRT2):
    """models/stylegan2/op/fused_act.py:23-39.  Bias sits on dim 1 for 2-D/4-D inputs and on
    the LAST dim for 3-D inputs (:26-32).  (The reference's unconditional .cuda() at :25 is a
    device move, not arithmetic.)"""
    if x.ndim == 3:
        shaped = bias.reshape(1, 1, -1)
    else:
        shaped = bias.reshape(1, -1, *([1] * (x.ndim - 2)))
    return F.leaky_relu(x + shaped, negative_slope=negative_slope) * scale


def pixel_norm_ref(x, dim=1):
    """models/stylegan2/model.py:11-17."""
    return x * torch.rsqrt(torch.mean(x ** 2, dim=dim, keepdim=True) + 1e-8)


def equal_linear_ref(x, weight, bias, lr_mul=1.0, activation=None):
    """models/stylegan2/model.py:149-159 (scale = lr_mul / sqrt(in_dim), :146)."""
    scale = (1 / math.sqrt(weight.shape[1])) * lr_mul
    if activation:
        return fused_leaky_relu_ref(F.linear(x, weight * scale), bias * lr_mul)
    return F.linear(x, weight * scale, bias=bias * lr_mul)


def nearest_resize_ref(mask, size):
    """F.interpolate(mask, size) default mode = nearest (attention/attention_model.py:548):
    src = floor(dst * in / out).  Integer restatement (exact for any ratio)."""
    b, c, h, w = mask.shape
    iy = (torch.arange(size) * h) // size
    ix = (torch.arange(size) * w) // size
    return mask[:, :, iy][:, :, :, ix]


def mask_blend_ref(edited, original, mask):
    """attention/attention_model.py:548-549: m.repeat(C) * out + (1 - m.repeat(C)) * orig."""
    m = nearest_resize_ref(mask, edited.shape[-1]).to(edited.dtype)
    return m * edited + (1 - m) * original


# --------------------------------------------------------------------------------------------
# modules (functional)
# --------------------------------------------------------------------------------------------
def modulated_conv2d_ref(x, style, weight, mod_weight, mod_bias, demodulate=True, upsample=False,
                         blur_kernel=None, input_is_stylespace=False):
    """models/stylegan2/model.py:234-276.  weight is [1,Cout,Cin,k,k]; returns (out, style)
    with style [B,1,Cin,1,1].  Up path: grouped conv_transpose2d stride 2 then Blur pad
    (pad0,pad1) from :199-206."""
    b, cin, h, w = x.shape
    _, cout, _, k, _ = weight.shape
    if not input_is_stylespace:
        style = equal_linear_ref(style, mod_weight, mod_bias).reshape(b, 1, cin, 1, 1)
    wscale = 1 / math.sqrt(cin * k * k)
    wmod = wscale * weight * style
    if demodulate:
        d = torch.rsqrt(wmod.pow(2).sum([2, 3, 4]) + 1e-8)
        wmod = wmod * d.reshape(b, cout, 1, 1, 1)
    if upsample:
        wt = wmod.transpose(1, 2).reshape(b * cin, cout, k, k)
        out = F.conv_transpose2d(x.reshape(1, b * cin, h, w), wt, padding=0, stride=2, groups=b)
        out = out.reshape(b, cout, out.shape[-2], out.shape[-1])
        factor = 2
        p = (blur_kernel.shape[0] - factor) - (k - 1)
        out = upfirdn2d_ref(out, blur_kernel, pad=((p + 1) // 2 + factor - 1, p // 2 + 1))
    else:
        out = F.conv2d(x.reshape(1, b * cin, h, w), wmod.reshape(b * cout, cin, k, k),
                       padding=k // 2, groups=b)
        out = out.reshape(b, cout, out.shape[-2], out.shape[-1])
    return out, style


def _styled_conv(sd, prefix, x, style, noise, upsample, stylespace):
    """models/stylegan2/model.py:334-340: modconv -> + noise_w * noise -> fused lrelu."""
    out, s = modulated_conv2d_ref(
        x, style, sd[f"{prefix}.conv.weight"], sd[f"{prefix}.conv.modulation.weight"],
        sd[f"{prefix}.conv.modulation.bias"], demodulate=True, upsample=upsample,
        blur_kernel=sd.get(f"{prefix}.conv.blur.kernel"), input_is_stylespace=stylespace)
    out = out + sd[f"{prefix}.noise.weight"] * noise          # :290
    out = fused_leaky_relu_ref(out, sd[f"{prefix}.activate.bias"])
    return out, s


def _to_rgb(sd, prefix, x, style, skip, stylespace):
    """models/stylegan2/model.py:353-362: 1x1 modconv (no demod) + bias + Upsample(skip)."""
    out, s = modulated_conv2d_ref(
        x, style, sd[f"{prefix}.conv.weight"], sd[f"{prefix}.conv.modulation.weight"],
        sd[f"{prefix}.conv.modulation.bias"], demodulate=False, input_is_stylespace=stylespace)
    out = out + sd[f"{prefix}.bias"]
    if skip is not None:
        kern = sd[f"{prefix}.upsample.kernel"]
        p = kern.shape[0] - 2
        out = out + upfirdn2d_ref(skip, kern, up=2, down=1, pad=((p + 1) // 2 + 1, p // 2))
    return out, s


def mapping_ref(sd, z, n_mlp=8, lr_mlp=0.01):
    """Generator.style: PixelNorm + n_mlp x EqualLinear(fused_lrelu) (model.py:379-390)."""
    h = pixel_norm_ref(z)
    for i in range(n_mlp):
        h = equal_linear_ref(h, sd[f"style.{i + 1}.weight"], sd[f"style.{i + 1}.bias"],
                             lr_mul=lr_mlp, activation="fused_lrelu")
    return h


def generator_forward_ref(sd, styles, size, n_mlp=8, return_latents=False, return_features=False,
                          inject_index=None, truncation=1, truncation_latent=None,
                          input_is_latent=False, input_is_stylespace=False, noise=None,
                          attention_layer=0, attention_map=None, feature_map=None):
    """attention/attention_model.py:473-676 with randomize_noise=False (fixed noise buffers,
    :492-498).  Returns the same tuple arities as the reference (:666-676)."""
    log_size = int(math.log2(size))
    n_latent = log_size * 2 - 2
    num_layers = (log_size - 2) * 2 + 1
    if not input_is_latent and not input_is_stylespace:
        styles = [mapping_ref(sd, s, n_mlp) for s in styles]
    if noise is None:
        noise = [sd[f"noises.noise_{i}"] for i in range(num_layers)]
    if truncation < 1 and not input_is_stylespace:
        styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
    if input_is_stylespace:
        latent = styles[0]
    elif len(styles) < 2:
        latent = styles[0]
        if latent.ndim < 3:
            latent = latent.unsqueeze(1).repeat(1, n_latent, 1)
    else:
        assert inject_index is not None, "oracle needs an explicit inject_index for style mixing"
        latent = torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                            styles[1].unsqueeze(1).repeat(1, n_latent - inject_index, 1)], 1)

    blending = attention_map is not None
    captured, style_vector = [], []
    state = {"layer": 0, "carry": False}

    def pick(idx):
        return latent[idx] if input_is_stylespace else latent[:, idx]

    def after_conv(out):
        # :542-550 (and siblings): capture, or blend when this is the attention layer
        if blending:
            state["layer"] += 1
            if state["layer"] == attention_layer:
                state["carry"] = True
                out = mask_blend_ref(out, feature_map[state["layer"] - 1], attention_map)
        captured.append(out)
        return out

    def after_rgb(skip):
        # :554-562: the ToRGB following the attention layer is blended as well (`or this_layer`)
        if blending:
            state["layer"] += 1
            if state["layer"] == attention_layer or state["carry"]:
                state["carry"] = False
                skip = mask_blend_ref(skip, feature_map[state["layer"] - 1], attention_map)
        captured.append(skip)
        return skip

    batch = (latent[0] if input_is_stylespace else latent).shape[0]
    out = sd["input.input"].repeat(batch, 1, 1, 1)                       # model.py:299-303
    out, s = _styled_conv(sd, "conv1", out, pick(0), noise[0], False, input_is_stylespace)
    out = after_conv(out)
    style_vector.append(s)
    skip, s = _to_rgb(sd, "to_rgb1", out, pick(1), None, input_is_stylespace)
    skip = after_rgb(skip)
    style_vector.append(s)
    i = 2 if input_is_stylespace else 1
    for j in range(log_size - 2):
        out, s1 = _styled_conv(sd, f"convs.{2 * j}", out, pick(i), noise[1 + 2 * j], True,
                               input_is_stylespace)
        out = after_conv(out)
        out, s2 = _styled_conv(sd, f"convs.{2 * j + 1}", out, pick(i + 1), noise[2 + 2 * j], False,
                               input_is_stylespace)
        out = after_conv(out)
        skip, s3 = _to_rgb(sd, f"to_rgbs.{j}", out, pick(i + 2), skip, input_is_stylespace)
        skip = after_rgb(skip)
        style_vector.extend([s1, s2, s3])
        i += 3 if input_is_stylespace else 2                               # :630 / :664
    image = skip
    if return_latents:
        return image, latent, style_vector
    if return_features:
        return image, latent, style_vector, captured
    return image, None
