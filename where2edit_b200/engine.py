"""bf16 tensor-core synthesis engine: the no-grad fast path behind `Generator.forward` when
`precision == "bf16"` (attention/attention_model.py:473-676 semantics).

Data layout in HBM: activations are channels-last bf16 `[B,H,W,C]` (C innermost = the K dimension
of the implicit GEMM, so TMA boxes land directly in the UMMA operand layout).  Every producer
kernel writes its result ALREADY multiplied by the style of the layer that will consume it
(`out_mod`), which is how the style modulation of models/stylegan2/model.py:238-239 is folded
into the activation path: all samples share one bf16 weight tensor `[9][Cout][Cin]` per layer and
the demodulation (model.py:242-243) is a per-(sample, channel) scale in the GEMM epilogue.

Per resolution block (model.py:306-362):
    up-conv   w2e_modconv_tc2 (transposed: the 4 parity classes of conv_transpose2d stride 2 in one pass) -> z [(2h+1)^2]
    blur      w2e_blur_act_nhwc: 4x4 FIR + noise + bias + lrelu*sqrt2, writes act * s_conv
    conv      w2e_modconv_tc2[_rgb] (9 taps) + demod + noise + bias + lrelu*sqrt2 (+ the following ToRGB in the
              epilogue), writes act (for ToRGB / feature capture) and act * s_next_up
    ToRGB     fused above 16^2; w2e_torgb_nhwc otherwise: 1x1 modulated conv + bias + polyphase skip upsample
Styles, demodulation coefficients and the tiny modulation linears stay fp32.
"""
import os

import torch

from . import _native as N
from . import functional as K
from .op.upfirdn2d import kernel_taps


class SynthesisEngine:
    def __init__(self, gen):
        self.gen = gen
        self._w = {}
        self._err = N.ErrorFlag("libw2e modconv_tc2, bf16 synthesis engine")
        # which of the two (bit-identical) blur kernels runs.  Measured (B200, batch 32, tools/ab_layers.py): the templated
        # kernel (variant 1) wins at >= 64 channels (0.442 -> 0.391 ms at 64@512^2, 0.230 -> 0.205 at 128@256^2), the
        # run-time-tile-shape kernel (variant 0) at 32 channels (0.891 vs 1.172 ms at 32@1024^2): "auto" picks by the
        # channel count; W2E_BLUR_V2=0 / 1 forces one kernel everywhere (A/B).
        env = os.environ.get("W2E_BLUR_V2", "auto")
        self.blur_variant = int(env) if env in ("0", "1") else "auto"
        # x-pair mode of the RGB-only 32-channel last layer (w2e_modconv_tc2_rgb_pair); W2E_PAIR=0 disables (A/B)
        self.pair_mode = os.environ.get("W2E_PAIR", "1") == "1"
        # per-call tuning / A-B switches of the tcgen05 convolution (_native.tc2_config(...)); None = defaults
        self.tc2_cfg = None
        # dtype of the image the LAST layer's epilogue writes (torch.float32 as the reference, or torch.bfloat16:
        # half the bytes for a caller that gathers / copies the images out)
        self.image_dtype = torch.float32
        self.image_out = None   # optional caller-owned [B,3,H,W] buffer for the image (Generator.set_image_output)
        # W2E_FUSE_RGB=0 keeps ToRGB as its own kernel (w2e_torgb_nhwc) instead of the conv epilogue
        self.fuse_rgb = os.environ.get("W2E_FUSE_RGB", "1") == "1"
        # W2E_FUSE_UPBLUR=1 runs the up-convolution and its Blur as ONE kernel where w2e_modconv_tc2_upblur covers
        # the shape.  Off by default: parity-green, but its CUDA-core FIR phase is not yet faster than the
        # HBM round trip it removes.
        self.fuse_upblur = os.environ.get("W2E_FUSE_UPBLUR", "0") == "1"
        if not N.load().w2e_modconv_tc_supported():
            raise RuntimeError("where2edit_b200: precision='bf16' needs an sm_100 (B200) device and a driver with "
                               "cuTensorMapEncodeTiled; there is no fallback -- use precision='fp32'")

    # ------------------------------------------------------------------ cached derived weights
    def _tc_weight(self, conv):
        """bf16 [k*k][Cout][Cin] with the equalised-lr scale folded in (model.py:216-217)."""
        pw = conv.packed()
        pw.tc_fwd()
        return pw

    def error_flag(self, device):
        return self._err.tensor(device)

    def assert_ok(self):
        """Synchronising check of the pipeline-timeout flag of the tensor-core kernels (run() itself polls the
        flag of the previous calls without synchronising, see _native.ErrorFlag)."""
        self._err.check()


    # ------------------------------------------------------------------ styles / demodulation plan
    def _style_plan(self, layers, dev):
        """Concatenated modulation weights / wsq tables of every styled layer plus the device-side
        block->layer maps of w2e_style_mod_all / w2e_style_demod_all; rebuilt when a weight changes."""
        key = [str(dev)]
        for module, _ in layers:
            conv = module.conv
            key += [conv.weight.data_ptr(), conv.weight._version, conv.modulation.weight.data_ptr(),
                    conv.modulation.weight._version, conv.modulation.bias._version]
        key = tuple(key)
        plan = getattr(self, "_plan", None)
        if plan is not None and plan["key"] == key:
            return plan
        rows = self.gen.latent_rows(False)
        w_all, b_all, meta_mod, blk_mod = [], [], [], []
        wsq_all, meta_dem, d_off, blk_dem = [], [], [], []
        choff = coff = wsq_off = 0
        seg = []   # per layer: (channel offset, Cin, demod index or None)
        for li, ((module, kind), row) in enumerate(zip(layers, rows)):
            conv = module.conv
            mod = conv.modulation
            cin = conv.in_channel
            if cin % 32:
                raise RuntimeError(f"where2edit_b200 bf16 engine: in_channel {cin} is not a multiple of 32")
            w_all.append(mod.weight.detach().to(dev, torch.float32) * mod.scale)
            b_all.append(mod.bias.detach().to(dev, torch.float32) * mod.lr_mul)
            meta_mod.append([row, choff, cin, 0])
            blk_mod += [li] * (cin // 32)
            di = None
            if kind != "rgb":
                pw = self._tc_weight(conv)
                cout = conv.out_channel
                if cout % 8:
                    raise RuntimeError(f"where2edit_b200 bf16 engine: out_channel {cout} is not a multiple of 8")
                di = len(meta_dem)
                wsq_all.append(pw.wsq.reshape(-1))
                meta_dem.append([choff, cin, cout, wsq_off])
                d_off.append(coff)
                blk_dem += [di] * (cout // 8)
                wsq_off += cin * cout
                coff += cout
            seg.append((choff, cin, di))
            choff += cin
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
        self._plan = {
            "key": key, "seg": seg, "sum_c": choff, "sum_o": coff, "d_off": d_off, "cout": [m[2] for m in meta_dem],
            "w_all": torch.cat(w_all).contiguous(), "b_all": torch.cat(b_all).contiguous(),
            "wsq_all": torch.cat(wsq_all).contiguous(), "meta_mod": i32(meta_mod), "blk_mod": i32(blk_mod),
            "meta_dem": i32(meta_dem), "blk_dem": i32(blk_dem), "d_off_dev": i32(d_off),
            "style_dim": w_all[0].shape[1],
        }
        return self._plan

    def _styles_and_demods(self, layers, rows, latent, stylespace, batch, dev):
        plan = self._style_plan(layers, dev)
        lib = N.load()
        if stylespace:
            parts = []
            for (choff, cin, _), row in zip(plan["seg"], rows):
                s = latent[row]
                if s.numel() != batch * cin:
                    raise ValueError(f"stylespace entry {row} has {tuple(s.shape)}, expected [{batch},1,{cin},1,1]")
                parts.append(s.detach().reshape(-1).to(torch.float32))
            s_all = torch.cat(parts).contiguous()
        else:
            lat = latent.detach().to(torch.float32).contiguous()
            if lat.ndim != 3 or lat.shape[2] != plan["style_dim"]:
                raise ValueError(f"W+ latent must be [B, n_latent, {plan['style_dim']}], got {tuple(lat.shape)}")
            s_all = torch.empty(batch * plan["sum_c"], device=dev, dtype=torch.float32)
            N.note(kind="style", tag="all modulations")
            N.check(lib.w2e_style_mod_all(N.ptr(lat), lat.stride(0), lat.stride(1), N.ptr(plan["w_all"]),
                                          N.ptr(plan["b_all"]), N.ptr(plan["blk_mod"]), N.ptr(plan["meta_mod"]),
                                          N.ptr(s_all), batch, plan["style_dim"], plan["blk_mod"].numel(),
                                          N.stream_ptr()), "style_mod_all")
        d_all = torch.empty(batch * plan["sum_o"], device=dev, dtype=torch.float32)
        N.note(kind="style", tag="all demodulations")
        N.check(lib.w2e_style_demod_all(N.ptr(s_all), N.ptr(plan["wsq_all"]), N.ptr(plan["blk_dem"]),
                                        N.ptr(plan["meta_dem"]), N.ptr(plan["d_off_dev"]), N.ptr(d_all), batch,
                                        plan["blk_dem"].numel(), N.stream_ptr()), "style_demod_all")
        styles, demods = [], []
        for choff, cin, di in plan["seg"]:
            styles.append(s_all[batch * choff:batch * (choff + cin)].view(batch, cin))
            if di is None:
                demods.append(None)
            else:
                o, c = plan["d_off"][di], plan["cout"][di]
                demods.append(d_all[batch * o:batch * (o + c)].view(batch, c))
        return styles, demods

    # ------------------------------------------------------------------ kernel launches
    def _conv2(self, xs, pw, d, noise, noise_w, bias, next_scale, want_out, want_mod, transposed, act):
        """Persistent v2 kernel: plain 3x3 (transposed=False) or the fused 4-class conv_transpose x2."""
        b, h, w, _ = xs.shape
        dev = xs.device
        oh, ow = (2 * h + 1, 2 * w + 1) if transposed else (h, w)
        out = torch.empty((b, oh, ow, pw.cout), device=dev, dtype=torch.bfloat16) if want_out else None
        out_mod = torch.empty((b, oh, ow, pw.cout), device=dev, dtype=torch.bfloat16) if want_mod else None
        nb = 0 if noise is None else noise.shape[0]
        gh, gw = (h, w)
        N.note(kind="modconv", flops=2.0 * 9 * pw.cin * pw.cout * b * gh * gw,
               tag=f"{'up ' if transposed else ''}{pw.cin}->{pw.cout}@{gh}x{gw}")
        N.check(N.load().w2e_modconv_tc2(
            N.ptr(xs), N.ptr(pw.tc), N.ptr(d), N.ptr(bias), N.ptr(noise), N.ptr(noise_w), nb, N.ptr(next_scale),
            N.ptr(out), N.ptr(out_mod), N.ptr(self.error_flag(dev)), b, pw.cin, pw.cout, h, w, int(transposed), act,
            N.tc2_cfg(self.tc2_cfg), N.stream_ptr()), "modconv_tc2")
        return out, out_mod

    def _conv2_rgb(self, xs, pw, d, noise, noise_w, bias, next_scale, want_out, want_mod, rgb_module, rgb_style, skip,
                   rgb_dtype=torch.float32, rgb_out=None):
        """Plain 3x3 StyledConv with the following ToRGB fused into its epilogue."""
        b, h, w, _ = xs.shape
        dev = xs.device
        out = torch.empty((b, h, w, pw.cout), device=dev, dtype=torch.bfloat16) if want_out else None
        out_mod = torch.empty((b, h, w, pw.cout), device=dev, dtype=torch.bfloat16) if want_mod else None
        # Cout >= 256 runs as two channel blocks that ADD their partial ToRGB sums: the image starts at zero
        if pw.cout >= 256:
            rgb_dtype = torch.float32
        if rgb_out is not None and tuple(rgb_out.shape) == (b, 3, h, w) and rgb_out.dtype == rgb_dtype and pw.cout < 256:
            rgb = rgb_out
        else:
            rgb = (torch.zeros if pw.cout >= 256 else torch.empty)((b, 3, h, w), device=dev, dtype=rgb_dtype)
        rpw = rgb_module.conv.packed()
        taps1d = None
        if skip is not None:
            taps2d = kernel_taps(rgb_module.upsample.kernel)
            taps1d = K.separable_taps(taps2d) if len(taps2d) == 16 else None
            if taps1d is None:
                raise RuntimeError("where2edit_b200 bf16 engine: the ToRGB upsample kernel must be a separable 4x4 FIR")
            skip = skip.contiguous()
        rbias = rgb_module.bias.detach().reshape(3).to(torch.float32).contiguous()
        nb = 0 if noise is None else noise.shape[0]
        N.note(kind="modconv", flops=2.0 * 9 * pw.cin * pw.cout * b * h * w, tag=f"{pw.cin}->{pw.cout}@{h}x{w}+rgb")
        if self.pair_mode and not want_mod and pw.cin == 32 and pw.cout == 32 and w % 16 == 0 and h % 2 == 0 and h > 16:
            # RGB-only 32-channel layer: on pixel pairs (N = 64 MMAs)
            rc = N.load().w2e_modconv_tc2_rgb_pair(
                N.ptr(xs), N.ptr(pw.tc_pair()), N.ptr(d), N.ptr(bias), N.ptr(noise), N.ptr(noise_w), nb, N.ptr(out),
                N.ptr(self.error_flag(dev)), b, h, w, N.ACT_LRELU, N.ptr(rpw.rgb), N.ptr(rgb_style),
                N.ptr(rbias), N.ptr(skip), N.host_floats(taps1d) if taps1d is not None else None, N.ptr(rgb),
                N.dtype_code(rgb), N.tc2_cfg(self.tc2_cfg), N.stream_ptr())
            if rc != N.ERR_UNSUPPORTED:
                N.check(rc, "modconv_tc2_rgb_pair")
                return out, None, rgb
            N.STATS.launches["w2e_modconv_tc2_rgb_pair"] -= 1   # nothing was launched: the ordinary kernel below runs
            if N.STATS.trace:
                N.STATS.trace.pop()
            N.note(kind="modconv", flops=2.0 * 9 * pw.cin * pw.cout * b * h * w, tag=f"{pw.cin}->{pw.cout}@{h}x{w}+rgb")
        N.check(N.load().w2e_modconv_tc2_rgb(
            N.ptr(xs), N.ptr(pw.tc), N.ptr(d), N.ptr(bias), N.ptr(noise), N.ptr(noise_w), nb, N.ptr(next_scale),
            N.ptr(out), N.ptr(out_mod), N.ptr(self.error_flag(dev)), b, pw.cin, pw.cout, h, w, N.ACT_LRELU,
            N.ptr(rpw.rgb), N.ptr(rgb_style), N.ptr(rbias), N.ptr(skip),
            N.host_floats(taps1d) if taps1d is not None else None, N.ptr(rgb), N.dtype_code(rgb),
            N.tc2_cfg(self.tc2_cfg), N.stream_ptr()), "modconv_tc2_rgb")
        return out, out_mod, rgb

    def _upblur(self, xs, pw, d, blur_kernel, pad, bias, noise, noise_w, next_scale, want_out, want_mod):
        """Up-convolution + Blur + noise + bias + lrelu in ONE launch (w2e_modconv_tc2_upblur); returns None
        when the fused kernel does not cover the shape (the caller then runs the two-kernel path)."""
        b, h, w, _ = xs.shape
        taps = kernel_taps(blur_kernel)
        if len(taps) != 16 or tuple(pad) != (1, 1) or pw.cout > 64 or h <= 16:
            return None
        dev = xs.device
        out = torch.empty((b, 2 * h, 2 * w, pw.cout), device=dev, dtype=torch.bfloat16) if want_out else None
        out_mod = torch.empty((b, 2 * h, 2 * w, pw.cout), device=dev, dtype=torch.bfloat16) if want_mod else None
        nb = 0 if noise is None else noise.shape[0]
        # 2 x 2 halo recomputed per tile: the MMA does 64*16/(60*12) = 1.42x the algorithmic work
        N.note(kind="modconv", flops=2.0 * 9 * pw.cin * pw.cout * b * h * w, tag=f"up+blur {pw.cin}->{pw.cout}@{h}x{w}",
               bytes=2.0 * b * (h * w * pw.cin + 4 * h * w * pw.cout * (int(want_out) + int(want_mod))))
        rc = N.load().w2e_modconv_tc2_upblur(
            N.ptr(xs), N.ptr(pw.tc), N.ptr(d), N.host_floats(taps), N.ptr(bias), N.ptr(noise), N.ptr(noise_w), nb,
            N.ptr(next_scale), N.ptr(out), N.ptr(out_mod), N.ptr(self.error_flag(dev)), b, pw.cin, pw.cout, h, w,
            N.ACT_LRELU, N.tc2_cfg(self.tc2_cfg), N.stream_ptr())
        if rc == N.ERR_UNSUPPORTED:
            N.STATS.note = None
            return None
        N.check(rc, "modconv_tc2_upblur")
        return out, out_mod

    def _blur(self, z, blur_kernel, pad, bias, noise, noise_w, next_scale, want_out, want_mod, out_hw):
        b, ih, iw, c = z.shape
        dev = z.device
        out = torch.empty((b, out_hw[0], out_hw[1], c), device=dev, dtype=torch.bfloat16) if want_out else None
        out_mod = torch.empty((b, out_hw[0], out_hw[1], c), device=dev, dtype=torch.bfloat16) if want_mod else None
        nb = 0 if noise is None else noise.shape[0]
        n_out = int(want_out) + int(want_mod)
        N.note(kind="upfirdn2d", bytes=2.0 * b * c * (ih * iw + out_hw[0] * out_hw[1] * n_out),
               tag=f"blur {c}@{out_hw[0]}")
        N.check(N.load().w2e_blur_act_nhwc(
            N.ptr(z), N.host_floats(kernel_taps(blur_kernel)), N.ptr(bias), N.ptr(noise), N.ptr(noise_w), nb,
            N.ptr(next_scale), N.ptr(out), N.ptr(out_mod), b, c, ih, iw, pad[0], pad[0], out_hw[0], out_hw[1],
            N.ACT_LRELU, (int(c >= 64) if self.blur_variant == "auto" else self.blur_variant), N.stream_ptr()),
            "blur_act_nhwc")
        return out, out_mod

    def _to_nchw(self, x):
        b, h, w, c = x.shape
        y = torch.empty((b, c, h, w), device=x.device, dtype=torch.float32)
        N.note(kind="layout", bytes=6.0 * x.numel(), tag=f"nhwc->nchw {c}@{h}")
        N.check(N.load().w2e_nhwc_to_nchw_f32(N.ptr(x), N.ptr(y), b, c, h * w, N.BF16, N.stream_ptr()), "nhwc_to_nchw_f32")
        return y

    def _to_nhwc(self, x, style, batch):
        bx, c, h, w = x.shape
        x = x.to(torch.float32).contiguous()
        y = torch.empty((batch, h, w, c), device=x.device, dtype=torch.bfloat16)
        N.note(kind="layout", bytes=4.0 * x.numel() + 2.0 * y.numel(), tag=f"nchw->nhwc {c}@{h}")
        N.check(N.load().w2e_nchw_to_nhwc_mod(N.ptr(x), N.ptr(style), N.ptr(y), batch, bx, c, h * w, N.BF16, N.stream_ptr()),
                "nchw_to_nhwc_mod")
        return y

    def _torgb(self, x, module, s, skip):
        b, h, w, c = x.shape
        pw = module.conv.packed()
        rgb = torch.empty((b, 3, h, w), device=x.device, dtype=torch.float32)
        taps1d = None
        if skip is not None:
            taps2d = kernel_taps(module.upsample.kernel)
            taps1d = K.separable_taps(taps2d) if len(taps2d) == 16 else None
            if taps1d is None:
                raise RuntimeError("where2edit_b200 bf16 engine: the ToRGB upsample kernel must be a separable 4x4 FIR")
            skip = skip.contiguous()
        bias = module.bias.detach().reshape(3).to(torch.float32).contiguous()
        N.note(kind="torgb", bytes=b * h * w * (2.0 * c + 12.0 + (3.0 if skip is not None else 0.0)),
               tag=f"torgb {c}@{h}")
        N.check(N.load().w2e_torgb_nhwc(N.ptr(x), N.ptr(pw.rgb), N.ptr(s), N.ptr(bias), N.ptr(skip),
                                        N.host_floats(taps1d) if taps1d is not None else None, N.ptr(rgb), b, c, h, w,
                                        N.stream_ptr()), "torgb_nhwc")
        return rgb

    def _blend(self, out, orig_nchw, mask, next_scale, want_mod):
        b, h, w, c = out.shape
        if tuple(orig_nchw.shape) != (b, c, h, w):
            raise ValueError(f"feature_map entry {tuple(orig_nchw.shape)} does not match layer output {(b, c, h, w)}")
        if mask.ndim != 4 or mask.shape[1] != 1 or mask.shape[0] != b:
            raise ValueError(f"attention_map must be [B,1,h,w], got {tuple(mask.shape)}")
        orig = self._to_nhwc(orig_nchw.detach(), None, b)
        mask = mask.detach().to(torch.float32).contiguous()
        new = torch.empty_like(out)
        new_mod = torch.empty_like(out) if want_mod else None
        N.note(kind="blend", bytes=2.0 * out.numel() * (3 + int(want_mod)), tag=f"blend {c}@{h}")
        N.check(N.load().w2e_blend_nhwc(N.ptr(out), N.ptr(orig), N.ptr(mask), N.ptr(next_scale), N.ptr(new),
                                        N.ptr(new_mod), b, c, h, w, mask.shape[2], mask.shape[3], N.stream_ptr()),
                "blend_nhwc")
        return new, new_mod

    # ------------------------------------------------------------------ the forward
    @torch.no_grad()
    def run(self, latent, stylespace, noise, want_features=False, attention_layer=0, attention_map=None,
            feature_map=None, capture_layers=None):
        """One forward.  Launches on the generator's device (made current for the call, so a model on cuda:1
        works while cuda:0 is current); a pipeline timeout of an EARLIER call raises here (no synchronisation)."""
        self._err.poll()
        with torch.cuda.device(self.gen.input.input.device):
            out = self._run(latent, stylespace, noise, want_features, attention_layer, attention_map, feature_map,
                            capture_layers)
            self._err.publish()
        return out

    def _run(self, latent, stylespace, noise, want_features, attention_layer, attention_map, feature_map,
             capture_layers=None):
        gen = self.gen
        keep = None if capture_layers is None else set(int(i) for i in capture_layers)   # positions of the returned list
        layers = gen.styled_layers()
        rows = gen.latent_rows(stylespace)
        pick = (lambda r: latent[r]) if stylespace else (lambda r: latent[:, r])
        batch = (latent[0] if stylespace else latent).shape[0]
        dev = gen.input.input.device
        N.require_cuda(latent[0] if stylespace else latent)

        # styles and demodulation coefficients of ALL layers up front, in two launches (style_plan.cu)
        styles, demods = self._styles_and_demods(layers, rows, latent, stylespace, batch, dev)

        def consumer_style(idx):
            """style of the next 3x3 conv after layer idx (None after the last one)."""
            for j in range(idx + 1, len(layers)):
                if layers[j][1] != "rgb":
                    return styles[j]
            return None

        captured, style_vector = [], []
        fused_rgb = None
        carry = False
        skip = None
        noise_idx = 0
        xs = self._to_nhwc(gen.input.input.detach(), styles[0], batch)   # ConstantInput * s_conv1
        act = None
        hw = (gen.input.input.shape[2], gen.input.input.shape[3])
        for idx, ((module, kind), s) in enumerate(zip(layers, styles)):
            layer = idx + 1
            blend_here = bool(attention_layer) and layer == attention_layer
            if kind == "rgb":
                if fused_rgb is not None:
                    skip, fused_rgb = fused_rgb, None
                else:
                    skip = self._torgb(act, module, s, skip)
                if attention_layer and (blend_here or carry):
                    carry = False
                    skip = K.mask_blend(skip, feature_map[layer - 1], attention_map)
                captured.append(skip if (keep is None or idx in keep) else None)
                style_vector.append(s.reshape(batch, 1, -1, 1, 1))
                continue
            want_here = want_features and (keep is None or idx in keep)   # this layer's map goes back to the caller
            conv = module.conv
            pw = self._tc_weight(conv)
            nz = noise[noise_idx]
            noise_idx += 1
            if nz is None:  # model.py:286-288: fresh per-sample noise
                f = 2 if kind == "up" else 1
                nz = torch.empty((batch, 1, hw[0] * f, hw[1] * f), device=dev, dtype=torch.float32).normal_()
            nz = nz.to(torch.float32).contiguous()
            noise_w = module.noise.weight.detach().to(torch.float32).contiguous()
            bias = module.activate.bias.detach().to(torch.float32).contiguous()
            nxt = consumer_style(idx)
            next_is_rgb = idx + 1 < len(layers) and layers[idx + 1][1] == "rgb"
            need_out = next_is_rgb or want_here or blend_here
            need_mod = nxt is not None and not blend_here
            fuse = (kind == "conv" and next_is_rgb and self.fuse_rgb and pw.cout <= 512
                    and hw[0] > 16 and not (attention_layer and attention_layer in (layer, layer + 1)))
            if fuse:
                last = idx + 2 == len(layers)   # the image itself: written in the dtype the caller asked for
                act, xs_next, fused_rgb = self._conv2_rgb(xs, pw, demods[idx], nz, noise_w, bias, nxt,
                                                          want_here, need_mod, layers[idx + 1][0],
                                                          styles[idx + 1], skip,
                                                          self.image_dtype if last else torch.float32,
                                                          self.image_out if last else None)
            elif kind == "conv":
                act, xs_next = self._conv2(xs, pw, demods[idx], nz, noise_w, bias, nxt, need_out, need_mod,
                                           False, N.ACT_LRELU)
            else:
                h, w = hw
                fused = None
                if self.fuse_upblur:
                    fused = self._upblur(xs, pw, demods[idx], conv.blur.kernel, conv.blur.pad, bias, nz, noise_w, nxt,
                                         need_out, need_mod)
                hw = (2 * h, 2 * w)
                if fused is not None:
                    act, xs_next = fused
                else:
                    z, _ = self._conv2(xs, pw, demods[idx], None, None, None, None, True, False, True, N.ACT_NONE)
                    act, xs_next = self._blur(z, conv.blur.kernel, conv.blur.pad, bias, nz, noise_w, nxt, need_out,
                                              need_mod, hw)
            if blend_here:
                carry = True
                act, xs_next = self._blend(act, feature_map[layer - 1], attention_map, nxt, nxt is not None)
            xs = xs_next
            if want_features:
                captured.append(self._to_nchw(act) if want_here else None)
            style_vector.append(s.reshape(batch, 1, -1, 1, 1))
        if skip.dtype != self.image_dtype:   # image produced by an unfused ToRGB / blend (fp32): convert once
            skip = K.quantize_u8(skip) if self.image_dtype == torch.uint8 else skip.to(self.image_dtype)
        if self.image_out is not None and skip.data_ptr() != self.image_out.data_ptr() \
                and tuple(self.image_out.shape) == tuple(skip.shape):
            self.image_out.copy_(skip)
            skip = self.image_out
        return skip, style_vector, captured
