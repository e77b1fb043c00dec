"""Channels-last bf16 forward + backward engine: `Generator.forward` with `precision == "bf16"` UNDER AUTOGRAD
(BASELINE config 4: the latent / style optimisation loops of attention/run_attention.py:1233-1424 and
mapper/training/coach.py:81-92, whose losses backpropagate through the frozen generator into W+ or stylespace codes).

Forward: the kernels of `engine.SynthesisEngine` (tcgen05 modulated convolutions, TMA blur, fused ToRGB), which here
also keep every layer's unmodulated activation and the pre-blur tensors.  Backward (math: SURVEY.md appendix C), layer
by layer in reverse, all channels-last bf16, no NCHW <-> NHWC pass anywhere:

    plain conv l   w2e_grad_assemble_nhwc   g_a = gxs_next*s_next + ToRGB pull-back; gz = g_a*lrelu'(a)*d;
                                            reduces dL/ds_next (direct term), dL/ds_rgb, d*dL/dd
                   w2e_modconv_tc2          dgrad = the same tensor-core kernel with flipped taps, Cin <-> Cout
    up conv l      w2e_grad_assemble_nhwc   (activation sits after the Blur: no demodulation here)
                   w2e_blur_act_nhwc        Blur^T (pad 2) fused with the multiplication by d
                   w2e_rowdot_nhwc          d*dL/dd = sum gz*z / d
                   w2e_modconv_tc2_view x4  dgrad of conv_transpose2d(stride 2): one launch per output-parity class of
                                            the gradient, 4 / 2 / 2 / 1 taps, + w2e_sum4_nhwc
    ToRGB          folded into grad_assemble; the skip path is w2e_upfirdn2d_bwd on the 3-channel image gradient
    styles         gs = direct + 2*s*((-d^2 G / 2) @ wsq); W+ through the modulation linears (small library GEMMs)

The generator is frozen in every caller (run_attention.py:1051), so no weight gradient is produced here; a generator
with trainable parameters, a region blend or per-sample random noise takes the module path (model.Generator).
"""
import torch

from . import _native as N
from . import functional as K
from .engine import SynthesisEngine
from .op.upfirdn2d import kernel_taps

# window taps (bit ky*3+kx of the 3x3 kernel) used by the dgrad of the transposed convolution per gradient class (py, px)
_CLASS_TAPS = {(0, 0): (1 << 4) | (1 << 5) | (1 << 7) | (1 << 8), (0, 1): (1 << 4) | (1 << 7),
               (1, 0): (1 << 4) | (1 << 5), (1, 1): 1 << 4}


class _Saved:
    pass


class TrainEngine(SynthesisEngine):
    dgrad_up_fused = True      # A/B switch: False = never the one-launch dgrad of the transposed convolution
    dgrad_up_in_place = True   # A/B switch: False = always separate class results + w2e_sum4_nhwc

    # ------------------------------------------------------------------ forward (keeps what the backward needs)
    def forward_train(self, latent, stylespace, noise, attention_layer=0, attention_map=None, feature_map=None):
        gen = self.gen
        layers = gen.styled_layers()
        rows = gen.latent_rows(stylespace)
        batch = (latent[0] if stylespace else latent).shape[0]
        dev = gen.input.input.device
        styles, demods = self._styles_and_demods(layers, rows, latent, stylespace, batch, dev)

        def consumer_style(idx):
            for j in range(idx + 1, len(layers)):
                if layers[j][1] != "rgb":
                    return styles[j]
            return None

        sv = _Saved()
        sv.batch, sv.stylespace, sv.styles, sv.demods = batch, stylespace, styles, demods
        sv.layers, sv.rows = layers, rows
        sv.act, sv.z, sv.noise, sv.hw = {}, {}, {}, {}
        sv.blend = {}        # layer index -> (tensor before the blend, original feature, kind): region blend sites
        sv.mask = attention_map.detach().to(torch.float32).contiguous() if attention_layer else None
        carry = False
        const = gen.input.input.detach()
        sv.const = self._to_nhwc(const, None, 1)                     # unmodulated, batch-broadcast
        xs = self._to_nhwc(const, styles[0], batch)
        hw = (const.shape[2], const.shape[3])
        skip = None
        fused_rgb = None
        act = None
        noise_idx = 0
        style_vector = []
        for idx, ((module, kind), s) in enumerate(zip(layers, styles)):
            if kind == "rgb":
                if fused_rgb is not None:
                    skip, fused_rgb = fused_rgb, None
                else:
                    skip = self._torgb(act, module, s, skip)
                if attention_layer and (idx + 1 == attention_layer or carry):     # attention_model.py:558-561 (`or this_layer`)
                    carry = False
                    orig = feature_map[idx].detach().to(torch.float32).contiguous()
                    sv.blend[idx] = (skip, orig, "rgb")
                    skip = K.mask_blend(skip, orig, sv.mask)
                style_vector.append(s.reshape(batch, 1, -1, 1, 1))
                continue
            conv = module.conv
            pw = self._tc_weight(conv)
            nz = noise[noise_idx].to(torch.float32).contiguous()
            noise_idx += 1
            noise_w = module.noise.weight.detach().to(torch.float32).contiguous()
            bias = module.activate.bias.detach().to(torch.float32).contiguous()
            nxt = consumer_style(idx)
            next_is_rgb = idx + 1 < len(layers) and layers[idx + 1][1] == "rgb"
            blend_here = bool(attention_layer) and idx + 1 == attention_layer
            need_mod = nxt is not None and not blend_here
            if kind == "conv":
                fuse = (next_is_rgb and self.fuse_rgb and pw.cout <= 512 and hw[0] > 16
                        and not (attention_layer and attention_layer in (idx + 1, idx + 2)))
                if fuse:
                    act, xs_next, fused_rgb = self._conv2_rgb(xs, pw, demods[idx], nz, noise_w, bias, nxt, True, need_mod,
                                                              layers[idx + 1][0], styles[idx + 1], skip)
                else:
                    act, xs_next = self._conv2(xs, pw, demods[idx], nz, noise_w, bias, nxt, True, need_mod, False,
                                               N.ACT_LRELU)
            else:
                h, w = hw
                z, _ = self._conv2(xs, pw, demods[idx], None, None, None, None, True, False, True, N.ACT_NONE)
                hw = (2 * h, 2 * w)
                act, xs_next = self._blur(z, conv.blur.kernel, conv.blur.pad, bias, nz, noise_w, nxt, True, need_mod, hw)
                sv.z[idx] = z
            sv.act[idx], sv.noise[idx], sv.hw[idx] = act, (nz, noise_w, bias), hw
            if blend_here:      # attention_model.py:546-549: m * out + (1 - m) * original feature, consumed by what follows
                carry = True
                blended, xs_next = self._blend(act, feature_map[idx], sv.mask, nxt, nxt is not None)
                sv.blend[idx] = (blended, None, "conv")
                sv.blend_orig = getattr(sv, "blend_orig", {})
                sv.blend_orig[idx] = feature_map[idx]
                act = blended   # the following ToRGB (and the next convolution, through xs_next) read the blended tensor
            xs = xs_next
            style_vector.append(s.reshape(batch, 1, -1, 1, 1))
        return skip, style_vector, sv

    # ------------------------------------------------------------------ backward
    def _workspace(self, b, hw, c, dev):
        n = int(N.load().w2e_grad_assemble_workspace(b, hw, c))
        ws = getattr(self, "_ws", None)
        if ws is None or ws.numel() < n or ws.device != dev:
            ws = self._ws = torch.empty(max(n, 1), device=dev, dtype=torch.float32)
        return ws

    def _assemble(self, gxs, s_next, act, g_rgb, w_rgb, s_rgb, noise, demod, want_gz=True, act_kind=N.ACT_LRELU):
        b, h, w, c = act.shape
        dev = act.device
        gz = torch.empty_like(act) if want_gz else None
        sums = torch.empty((b, 3, c), device=dev, dtype=torch.float32)
        nz, noise_w, bias = noise if noise is not None else (None, None, None)
        N.note(kind="bwd_elementwise", bytes=2.0 * act.numel() * (2 + int(gxs is not None)), tag=f"grad_assemble {c}@{h}")
        N.check(N.load().w2e_grad_assemble_nhwc(
            N.ptr(gxs), N.ptr(s_next), N.ptr(act), N.ptr(g_rgb), N.ptr(w_rgb), N.ptr(s_rgb), N.ptr(nz), N.ptr(noise_w),
            0 if nz is None else nz.shape[0], N.ptr(bias), N.ptr(demod), act_kind, N.ptr(gz), N.ptr(sums),
            N.ptr(self._workspace(b, h * w, c, dev)), b, h * w, c, N.stream_ptr()), "grad_assemble_nhwc")
        return gz, sums

    def _rowdot(self, a, b_):
        b, h, w, c = a.shape
        dot = torch.empty((b, c), device=a.device, dtype=torch.float32)
        N.note(kind="bwd_elementwise", bytes=4.0 * a.numel(), tag=f"rowdot {c}@{h}")
        N.check(N.load().w2e_rowdot_nhwc(N.ptr(a), N.ptr(b_), b_.shape[0], N.ptr(dot), N.ptr(self._workspace(b, h * w, c, a.device)),
                                         b, h * w, c, N.stream_ptr()), "rowdot_nhwc")
        return dot

    def _dgrad_plain(self, gz, pw):
        """gxs [B,H,W,Cin] = conv^T(gz): the forward kernel with flipped taps and swapped channel roles."""
        b, h, w, _ = gz.shape
        out = torch.empty((b, h, w, pw.cin), device=gz.device, dtype=torch.bfloat16)
        N.note(kind="modconv", flops=2.0 * 9 * pw.cin * pw.cout * b * h * w, tag=f"dgrad {pw.cout}->{pw.cin}@{h}x{w}")
        if self.pair_mode and pw.cin == 32 and pw.cout == 32 and w % 16 == 0 and h > 16:
            # the 32-channel layer on pixel pairs (N = 64 MMAs), as in the forward
            rc = N.load().w2e_modconv_tc2_pair(N.ptr(gz), N.ptr(pw.tc_dgrad_pair()), N.ptr(out),
                                               N.ptr(self.error_flag(gz.device)), b, h, w, N.tc2_cfg(self.tc2_cfg),
                                               N.stream_ptr())
            if rc != N.ERR_UNSUPPORTED:
                N.check(rc, "modconv_tc2_pair")
                return out
            N.STATS.launches["w2e_modconv_tc2_pair"] -= 1
            if N.STATS.trace:
                N.STATS.trace.pop()
            N.note(kind="modconv", flops=2.0 * 9 * pw.cin * pw.cout * b * h * w, tag=f"dgrad {pw.cout}->{pw.cin}@{h}x{w}")
        N.check(N.load().w2e_modconv_tc2(
            N.ptr(gz), N.ptr(pw.tc_dgrad()), None, None, None, None, 0, None, N.ptr(out), None,
            N.ptr(self.error_flag(gz.device)), b, pw.cout, pw.cin, h, w, 0, N.ACT_NONE, N.tc2_cfg(self.tc2_cfg),
            N.stream_ptr()), "modconv_tc2 (dgrad)")
        return out

    def _dgrad_up(self, gz, pw, h, w):
        """gxs [B,h,w,Cin] of the transposed x2 convolution from its (2h+1)^2 output gradient: a stride-2 convolution,
        run per output-parity class of gz (strided views) with the taps that class feeds, then summed."""
        b, zh, zw, c = gz.shape
        lib = N.load()
        out = torch.empty((b, h, w, pw.cin), device=gz.device, dtype=torch.bfloat16)
        if self.dgrad_up_fused and 9 * pw.cin * pw.cout * 2 <= 80 * 1024 and h > 16:
            # one launch: the four parity-class tiles of gz feed one accumulator (gz read once, gx written once); only
            # where the nine weight blocks fit in shared memory next to them (the 64 -> 32 channel layer)
            N.note(kind="modconv", flops=2.0 * 9 * pw.cin * pw.cout * b * h * w, tag=f"dgrad up {pw.cout}->{pw.cin}@{h}x{w} fused")
            rc = lib.w2e_modconv_tc2_dgrad_up(N.ptr(gz), N.ptr(pw.tc_dgrad_up_fused()), None, N.ptr(out),
                                              N.ptr(self.error_flag(gz.device)), b, pw.cout, pw.cin, h, w,
                                              N.tc2_cfg(self.tc2_cfg), N.stream_ptr())
            if rc != N.ERR_UNSUPPORTED:
                N.check(rc, "modconv_tc2_dgrad_up")
                return out
            N.STATS.launches["w2e_modconv_tc2_dgrad_up"] -= 1
            if N.STATS.trace:
                N.STATS.trace.pop()
        if self.dgrad_up_fused and pw.cout in (64, 128) and h >= 16:
            # K-loop form: one accumulator, the K loop walks the four parity classes of gz (weights through the ring).
            # Measured at batch 16: 64->128@256^2 0.53 -> 0.24 ms, 128->256@128^2 0.34 -> 0.26; at 256 / 512 channels the
            # per-class launches are as fast or faster (0.32 vs 0.30, 0.22 vs 0.24 ms) and stay.
            masks = sum(_CLASS_TAPS[(c >> 1, c & 1)] << (9 * c) for c in range(4))
            N.note(kind="modconv", flops=2.0 * 9 * pw.cin * pw.cout * b * h * w, tag=f"dgrad up {pw.cout}->{pw.cin}@{h}x{w} fused")
            rc = lib.w2e_modconv_tc2_dgrad_up_k(N.ptr(gz), N.ptr(pw.tc_dgrad_up_k()), masks, None, N.ptr(out),
                                                N.ptr(self.error_flag(gz.device)), b, pw.cout, pw.cin, h, w,
                                                N.tc2_cfg(self.tc2_cfg), N.stream_ptr())
            if rc != N.ERR_UNSUPPORTED:
                N.check(rc, "modconv_tc2_dgrad_up_k")
                return out
            N.STATS.launches["w2e_modconv_tc2_dgrad_up_k"] -= 1
            if N.STATS.trace:
                N.STATS.trace.pop()
        wts = pw.tc_dgrad_up()
        # >= 32 rows: the four class launches accumulate IN PLACE (class 00 stores the h x w region, the others add to
        # it through TMA reduce-add); smaller maps: separate results + w2e_sum4_nhwc
        in_place = self.dgrad_up_in_place and h >= 32
        parts = []
        for py in (0, 1):
            for px in (0, 1):
                hc, wc = h + 1 - py, w + 1 - px
                view = gz.reshape(-1)[(py * zw + px) * c:]
                y = out if in_place else torch.empty((b, hc, wc, pw.cin), device=gz.device, dtype=torch.bfloat16)
                ntaps = bin(_CLASS_TAPS[(py, px)]).count("1")
                N.note(kind="modconv", flops=2.0 * ntaps * pw.cin * pw.cout * b * hc * wc,
                       tag=f"dgrad up {pw.cout}->{pw.cin}@{hc}x{wc} class {py}{px}")
                N.check(lib.w2e_modconv_tc2_view(
                    N.ptr(view), N.ptr(wts[(py, px)]), None, None, N.ptr(y), None, N.ptr(self.error_flag(gz.device)),
                    b, pw.cout, pw.cin, hc, wc, 2 * c, 2 * zw * c, zh * zw * c, _CLASS_TAPS[(py, px)],
                    h if in_place else hc, w if in_place else wc, int(in_place and (py, px) != (0, 0)),
                    N.tc2_cfg(self.tc2_cfg), N.stream_ptr()), "modconv_tc2_view (dgrad up)")
                parts.append(y)
        if in_place:
            return out
        N.note(kind="bwd_elementwise", bytes=2.0 * 5 * out.numel(), tag=f"sum4 {pw.cin}@{h}")
        N.check(lib.w2e_sum4_nhwc(N.ptr(parts[0]), N.ptr(parts[1]), N.ptr(parts[2]), N.ptr(parts[3]), N.ptr(out), b, h, w,
                                  pw.cin, N.stream_ptr()), "sum4_nhwc")
        return out

    def _blur_t(self, g, blur_kernel, demod, out_hw):
        """Blur^T (upfirdn2d with the flipped kernel and pad (2,2): (2h)^2 -> (2h+1)^2) times demod[b,c]."""
        b, ih, iw, c = g.shape
        out = torch.empty((b, out_hw[0], out_hw[1], c), device=g.device, dtype=torch.bfloat16)
        taps = list(reversed(kernel_taps(blur_kernel)))
        N.note(kind="upfirdn2d", bytes=2.0 * b * c * (ih * iw + out_hw[0] * out_hw[1]), tag=f"blur^T {c}@{ih}")
        N.check(N.load().w2e_blur_act_nhwc(
            N.ptr(g), N.host_floats(taps), None, None, None, 0, N.ptr(demod), None, N.ptr(out), b, c, ih, iw, 2, 2,
            out_hw[0], out_hw[1], N.ACT_NONE, (int(c >= 64) if self.blur_variant == "auto" else self.blur_variant),
            N.stream_ptr()), "blur_act_nhwc (transposed)")
        return out

    def _blend_backward(self, g_blended, act, orig_nchw, mask):
        """g w.r.t. m * act + (1 - m) * orig (channels-last bf16, one layer of <= 64^2 in the published configuration)
        -> (g_act bf16 channels-last, g_mask [B,1,mh,mw]) through the fp32 blend-backward kernel between two layout
        passes (attention_model.py:548-549; no gradient to the original feature, run_attention.py:1195-1203)."""
        b, h, w, c = act.shape
        lib = N.load()
        g32, a32 = self._to_nchw(g_blended), self._to_nchw(act)
        o32 = orig_nchw.detach().to(torch.float32).contiguous()
        ge = torch.empty_like(g32)
        gm = torch.empty_like(mask)
        ws = torch.empty((b, h, w), device=g32.device, dtype=torch.float32)
        N.check(lib.w2e_mask_blend_bwd(N.ptr(g32), N.ptr(a32), N.ptr(o32), N.ptr(mask), N.ptr(ge), N.ptr(gm), N.ptr(ws), b, c, h,
                                       w, mask.shape[2], mask.shape[3], N.stream_ptr()), "mask_blend_bwd")
        return self._to_nhwc(ge, None, b), gm

    def backward_train(self, sv, g_img, feature_map=None):
        gen = self.gen
        layers, styles, demods, batch = sv.layers, sv.styles, sv.demods, sv.batch
        dev = g_img.device
        lib = N.load()
        g_rgb = g_img.detach().to(torch.float32).contiguous()
        g_mask = None
        gs = [None] * len(layers)        # dL/d(style) per styled layer, [B, Cin]
        gxs_next, s_next, next_idx = None, None, None
        pending = None                   # (g_rgb at this resolution, rgb layer index)
        conv_ids = [i for i, (_, k) in enumerate(layers) if k != "rgb"]
        demod_terms = {}
        for idx in range(len(layers) - 1, -1, -1):
            module, kind = layers[idx]
            if kind == "rgb":
                if idx in sv.blend:      # blended skip image: g_skip = m * g, g_mask += sum g * (skip - orig)
                    skip_pre, orig, _ = sv.blend[idx]
                    b, _, h, w = g_rgb.shape
                    ge, gm = torch.empty_like(g_rgb), torch.empty_like(sv.mask)
                    ws = torch.empty((b, h, w), device=dev, dtype=torch.float32)
                    N.check(lib.w2e_mask_blend_bwd(N.ptr(g_rgb), N.ptr(skip_pre), N.ptr(orig), N.ptr(sv.mask), N.ptr(ge), N.ptr(gm),
                                                   N.ptr(ws), b, 3, h, w, sv.mask.shape[2], sv.mask.shape[3], N.stream_ptr()),
                            "mask_blend_bwd")
                    g_rgb = ge
                    g_mask = gm if g_mask is None else g_mask + gm
                pending = (g_rgb, idx)
                if hasattr(module, "upsample"):     # the skip path: gradient of upfirdn2d(skip, up=2, pad=(2,1))
                    b, _, h, w = g_rgb.shape
                    taps2d = kernel_taps(module.upsample.kernel)
                    taps1d = K.separable_taps(taps2d) if len(taps2d) == 16 else None
                    g_skip = torch.empty((b, 3, h // 2, w // 2), device=dev, dtype=torch.float32)
                    N.note(kind="bwd_elementwise", bytes=4.0 * 1.25 * g_rgb.numel(), tag=f"skip grad @{h}")
                    if taps1d is not None:
                        N.check(lib.w2e_skip_grad(N.ptr(g_rgb), N.ptr(g_skip), N.host_floats(taps1d), b * 3, h // 2, w // 2,
                                                  N.stream_ptr()), "skip_grad")
                    else:   # unusual (non-separable) upsample filter: the generic upfirdn2d backward
                        N.check(lib.w2e_upfirdn2d_bwd(N.ptr(g_rgb), N.ptr(g_skip), N.host_floats(taps2d), b * 3, h // 2, w // 2,
                                                      4, 4, 2, 2, 1, 1, 2, 1, 2, 1, N.F32, N.stream_ptr()), "upfirdn2d_bwd")
                    g_rgb = g_skip
                else:
                    g_rgb = None
                continue
            conv = module.conv
            pw = self._tc_weight(conv)
            act = sv.act[idx]
            d = demods[idx]
            rgb_args = (None, None, None)
            if pending is not None:
                gr, ridx = pending
                rpw = layers[ridx][0].conv.packed()
                rgb_args = (gr, rpw.rgb, styles[ridx])
            if idx in sv.blend:
                # the consumers read the BLENDED tensor: pull their gradients back to it (no activation, no demodulation:
                # act_kind NONE), through the blend to the layer's own output, then through the activation
                blended = sv.blend[idx][0]
                g_bl, sums = self._assemble(gxs_next, s_next, blended, rgb_args[0], rgb_args[1], rgb_args[2], None, None,
                                            act_kind=N.ACT_NONE)
                g_act, gm = self._blend_backward(g_bl, act, sv.blend_orig[idx], sv.mask)
                g_mask = gm if g_mask is None else g_mask + gm
                gz, sums2 = self._assemble(g_act, None, act, None, None, None, sv.noise[idx] if kind == "conv" else None,
                                           d if kind == "conv" else None)
                sums = torch.stack([sums[:, 0], sums[:, 1], sums2[:, 2]], dim=1)
            else:
                gz, sums = self._assemble(gxs_next, s_next, act, rgb_args[0], rgb_args[1], rgb_args[2],
                                          sv.noise[idx] if kind == "conv" else None, d if kind == "conv" else None)
            if next_idx is not None:
                gs[next_idx] = sums[:, 0]                       # direct term of the NEXT convolution's style gradient
            if pending is not None:
                gs[pending[1]] = sums[:, 1]
                pending = None
            # demodulation term of this layer's style gradient (appendix C): with G = sum g_y * y (y = demodulated conv
            # output), dL/dd = G / d and gs_demod = 2 s ((-G d^2 / 2) @ wsq) = -s ((G d^2) @ wsq)
            if kind == "conv":
                gd2 = sums[:, 2] * d * d                        # G = sums[:, 2]
                gxs = self._dgrad_plain(gz, pw)
            else:
                h2, w2 = sv.hw[idx]
                h, w = h2 // 2, w2 // 2
                gzu = self._blur_t(gz, conv.blur.kernel, d, (2 * h + 1, 2 * w + 1))
                gd2 = self._rowdot(gzu, sv.z[idx]) * d          # gzu = g_y * d, z = y: sum gzu*z = G d
                gxs = self._dgrad_up(gzu, pw, h, w)
            demod_terms[idx] = gd2 @ pw.wsq
            gxs_next, s_next, next_idx = gxs, styles[idx], idx
        # first layer: its input is the constant
        gs[conv_ids[0]] = self._rowdot(gxs_next, sv.const)
        for idx in conv_ids:
            gs[idx] = torch.addcmul(gs[idx], styles[idx], demod_terms[idx], value=-1.0)
        sv.g_mask = g_mask
        return gs

    def latent_gradient(self, sv, gs, latent_shape):
        """dL/dW+ [B, n_latent, style_dim] from the per-layer style gradients, through the modulation linears
        s_l = w[:, row_l] @ (W_l * scale)^T + b_l (models/stylegan2/model.py:149-159, 238)."""
        plan = self._plan
        # layers that read the same W+ row (a ToRGB and the next block's first convolution) are neighbours in the
        # concatenated modulation weight: one GEMM per row
        gs_all = torch.cat(gs, dim=1)
        per_row = {}
        for (choff, cin, _), row in zip(plan["seg"], sv.rows):
            lo, hi = per_row.get(row, (choff, choff))
            per_row[row] = (min(lo, choff), max(hi, choff + cin))
        zero = None
        cols = []
        for row in range(latent_shape[1]):
            if row in per_row:
                lo, hi = per_row[row]
                cols.append(gs_all[:, lo:hi] @ plan["w_all"][lo:hi])
            else:
                if zero is None:
                    zero = torch.zeros((latent_shape[0], latent_shape[2]), device=gs_all.device, dtype=torch.float32)
                cols.append(zero)
        return torch.stack(cols, dim=1)


class _SynthesisFn(torch.autograd.Function):
    """image = G(latent) on the channels-last bf16 engine, differentiable w.r.t. the W+ latent or the stylespace list."""

    @staticmethod
    @K._amp_fwd
    def forward(ctx, engine, stylespace, noise, blend, mask, *inputs):
        latent = list(inputs) if stylespace else inputs[0]
        with torch.no_grad():
            if blend is None:
                image, style_vector, sv = engine.forward_train(latent, stylespace, noise)
            else:
                image, style_vector, sv = engine.forward_train(latent, stylespace, noise, blend[0], mask, blend[1])
        ctx.engine, ctx.sv, ctx.stylespace = engine, sv, stylespace
        ctx.shapes = [tuple(t.shape) for t in inputs]
        ctx.has_mask = mask is not None
        ctx.mark_non_differentiable(*style_vector)
        return (image, *style_vector)

    @staticmethod
    @torch.autograd.function.once_differentiable
    @K._amp_bwd
    def backward(ctx, g_img, *_):
        engine, sv = ctx.engine, ctx.sv
        with torch.cuda.device(g_img.device):
            gs = engine.backward_train(sv, g_img)
            if ctx.stylespace:
                grads = [g.reshape(shape) for g, shape in zip(gs, ctx.shapes)]
            else:
                grads = [engine.latent_gradient(sv, gs, ctx.shapes[0])]
        g_mask = sv.g_mask if (ctx.has_mask and ctx.needs_input_grad[4]) else None
        ctx.sv = None
        engine._err.publish()
        return (None, None, None, None, g_mask, *grads)


def synthesize_with_grad(engine, latent, stylespace, noise, attention_layer=0, attention_map=None, feature_map=None):
    """image (differentiable w.r.t. the W+ latent / the stylespace codes and, with a region blend, the attention map)
    and the detached post-modulation styles."""
    inputs = list(latent) if stylespace else [latent]
    engine._err.poll()
    blend = (int(attention_layer), list(feature_map)) if attention_layer else None
    with torch.cuda.device(engine.gen.input.input.device):
        out = _SynthesisFn.apply(engine, stylespace, noise, blend, attention_map if attention_layer else None, *inputs)
    return out[0], list(out[1:])
