"""Drop-in generator modules for the Where2edit hot path.

Same class names, constructor arguments, parameter/buffer names and shapes (state-dict
compatible) and return arities as the reference's models/stylegan2/model.py and its superset
attention/attention_model.py -- but every hot operation runs in libw2e's sm_100a kernels:

  ModulatedConv2d  -> shared-weight modulated convolution: the style scales the ACTIVATION on
                      load and demodulation is an epilogue scale, so one weight tensor serves
                      the whole batch (reference: per-sample weights + grouped conv,
                      models/stylegan2/model.py:239-274)
  StyledConv       -> conv + (blur) + noise + bias + leaky-ReLU*sqrt(2) fused
  ToRGB            -> 1x1 modulated conv + bias + polyphase skip upsample + add, one kernel
  Generator        -> attention/attention_model.py:473-676 semantics incl. feature capture and
                      region-mask blend

Precision modes: "fp32" (exact CUDA-core path, differentiable, <=1e-4 parity target) and
"bf16" (tcgen05 tensor-core path through where2edit_b200.engine, inference).
There is no CPU path: modules can be constructed and (de)serialised on the CPU, but forward
requires CUDA tensors.

The generator is treated as FROZEN, as in every caller of the reference (the optimisers only hold
mapper parameters, attention/run_attention.py:1051, mapper/training/coach.py): gradients flow
to activations, styles / W+ latents and the attention mask, not to generator weights.
"""
import math
import random
import warnings

import torch
from torch import nn
from torch.nn import functional as F

from . import _native as N
from . import functional as K
from .op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d


class PixelNorm(nn.Module):
    """models/stylegan2/model.py:11-17."""

    def __init__(self, dim=1):
        super().__init__()
        self.dim = dim

    def forward(self, input):
        return input * torch.rsqrt(torch.mean(input ** 2, dim=self.dim, keepdim=True) + 1e-8)


def make_kernel(k):
    """models/stylegan2/model.py:20-28."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    k /= k.sum()
    return k


class Upsample(nn.Module):
    """models/stylegan2/model.py:31-49."""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer("kernel", make_kernel(kernel) * (factor ** 2))
        p = self.kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):
    """models/stylegan2/model.py:52-70."""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        self.register_buffer("kernel", make_kernel(kernel))
        p = self.kernel.shape[0] - factor
        self.pad = ((p + 1) // 2, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=1, down=self.factor, pad=self.pad)


class Blur(nn.Module):
    """models/stylegan2/model.py:73-89."""

    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        kernel = make_kernel(kernel)
        if upsample_factor > 1:
            kernel = kernel * (upsample_factor ** 2)
        self.register_buffer("kernel", kernel)
        self.pad = pad

    def forward(self, input):
        return upfirdn2d(input, self.kernel, pad=self.pad)


class EqualConv2d(nn.Module):
    """models/stylegan2/model.py:92-127 -- caller-side utility (mapper heads); plain library conv."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride = stride
        self.padding = padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None

    def forward(self, input):
        return F.conv2d(input, self.weight * self.scale, bias=self.bias, stride=self.stride, padding=self.padding)

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]},"
                f" {self.weight.shape[2]}, stride={self.stride}, padding={self.padding})")


class EqualLinear(nn.Module):
    """models/stylegan2/model.py:130-165: F.linear with the equalised-lr scale; the optional
    'fused_lrelu' activation goes through libw2e's bias+act kernel."""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, input):
        if self.activation:
            out = F.linear(input, self.weight * self.scale)
            return fused_leaky_relu(out, self.bias * self.lr_mul)
        return F.linear(input, self.weight * self.scale, bias=self.bias * self.lr_mul)

    def __repr__(self):
        return f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})"


class ScaledLeakyReLU(nn.Module):
    """models/stylegan2/model.py:168-176."""

    def __init__(self, negative_slope=0.2):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, input):
        return F.leaky_relu(input, negative_slope=self.negative_slope) * math.sqrt(2)


class ModulatedConv2d(nn.Module):
    """models/stylegan2/model.py:179-276.  forward(input, style, input_is_stylespace=False) ->
    (out [B,Cout,H',W'], style [B,1,Cin,1,1])."""

    _warned = False

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.eps = 1e-8
        self.weight_grad = False   # opt-in gradient for self.weight (see forward)
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample
        if upsample:
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + factor - 1, p // 2 + 1), upsample_factor=factor)
        if downsample:
            factor = 2
            p = (len(blur_kernel) - factor) + (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2, p // 2))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate
        self._packed = None  # derived weight layouts, rebuilt when `weight` changes

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, "
                f"upsample={self.upsample}, downsample={self.downsample})")

    def packed(self):
        """Batch-shared weight layouts (K.PackedWeight), cached per weight version/device."""
        w = self.weight
        key = (w.data_ptr(), w._version, str(w.device))
        # A weight that is being TRAINED (weight_grad, e.g. the cluster mapper's attention heads) is repacked on every
        # call: optimisers that write through `.data` (the reference's Ranger, mapper/training/ranger.py:155-162) do not
        # bump the version counter, and a stale layout would silently keep the old weights.
        if self._packed is None or self._packed.key != key or self.weight_grad:
            self._packed = K.PackedWeight(w.detach(), self.scale, key)
        return self._packed

    def styles(self, style, input_is_stylespace=False):
        """[B,style_dim] (or stylespace [B,1,Cin,1,1]) -> s [B,Cin] fp32; model.py:237-238."""
        batch = style.shape[0]
        if input_is_stylespace:
            return style.reshape(batch, self.in_channel).to(torch.float32)
        return self.modulation(style).reshape(batch, self.in_channel)

    def forward(self, input, style, input_is_stylespace=False):
        N.require_cuda(input, style)
        if self.downsample:
            raise NotImplementedError(
                "ModulatedConv2d(downsample=True) is never reached by the synthesis path (only the "
                "discriminator-era code uses it) and is out of scope of where2edit_b200")
        want_wgrad = self.weight_grad and self.weight.requires_grad and torch.is_grad_enabled()
        if (self.training and self.weight.requires_grad and torch.is_grad_enabled() and not self.weight_grad
                and not ModulatedConv2d._warned):
            ModulatedConv2d._warned = True   # once per process
            warnings.warn(
                "where2edit_b200.ModulatedConv2d does not compute gradients for its convolution weight (every "
                "caller of the synthesis path keeps the generator frozen / in eval mode): this module is in "
                "training mode with a trainable weight, whose .grad will stay None (set `module.weight_grad = True` "
                "or call where2edit_b200.enable_weight_gradients(model) to have it computed). Gradients to the "
                "input, the style and the modulation layer are always computed.", RuntimeWarning, stacklevel=2)
        batch = input.shape[0]
        s = self.styles(style, input_is_stylespace)
        pw = self.packed()
        d = K.demod_coefficients(s, pw.wsq) if self.demodulate else None
        out = K.modulated_conv2d(input, s, d, pw, self.kernel_size, self.upsample)
        if want_wgrad:   # opt-in: trainable convolution weight (library weight-gradient, functional.modconv_weight_grad)
            out = K.weight_grad_tap(out, self.weight, input, s, d, self.scale, self.kernel_size, self.upsample)
        if self.upsample:
            out = self.blur(out)
        return out, s.reshape(batch, 1, self.in_channel, 1, 1)


class NoiseInjection(nn.Module):
    """models/stylegan2/model.py:279-290."""

    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            batch, _, height, width = image.shape
            noise = image.new_empty(batch, 1, height, width).normal_()
        return image + self.weight * noise


class ConstantInput(nn.Module):
    """models/stylegan2/model.py:293-303."""

    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


class StyledConv(nn.Module):
    """models/stylegan2/model.py:306-340: ModulatedConv2d -> NoiseInjection -> FusedLeakyReLU;
    forward(input, style, noise=None, input_is_stylespace=False) -> (out, style)."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, noise=None, input_is_stylespace=False):
        out, style = self.conv(input, style, input_is_stylespace=input_is_stylespace)
        if self.conv.weight_grad and torch.is_grad_enabled():
            # trainable StyledConv (opt-in): module-by-module like model.py:337-339, so that noise.weight and
            # activate.bias receive gradients too (the fused epilogue treats them as constants)
            return self.activate(self.noise(out, noise=noise)), style
        if noise is None:  # model.py:286-288: fresh per-sample noise
            noise = out.new_empty(out.shape[0], 1, out.shape[2], out.shape[3]).normal_()
        out = K.noise_bias_act(out, self.activate.bias, noise, self.noise.weight,
                               self.activate.negative_slope, self.activate.scale)
        return out, style


class ToRGB(nn.Module):
    """models/stylegan2/model.py:343-362; forward(input, style, skip=None, input_is_stylespace=False)."""

    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, input, style, skip=None, input_is_stylespace=False):
        N.require_cuda(input, style)
        batch = input.shape[0]
        s = self.conv.styles(style, input_is_stylespace)
        pw = self.conv.packed()
        up_kernel = self.upsample.kernel if skip is not None else None
        out = K.to_rgb(input, s, pw, self.bias, skip, up_kernel)
        return out, s.reshape(batch, 1, self.conv.in_channel, 1, 1)


class Generator(nn.Module):
    """attention/attention_model.py:365-676 (superset of models/stylegan2/model.py:365-574)."""

    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01,
                 precision="fp32"):
        super().__init__()
        self.size = size
        self.style_dim = style_dim
        layers = [PixelNorm()]
        for _ in range(n_mlp):
            layers.append(EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation="fused_lrelu"))
        self.style = nn.Sequential(*layers)
        self.channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier,
                         128: 128 * channel_multiplier, 256: 64 * channel_multiplier,
                         512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}
        self.input = ConstantInput(self.channels[4])
        self.conv1 = StyledConv(self.channels[4], self.channels[4], 3, style_dim, blur_kernel=blur_kernel)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        in_channel = self.channels[4]
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            self.noises.register_buffer(f"noise_{layer_idx}", torch.randn(1, 1, 2 ** res, 2 ** res))
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True, blur_kernel=blur_kernel))
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(out_channel, style_dim))
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - 2
        self.precision = "fp32"
        self._engine = None
        self.image_dtype = torch.float32   # dtype of the returned image in bf16 no-grad mode (set_image_output)
        self.image_out = None              # optional caller-owned buffer the last layer writes the image into
        self.bf16_backward = "engine"      # "modules": keep the fp32 module path under autograd even when the engine applies
        self.set_precision(precision)

    # ------------------------------------------------------------------ helpers kept from the reference
    def make_noise(self):
        device = self.input.input.device
        noises = [torch.randn(1, 1, 2 ** 2, 2 ** 2, device=device)]
        for i in range(3, self.log_size + 1):
            for _ in range(2):
                noises.append(torch.randn(1, 1, 2 ** i, 2 ** i, device=device))
        return noises

    def mean_latent(self, n_latent):
        latent_in = torch.randn(n_latent, self.style_dim, device=self.input.input.device)
        return self.style(latent_in).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    def set_precision(self, precision):
        """'fp32': exact CUDA-core path (<= 1e-4 of the reference; differentiable).
        'bf16': tcgen05 tensor-core path -- the channels-last bf16 engine for no-grad synthesis, bf16-operand
                convolutions (forward and dgrad) inside the fp32 module path under autograd.
        'tf32': the fp32 module path with every 3x3 modulated convolution (forward and dgrad) on tcgen05
                kind::tf32 (fp32 tensors in HBM, 10-bit-mantissa operands, fp32 accumulate): about 8x closer to
                the reference than bf16, for edits whose loss needs it.
        There is no fallback: the tensor-core modes raise if the device is not sm_100."""
        if precision not in ("fp32", "bf16", "tf32"):
            raise ValueError(f"precision must be 'fp32', 'bf16' or 'tf32', got {precision!r}")
        self.precision = precision
        return self

    def set_image_output(self, dtype=torch.float32, buffer=None):
        """bf16 no-grad mode only: the dtype the LAST layer's epilogue writes the image in (torch.float32 like the
        reference; torch.bfloat16: half the bytes for a caller that gathers or copies the images out; torch.uint8: a
        quarter -- the quantisation torchvision.utils.save_image(normalize=True, range=(-1, 1)) applies when the
        reference writes its results, run_attention.py:1470/1535, functional.quantize_u8 is the same arithmetic) and,
        optionally, a caller-owned contiguous [B,3,H,W] buffer of that dtype to write it into (e.g. this rank's slot
        of a peer-mapped all-gather buffer, parallel.PeerGather.own).  The buffer is used when its batch matches."""
        if dtype not in (torch.float32, torch.bfloat16, torch.uint8):
            raise ValueError("image dtype must be torch.float32, torch.bfloat16 or torch.uint8")
        if buffer is not None and (buffer.dtype != dtype or not buffer.is_contiguous()):
            raise ValueError("image buffer must be contiguous and of the requested dtype")
        self.image_dtype, self.image_out = dtype, buffer
        return self

    def _bf16_engine(self):
        from . import train_engine
        if self._engine is None:
            self._engine = train_engine.TrainEngine(self)
        self._engine.image_dtype, self._engine.image_out = self.image_dtype, self.image_out
        return self._engine

    def _train_engine_applies(self, latent, stylespace, noise, blending, want_features, attention_map=None, feature_map=None):
        """The channels-last backward engine covers the latent / style optimisation loops: frozen generator, fixed
        noise buffers, gradient to the W+ latent or the stylespace codes and, with a region blend
        (run_attention.py:1245), to the attention map.  Anything else (feature capture, trainable generator
        parameters, per-sample random noise, original features that require grad) takes the module path."""
        if want_features or any(n is None for n in noise) or self.bf16_backward == "modules":
            return False
        if any(p.requires_grad for p in self.parameters()):
            return False
        if blending and (feature_map is None or any(torch.is_tensor(f) and f.requires_grad for f in feature_map)):
            return False     # (the reference computes the original features under no_grad, run_attention.py:1195-1203)
        inputs = list(latent if stylespace else [latent]) + ([attention_map] if blending else [])
        return any(t.requires_grad for t in inputs)

    def assert_ok(self):
        """Synchronising check that no tensor-core kernel of this generator reported a pipeline timeout (the
        forward itself polls the flag of earlier calls without synchronising and raises when one did)."""
        if self._engine is not None:
            self._engine.assert_ok()
        K.tc_assert_ok()

    def styled_layers(self):
        """[(module, kind)] in execution order: kind in {'conv', 'up', 'rgb'} (26 entries at 1024)."""
        seq = [(self.conv1, "conv"), (self.to_rgb1, "rgb")]
        for j in range(self.log_size - 2):
            seq += [(self.convs[2 * j], "up"), (self.convs[2 * j + 1], "conv"), (self.to_rgbs[j], "rgb")]
        return seq

    # ------------------------------------------------------------------ forward
    def _assemble_latent(self, styles, inject_index, truncation, truncation_latent, input_is_latent,
                         input_is_stylespace):
        """attention/attention_model.py:489-528."""
        if not input_is_latent and not input_is_stylespace:
            styles = [self.style(s) for s in styles]
        if truncation < 1 and not input_is_stylespace:
            styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        if input_is_stylespace:
            return styles[0]
        if len(styles) < 2:
            latent = styles[0]
            if latent.ndim < 3:
                latent = latent.unsqueeze(1).repeat(1, self.n_latent, 1)
            return latent
        if inject_index is None:
            inject_index = random.randint(1, self.n_latent - 1)
        latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
        latent2 = styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)
        return torch.cat([latent, latent2], 1)

    def latent_rows(self, input_is_stylespace):
        """Index into `latent` used by each of the styled layers (W+ rows are shared between a
        ToRGB and the next block's first conv: i += 2; stylespace advances by 3;
        attention_model.py:598-630 / 632-664)."""
        rows = [0, 1]
        i = 2 if input_is_stylespace else 1
        for _ in range(self.log_size - 2):
            rows += [i, i + 1, i + 2]
            i += 3 if input_is_stylespace else 2
        return rows

    def blend_feature_layers(self, attention_layer):
        """Positions of the feature list that a forward blended at `attention_layer` reads (attention_model.py:546-561): the
        map of that layer and, when it is a convolution, the image of the next ToRGB (the `this_layer` carry) -- what to
        pass as `capture_layers` to the forward that produces the original's features."""
        kinds = [k for _, k in self.styled_layers()]
        if not 1 <= attention_layer <= len(kinds):
            raise ValueError(f"attention_layer must be in 1..{len(kinds)}")
        need = [attention_layer - 1]
        if kinds[attention_layer - 1] != "rgb":
            nxt = next((i for i in range(attention_layer, len(kinds)) if kinds[i] == "rgb"), None)
            if nxt is not None:
                need.append(nxt)
        return need

    def forward(self, styles, return_latents=False, return_features=False, inject_index=None, truncation=1,
                truncation_latent=None, input_is_latent=False, input_is_stylespace=False, noise=None,
                randomize_noise=True, attention_layer=0, attention_map=None, feature_map=None, capture_layers=None):
        """attention/attention_model.py:473-676.  One addition to the reference's signature: `capture_layers`, positions of
        the returned feature list to fill when `return_features=True` (None = all 26, as the reference; the others are
        None) -- the blended forward only reads the maps `blend_feature_layers(attention_layer)` names, and bringing
        every layer back as fp32 NCHW costs as much as the forward itself at 1024^2."""
        latent = self._assemble_latent(styles, inject_index, truncation, truncation_latent, input_is_latent,
                                       input_is_stylespace)
        if noise is None:
            if randomize_noise:
                noise = [None] * self.num_layers
            else:
                noise = [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)]

        blending = attention_map is not None
        if self.precision == "bf16" and not torch.is_grad_enabled():
            image, style_vector, captured = self._bf16_engine().run(
                latent, input_is_stylespace, noise, want_features=return_features and not return_latents,
                attention_layer=attention_layer if blending else 0, attention_map=attention_map,
                feature_map=feature_map, capture_layers=capture_layers)
        elif self.precision == "bf16" and self._train_engine_applies(latent, input_is_stylespace, noise, blending,
                                                                      return_features and not return_latents, attention_map,
                                                                      feature_map):
            # under autograd: channels-last bf16 forward + backward engine (train_engine.py)
            from . import train_engine
            image, style_vector = train_engine.synthesize_with_grad(
                self._bf16_engine(), latent, input_is_stylespace, noise, attention_layer if blending else 0, attention_map,
                feature_map)
            captured = []
        else:
            # module path.  precision "bf16" / "tf32": the 3x3 convolutions (forward and dgrad) run on the tensor
            # cores between layout passes, the rest of the differentiable path on the fp32 kernels
            prev = K.TC_AUTOGRAD
            K.TC_AUTOGRAD = self.precision if self.precision in ("bf16", "tf32") else False
            try:
                if K.TC_AUTOGRAD:
                    K.tc_poll()   # a pipeline timeout of an earlier forward / backward raises here (no sync)
                image, style_vector, captured = self._forward_modules(
                    latent, input_is_stylespace, noise, attention_layer if blending else 0, attention_map, feature_map)
                if K.TC_AUTOGRAD:
                    K.tc_publish()
            finally:
                K.TC_AUTOGRAD = prev

        if return_latents:
            return image, latent, style_vector
        if return_features:
            if capture_layers is not None:   # (the module / training paths capture everything: drop what was not asked for)
                keep = set(int(i) for i in capture_layers)
                captured = [c if i in keep else None for i, c in enumerate(captured)]
            return image, latent, style_vector, captured
        return image, None

    def _forward_modules(self, latent, stylespace, noise, attention_layer, attention_map, feature_map):
        rows = self.latent_rows(stylespace)
        pick = (lambda r: latent[r]) if stylespace else (lambda r: latent[:, r])
        batch = (latent[0] if stylespace else latent).shape[0]
        captured, style_vector = [], []
        carry = False
        out = self.input.input.repeat(batch, 1, 1, 1)
        skip = None
        noise_idx = 0
        for layer, ((module, kind), row) in enumerate(zip(self.styled_layers(), rows), start=1):
            if kind == "rgb":
                skip, s = module(out, pick(row), skip, input_is_stylespace=stylespace)
                if attention_layer and (layer == attention_layer or carry):
                    carry = False
                    skip = K.mask_blend(skip, feature_map[layer - 1], attention_map)
                captured.append(skip)
            else:
                out, s = module(out, pick(row), noise=noise[noise_idx], input_is_stylespace=stylespace)
                noise_idx += 1
                if attention_layer and layer == attention_layer:
                    carry = True
                    out = K.mask_blend(out, feature_map[layer - 1], attention_map)
                captured.append(out)
            style_vector.append(s)
        return skip, style_vector, captured
