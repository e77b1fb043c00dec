/* libw2e -- C ABI of the B200-native (sm_100a) StyleGAN2 synthesis hot path of Where2edit.
 *
 * The reference (Big-Brother-Pikachu/Where2edit) is pure Python/PyTorch and has no FFI of its
 * own; the boundary it offers for this path is the set of Python callables
 *   models/stylegan2/op/upfirdn2d.py:11   upfirdn2d(input, kernel, up, down, pad)
 *   models/stylegan2/op/upfirdn2d.py:19   upfirdn2d_native(input, kernel, 10 ints)
 *   models/stylegan2/op/fused_act.py:23   fused_leaky_relu(input, bias, negative_slope, scale)
 *   models/stylegan2/model.py:234         ModulatedConv2d.forward
 *   models/stylegan2/model.py:334         StyledConv.forward
 *   models/stylegan2/model.py:353         ToRGB.forward
 *   attention/attention_model.py:548      region-mask blend inside Generator.forward
 * Each entry point below names the reference lines it replaces.  A Python binding
 * (where2edit_b200/_native.py, ctypes) calls these with raw device pointers and the current
 * CUDA stream; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error (W2E_ERR_*); it never throws and
 *    never allocates device memory; w2e_last_error_string() describes the last failure of the
 *    calling thread.
 *  - all data pointers are DEVICE pointers unless the name says host; tensors are contiguous.
 *  - `stream` is a cudaStream_t passed as void*; work is enqueued, not synchronised.
 *  - dtype: W2E_F32 (float) or W2E_BF16 (__nv_bfloat16) for activation tensors; parameters
 *    (styles, demod, bias, noise, FIR taps) are always float.
 */
#ifndef W2E_H_
#define W2E_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define W2E_OK 0
#define W2E_ERR_INVALID 1      /* bad argument / unsupported shape */
#define W2E_ERR_CUDA 2         /* a CUDA runtime/driver call failed */
#define W2E_ERR_UNSUPPORTED 3  /* valid request this build cannot serve (e.g. not sm_100) */

#define W2E_F32 0
#define W2E_BF16 1
#define W2E_U8 2   /* image outputs only: uint8 by the quantisation of torchvision.utils.save_image(normalize=True,
                      range=(-1, 1)), which the reference applies to its results (attention/run_attention.py:1470, 1535):
                      u8 = trunc(clamp((clamp(v, -1, 1) + 1) / 2 * 255 + 0.5, 0, 255))                              */

#define W2E_ACT_NONE 0
#define W2E_ACT_LRELU 1

int w2e_version(void);
const char* w2e_last_error_string(void);
/* sm count / compute capability of the current device (host query, no stream work). */
int w2e_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- upfirdn2d  (models/stylegan2/op/upfirdn2d.py:19-60) ---------------------------------
 * NCHW planes: x [planes, in_h, in_w] -> y [planes, out_h, out_w], planes = N*C.
 * taps: host pointer to kh*kw floats, row-major, UNFLIPPED (the reference's `kernel`; the
 * convolution flips it, upfirdn2d.py:46).  Geometry = the reference's 8 ints.
 * out_h = (in_h*up_y + py0 + py1 - kh)/down_y + 1 (same for w); the caller allocates y.
 * Negative pads crop.  A separable 4-tap fast path is selected automatically.            */
int w2e_upfirdn2d_fwd(const void* x, void* y, const float* host_taps, int64_t planes, int in_h,
                      int in_w, int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                      int px0, int px1, int py0, int py1, int dtype, void* stream);
/* gradient w.r.t. x of the call above: gx [planes,in_h,in_w] from gy [planes,out_h,out_w]
 * (an upfirdn2d with the flipped kernel and up/down swapped).        */
int w2e_upfirdn2d_bwd(const void* gy, void* gx, const float* host_taps, int64_t planes, int in_h,
                      int in_w, int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                      int px0, int px1, int py0, int py1, int dtype, void* stream);

/* ---- bias + leaky-ReLU * gain  (models/stylegan2/op/fused_act.py:23-39) ------------------
 * x viewed as [outer, C, inner]; y = lrelu(x + bias[c] + noise_w * noise[.., inner]) * scale.
 * 2-D [B,C]: outer=B, inner=1.  3-D [B,L,C] (bias on the last dim, fused_act.py:26-32):
 * outer=B*L, inner=1.  4-D [B,C,H,W]: outer=B, inner=H*W.
 * noise (optional, may be NULL) implements NoiseInjection (models/stylegan2/model.py:279-290):
 * float [noise_batch, inner] with noise_batch in {1, outer}; noise_w is a DEVICE pointer to the
 * scalar NoiseInjection.weight.  bias may be NULL.                                          */
int w2e_bias_act_fwd(const void* x, const float* bias, const float* noise, const float* noise_w,
                     int noise_batch, void* y, int64_t outer, int C, int64_t inner, float slope,
                     float scale, int dtype, void* stream);
/* gx = gy * scale * (y > 0 ? 1 : slope)  (y = forward output).  If gbias != NULL also
 * gbias[c] = sum over outer,inner of gx (deterministic two-pass reduction; `workspace` must
 * hold w2e_bias_act_bwd_workspace(outer, C, inner) bytes).                                   */
int w2e_bias_act_bwd(const void* gy, const void* y, void* gx, float* gbias, void* workspace,
                     int64_t outer, int C, int64_t inner, float slope, float scale, int dtype,
                     void* stream);
int64_t w2e_bias_act_bwd_workspace(int64_t outer, int C, int64_t inner);

/* ---- style -> demodulation  (models/stylegan2/model.py:239-243, shared-weight form) -----
 * demod[b,o] = rsqrt(sum_i style[b,i]^2 * wsq[o,i] + 1e-8), wsq[o,i] = sum_k (scale*W[o,i,k])^2 */
int w2e_style_demod(const float* style, const float* wsq, float* demod, int B, int Cin, int Cout,
                    void* stream);

/* ---- all modulations / all demodulations of one Generator.forward in one launch each -------
 * (EqualLinear of every ModulatedConv2d, models/stylegan2/model.py:130-159 via :238; demod :242).
 * The host concatenates the per-layer tensors once:  w_all [sumCin, D] = modulation.weight*scale,
 * b_all [sumCin] = modulation.bias*lr_mul, wsq_all = the per-conv wsq[Cout_l, Cin_l] back to back.
 * block_layer[blk] names the layer of every 32-channel block (mod) / 8-channel block (demod);
 * meta4 holds per layer {latent row, first channel, Cin, -} (mod) or {s offset, Cin, Cout, wsq offset}
 * (demod); d_off[l] = first demod channel of layer l.  Results are per-layer contiguous blocks:
 * s_all[B*first + b*Cin + c], d_all[B*d_off + b*Cout + o].  latent is [B, rows, D] with the given
 * element strides.                                                                              */
int w2e_style_mod_all(const float* latent, int64_t stride_b, int64_t stride_row, const float* w_all,
                      const float* b_all, const int* block_layer, const int* meta4, float* s_all, int B,
                      int D, int nblocks, void* stream);
int w2e_style_demod_all(const float* s_all, const float* wsq_all, const int* block_layer, const int* meta4,
                        const int* d_off, float* d_all, int B, int nblocks, void* stream);

/* ---- modulated convolution, exact fp32 engine (models/stylegan2/model.py:249-274) ----------
 * Direct convolution on CUDA cores, fp32 FMA, NCHW.  One engine serves forward and dgrad:
 *   y[b,o,oy,ox] = out_scale[b,o] * sum_{c,t} in_scale[b,c] * x[b,c, j*in_stride+dy_t, i*in_stride+dx_t]
 *                                              * w[t][c][o]            (+ fused epilogue)
 * with (oy,ox) = (j*out_stride+py, i*out_stride+px) for (j,i) on a grid of grid_h x grid_w.
 * taps: host array of ntaps*3 ints {dy, dx, weight_slot}; w: float [slots][Cin][Cout].
 * in_scale / out_scale may be NULL (=1).  Epilogue (act=W2E_ACT_LRELU): + noise_w*noise[oy,ox]
 * + bias[o], leaky-ReLU(slope 0.2) * sqrt(2)  == NoiseInjection + FusedLeakyReLU.
 * Forward plain 3x3:  9 taps (dy,dx in -1..1), strides 1.   Forward up x2 (conv_transpose2d,
 * stride 2): four launches, one per output parity class.       */
int w2e_conv_engine_f32(const float* x, const float* w, const float* in_scale,
                        const float* out_scale, const float* bias, const float* noise,
                        const float* noise_w, int noise_batch, float* y, int B, int Cin, int Cout,
                        int in_h, int in_w, int out_h, int out_w, int grid_h, int grid_w,
                        int in_stride, int out_stride, int py, int px, const int* host_taps,
                        int ntaps, int act, int accumulate, void* stream);

/* ---- row dot / scale helper for the modconv backward ----------------
 * rows = B*C planes of `inner` pixels: dot[r] = sum_p a[r,p]*b[r,p]; if prod != NULL,
 * prod[r,p] = a[r,p] * scale[r] (scale may be NULL = 1).                                     */
int w2e_rowdot_f32(const float* a, const float* b, const float* scale, float* prod, float* dot,
                   int64_t rows, int64_t inner, void* stream);
/* The same reduction with each row split into nseg fixed segments (long rows of the high-resolution layers);
 * partial [rows, nseg] is caller-provided workspace, w2e_rowdot_segments() the recommended nseg.  Deterministic:
 * the partial sums are combined in segment order.  inner % 4 == 0, rows <= 65535, 16-byte aligned operands.  */
int64_t w2e_rowdot_segments(int64_t rows, int64_t inner);
int w2e_rowdot_seg_f32(const float* a, const float* b, const float* scale, float* prod, float* dot,
                       float* partial, int64_t rows, int64_t inner, int nseg, void* stream);

/* ---- ToRGB  (models/stylegan2/model.py:353-362) -------------------------------------------
 * rgb[b,o,p] = sum_c x[b,c,p]*style[b,c]*w[o,c] + bias[o] + upfirdn2d(skip, k4*4, up=2, pad=(2,1))
 * x: [B,Cin,H,W] (dtype, NCHW) ; w: float [3,Cin] pre-scaled by 1/sqrt(Cin) ; skip: float
 * [B,3,H/2,W/2] or NULL ; host_taps1d: the 4 taps of the separable skip filter along one axis
 * (already including the gain, e.g. [1,3,3,1]/8*2) ; rgb: float [B,3,H,W].                   */
int w2e_torgb_fwd(const void* x, const float* w, const float* style, const float* bias,
                  const float* skip, const float* host_taps1d, float* rgb, int B, int Cin, int H,
                  int W, int dtype, void* stream);
/* backward of the 1x1 modulated part: t[b,c,p] = sum_o g[b,o,p]*w[o,c];
 * gx = t*style[b,c] ; gstyle[b,c] = sum_p t*x.  (The skip branch's gradient is an
 * upfirdn2d_bwd of g; the bias gradient a plain reduction -- both done by the caller.)      */
int w2e_torgb_bwd(const float* g, const float* x, const float* w, const float* style, float* gx,
                  float* gstyle, int B, int Cin, int H, int W, void* stream);
/* the same with every (sample, channel) plane split into nseg fixed segments (w2e_rowdot_segments(B*Cin, H*W));
 * partial [B*Cin, nseg] is caller-provided workspace; H*W % 4 == 0, 16-byte aligned operands.                */
int w2e_torgb_bwd_seg(const float* g, const float* x, const float* w, const float* style, float* gx,
                      float* gstyle, float* partial, int B, int Cin, int H, int W, int nseg, void* stream);

/* ---- region-mask blend  (attention/attention_model.py:548-549 and siblings) ---------------
 * out = m*edited + (1-m)*orig, m = mask[b,0,floor(y*mh/H),floor(x*mw/W)] (F.interpolate
 * nearest, integer-exact), broadcast over C; the [B,C,H,W] mask is never materialised.      */
int w2e_mask_blend_fwd(const void* edited, const void* orig, const float* mask, void* out, int B,
                       int C, int H, int W, int mh, int mw, int dtype, void* stream);
/* g_edited = m*g ; g_mask[b,0,y',x'] = sum_c sum_{(y,x)->(y',x')} g*(edited-orig)
 * (no gradient to orig: callers compute it under no_grad, run_attention.py:1195-1203).       */
int w2e_mask_blend_bwd(const float* g, const float* edited, const float* orig, const float* mask,
                       float* g_edited, float* g_mask, float* workspace /* B*H*W floats */, int B,
                       int C, int H, int W, int mh, int mw, void* stream);

/* ==== bf16 tensor-core path: channels-last (NHWC) bf16 activations ==========================
 *
 * ---- modulated convolution on tcgen05 tensor cores (models/stylegan2/model.py:249-274) ------
 * Persistent, warp-specialised implicit GEMM (haloed TMA input tile shared by all 9 taps, accumulators in
 * TMEM, the four parity classes of the transposed convolution fused into one pass):
 *   acc[b,y,x,o] = sum_{taps} sum_c xs[b, y+dy, x+dx, c] * w[tap][o][c]                 (fp32 accumulate)
 *   v = acc * out_scale[b,o]                                  (demodulation, model.py:242-243)
 *   act == W2E_ACT_LRELU:  v = lrelu(v + noise_w*noise[y,x] + bias[o], 0.2) * sqrt(2)
 *   out[b,y,x,o] = v ;  out_mod[b,y,x,o] = v * next_scale[b,o]
 *   xs:  bf16 [B,in_h,in_w,Cin]  -- the input ALREADY multiplied by this layer's style (its
 *        producer wrote it through out_mod/next_scale), so all samples share one weight tensor;
 *        reads outside the image are zero (= the convolution padding, done by TMA).
 *   w:   bf16 [9][Cout][Cin], pre-scaled by 1/sqrt(Cin*k*k) (model.py:216-217).
 *   transposed == 0: plain 3x3, padding 1: xs [B,in_h,in_w,Cin] -> out [B,in_h,in_w,Cout].
 *   transposed == 1: conv_transpose2d stride 2: -> out [B,2*in_h+1,2*in_w+1,Cout] (the pre-blur tensor).
 *   out / out_mod: bf16 channels-last; either may be NULL.  error_flag: device int set to 1 if the kernel's
 *   pipeline timed out (never expected; turns a hang into a reportable error).
 * Requires Cin % 32 == 0, Cout % 16 == 0, 16-byte aligned tensors, an sm_100 device.
 *
 * cfg (may be NULL = defaults) carries the PER-CALL tuning / A-B switches; the library keeps no mutable
 * global tuning state.  Results are bit-identical for every setting.                           */
typedef struct w2e_tc2_config {
  int max_ctas;       /* cap on the number of persistent CTAs (0 = one per SM)                               */
  int ts_mode;        /* 1 = shared-memory-staged TMA-store epilogue whenever eligible, 0 = direct stores    */
  int flags;          /* bit 0 no edge-tile tap masking (transposed), bit 1 a single MMA-issuing warp,
                         bit 3 64-column tiles for the transposed conv, bit 4 256-pixel tiles for the 32-channel
                         transposed conv, bit 5 128-pixel tiles (two accumulator sets) for the 64-channel one,
                         bit 6 one weight request per filter tap instead of one per tap row, bit 7 one 256-column
                         tile instead of two 128-column tiles for the 256-channel conv with fused ToRGB, bit 8 no
                         per-class accumulator hand-over in the transposed conv, bit 9 no second planning pass with the
                         weight ring when the staged epilogue does not fit next to resident weights              */
  int cluster_log2;   /* weight-ring kernels as clusters of 2^n CTAs with TMA-multicast weight blocks (0..3)  */
  void* timeline;     /* device long long[64][8] or NULL: clock64 stamps of CTA 0's first 64 tiles
                         (tools/tc2_timeline.py)                                                              */
} w2e_tc2_config;

int w2e_modconv_tc_supported(void);
int w2e_modconv_tc2(const void* xs, const void* w, const float* out_scale, const float* bias,
                    const float* noise, const float* noise_w, int noise_batch, const float* next_scale,
                    void* out, void* out_mod, int* error_flag, int B, int Cin, int Cout, int in_h,
                    int in_w, int transposed, int act, const w2e_tc2_config* cfg, void* stream);
/* Plain 3x3 variant with the following ToRGB (models/stylegan2/model.py:353-362) fused into the
 * epilogue: every epilogue thread holds one pixel's full channel row, so
 *   rgb[b,o,y,x] = sum_c act[b,y,x,c]*rgb_style[b,c]*rgb_w[o,c] + rgb_bias[o] + upsample2(rgb_skip)[b,o,y,x]
 * is produced without re-reading the activation; out and out_mod may then both be NULL.
 * rgb_w: float [3,Cout] pre-scaled by 1/sqrt(Cout); rgb_skip: float [B,3,in_h/2,in_w/2] or NULL;
 * host_taps1d: the 4 taps of the separable skip filter (with gain); rgb: [B,3,in_h,in_w] of rgb_dtype
 * (W2E_F32, W2E_BF16 or W2E_U8: the image in the dtype the caller wants, straight from the epilogue).
 * Needs Cout <= 512 (the rgb image must be fp32 and zero-initialised when Cout >= 256: two channel blocks add
 * into it) and in_h > 16.                                                                        */
int w2e_modconv_tc2_rgb(const void* xs, const void* w, const float* out_scale, const float* bias,
                        const float* noise, const float* noise_w, int noise_batch,
                        const float* next_scale, void* out, void* out_mod, int* error_flag, int B,
                        int Cin, int Cout, int in_h, int in_w, int act, const float* rgb_w,
                        const float* rgb_style, const float* rgb_bias, const float* rgb_skip,
                        const float* host_taps1d, void* rgb, int rgb_dtype, const w2e_tc2_config* cfg,
                        void* stream);
/* x-pair mode of the RGB-only 32 -> 32 channel layer (the last layer of the 1024^2 generator): the same kernel on PIXEL
 * PAIRS, so that every tcgen05.mma has N = 64 instead of 32 (48 instead of 2 x 44 cycles per 16 K).  xs: the ordinary
 * channels-last input [B,H,W,32] (read as [B,H,W/2,64]); w_pair: bf16 [9][64][64],
 * w_pair[ky*3+dj+1][a*32+o][b*32+c] = w[ky*3+kx][o][c] with kx = 2*dj + b - a + 1 (zero outside 0..2);
 * out_scale [B,32], bias [32], rgb_w [3,32], rgb_style [B,32]: the ordinary 32-channel arrays.  H, W in pixels,
 * W a multiple of 16, H even.  out: bf16 [B,H,W,32] activation or NULL (no modulated copy in this mode); rgb, rgb_dtype as above.
 * models/stylegan2/model.py:234-276, 306-340 (StyledConv) + 343-362 (ToRGB).                                          */
int w2e_modconv_tc2_rgb_pair(const void* xs, const void* w_pair, const float* out_scale, const float* bias,
                             const float* noise, const float* noise_w, int noise_batch, void* out, int* error_flag, int B,
                             int H, int W, int act, const float* rgb_w, const float* rgb_style,
                             const float* rgb_bias, const float* rgb_skip, const float* host_taps1d, void* rgb,
                             int rgb_dtype, const w2e_tc2_config* cfg, void* stream);
/* Plain 32 -> 32 channel 3x3 convolution on pixel pairs, no epilogue terms (the dgrad of the last layer, model.py:249-274
 * under autograd): xs bf16 [B,H,W,32], w_pair bf16 [9][64][64] built like w2e_modconv_tc2_rgb_pair's from the [9][32][32]
 * operand, out bf16 [B,H,W,32].  W a multiple of 16, H > 16.                                                          */
int w2e_modconv_tc2_pair(const void* xs, const void* w_pair, void* out, int* error_flag, int B, int H, int W,
                         const w2e_tc2_config* cfg, void* stream);
/* Fused up-convolution + Blur + NoiseInjection + FusedLeakyReLU (models/stylegan2/model.py:249-260,
 * 279-290; op/fused_act.py:23-39): the transposed x2 modulated convolution of w2e_modconv_tc2 whose
 * (2h+1)^2 pre-blur result stays in shared memory and is filtered by the separable 4x4 FIR host_taps[16]
 * (upfirdn2d, pad (1,1)) inside the same kernel.  out / out_mod: [B, 2h, 2w, Cout] bf16 channels-last (either
 * may be null), noise: [noise_batch, 2h*2w] fp32 or null.  Returns W2E_ERR_UNSUPPORTED (and launches
 * nothing) for shapes the fused kernel does not cover (Cout > 64, fewer than 17 input rows, a
 * non-separable kernel): the caller then runs w2e_modconv_tc2 + w2e_blur_act_nhwc.                 */
int w2e_modconv_tc2_upblur(const void* xs, const void* w, const float* out_scale, const float* host_taps,
                           const float* bias, const float* noise, const float* noise_w, int noise_batch,
                           const float* next_scale, void* out, void* out_mod, int* error_flag, int B,
                           int Cin, int Cout, int in_h, int in_w, int act, const w2e_tc2_config* cfg,
                           void* stream);

/* Plain 3x3 convolution of w2e_modconv_tc2 over a STRIDED VIEW of a channels-last bf16 tensor (strides in elements,
 * multiples of 8) using only the filter taps whose bit is set in tap_mask (bit ky*3+kx).  Used by the dgrad of the
 * transposed x2 convolution (autograd of models/stylegan2/model.py:249-259): four launches, one per output-parity class
 * of the (2h+1)^2 upstream gradient.  out is [B,out_h,out_w,Cout] with out_h <= in_h, out_w <= in_w (the rest of the
 * convolution grid is clipped); accumulate != 0 ADDS the result to out (TMA reduce-add, bf16) so that the four class
 * launches accumulate in place; a clipped or accumulating output needs more than 16 rows (else W2E_ERR_UNSUPPORTED:
 * run with out_h = in_h, out_w = in_w and add the parts with w2e_sum4_nhwc).  No noise / bias / activation.        */
int w2e_modconv_tc2_view(const void* xs, const void* w, const float* out_scale, const float* next_scale,
                         void* out, void* out_mod, int* error_flag, int B, int Cin, int Cout, int in_h, int in_w,
                         int64_t stride_x, int64_t stride_y, int64_t stride_b, int tap_mask, int out_h, int out_w,
                         int accumulate, const w2e_tc2_config* cfg, void* stream);
/* Fused dgrad of the transposed x2 convolution (conv_transpose2d(stride 2) of models/stylegan2/model.py:249-258; autograd
 * reaches it from attention/run_attention.py:1419 and mapper/training/coach.py:91): gx [B,h,w,Cout] from the upstream
 * gradient gz [B,2h+1,2w+1,Cin] (channels-last bf16) in one launch, gx[j,i] = sum_{ky,kx} gz[2j+ky,2i+kx] . w[ky*3+kx] with
 * w bf16 [9][Cout][Cin]; out_scale [B,Cout] or NULL.  W2E_ERR_UNSUPPORTED when the weights do not fit in shared memory
 * next to the four parity-class tiles of gz (use w2e_modconv_tc2_view per class then).                                 */
int w2e_modconv_tc2_dgrad_up(const void* gz, const void* w, const float* out_scale, void* gx, int* error_flag, int B,
                             int Cin, int Cout, int h, int w_, const w2e_tc2_config* cfg, void* stream);
/* K-loop form of the same dgrad for the layers whose weights do not fit in shared memory next to four class tiles: the K
 * loop of one accumulator walks the four parity classes of gz [B,2h+1,2w+1,K] one after the other (class c = py * 2 + px
 * through its strided view, with the taps `tap_masks >> 9 c & 0x1ff` it feeds), weights streamed through the ring.
 * w4: bf16 [9][Cout][4 K] = the per-class weights of w2e_modconv_tc2_view concatenated along K in class order.
 * gx [B,h,w,Cout] is written once.  K % 64 == 0, h >= 16; W2E_ERR_UNSUPPORTED otherwise (use the per-class launches).  */
int w2e_modconv_tc2_dgrad_up_k(const void* gz, const void* w4, long long tap_masks, const float* out_scale, void* gx,
                               int* error_flag, int B, int K, int Cout, int h, int w_, const w2e_tc2_config* cfg,
                               void* stream);

/* tf32 mode (north_star (1): "bf16 and tf32 modes"): the same kernel with fp32 tensors in HBM (channels-last
 * activations xs / out / out_mod, weights [9][Cout][Cin]) read by tcgen05.mma kind::tf32 (10-bit mantissa operands,
 * fp32 accumulate); direct-store epilogue, no fused ToRGB.  Requires Cin % 16 == 0, Cout % 16 == 0.           */
int w2e_modconv_tc2_tf32(const float* xs, const float* w, const float* out_scale, const float* bias,
                         const float* noise, const float* noise_w, int noise_batch, const float* next_scale,
                         float* out, float* out_mod, int* error_flag, int B, int Cin, int Cout, int in_h,
                         int in_w, int transposed, int act, const w2e_tc2_config* cfg, void* stream);

/* ---- grouped region-attention heads of the cluster-style mapper ----
 * attention/run_attention.py:803-806, 829-841: per captured generator feature map one 1x1 StyledConv(C_h -> 32)
 * (models/stylegan2/model.py:234-290, 306-340, stylespace input) followed by F.interpolate(nearest) to size x size and
 * the concatenation of all results.  ONE launch for all heads, evaluated only at the pixels that survive the resize:
 *   out[b, 32 h + o, Y, X] = lrelu(d[b,o] * sum_c weight_h[o,c] style_h[b,c] feat_h[b,c,y,x] + nw_h noise_h + bias_h[o]) * sqrt2
 * feat / weight / style / bias / noise_w / noise: HOST arrays of nheads device pointers -- feat_h float [B,chans[h],res[h],
 * res[h]], weight_h [32,chans[h]] (equalised-lr scale folded in), style_h [B,chans[h]], bias_h [32], noise_w_h a device
 * scalar, noise_h [B,r,r] with r = min(res[h], size) or NULL (noise / noise_w may be NULL altogether).
 * out: float [B, 32 nheads, size, size]; demod: float [nheads, B, 32] (written; the backward reads it).  nheads <= 24.   */
int w2e_attn_heads_fwd(int nheads, const float* const* feat, const int* chans, const int* res,
                       const float* const* weight, const float* const* style, const float* const* bias,
                       const float* const* noise_w, const float* const* noise, float* out, float* demod, int B, int size,
                       void* stream);
/* Backward of the above (the reference trains the heads, run_attention.py:725-735; the feature maps come from the frozen
 * generator under no_grad, :1196-1203): gout / out [B, 32 nheads, size, size]; gacc: scratch of the same size;
 * red: float [nheads, B, 3, 32] (written): [.,.,0,o] = d bias_h[o], [.,.,1,o] internal, [.,.,2,0] = d noise_w_h, per sample
 * (the caller sums over B); dW / ds: HOST arrays of device pointers to the [32,chans[h]] / [B,chans[h]] gradients of
 * weight_h / style_h (written, demodulation terms included).  Fixed-order reductions, no atomics.                       */
int w2e_attn_heads_bwd(int nheads, const float* const* feat, const int* chans, const int* res,
                       const float* const* weight, const float* const* style, const float* const* bias,
                       const float* const* noise_w, const float* const* noise, const float* gout, const float* out,
                       const float* demod, float* gacc, float* red, float* const* dW, float* const* ds, int B, int size,
                       void* stream);

/* ---- layout transforms (dtype = W2E_BF16 or W2E_F32: element type of the CHANNELS-LAST tensor) ----
 * x fp32 [Bx,C,HW] (Bx == 1 broadcasts, e.g. ConstantInput, model.py:293-303) -> y bf16 [B,HW,C],
 * multiplied by style[b,c] when style != NULL.                                                */
int w2e_nchw_to_nhwc_mod(const float* x, const float* style, void* y, int B, int Bx, int C,
                         int64_t HW, int dtype, void* stream);
/* x bf16 [B,HW,C] -> y fp32 [B,C,HW]  (feature capture, attention_model.py:542-543).          */
int w2e_nhwc_to_nchw_f32(const void* x, float* y, int B, int C, int64_t HW, int dtype, void* stream);
/* Pieces of the tensor-core dgrad of the transposed x2 convolution (autograd of model.py:249-259): the
 * output-parity class (py, px) of a fp32 NCHW gradient [B,C,H,W] as bf16 channels-last
 * y[b,j,i,c] = x[b,c,2j+py,2i+px]*scale[b,c] ([B,(H-py+1)/2,(W-px+1)/2,C]); and the sum of the four
 * per-class convolution results y_pq [B,h+1-p,w+1-q,C] over their common h x w region as fp32 NCHW.   */
int w2e_nchw_class_to_nhwc_mod(const float* x, const float* scale, void* y, int B, int C, int H, int W, int py,
                               int px, int dtype, void* stream);
int w2e_nhwc_sum4_to_nchw_f32(const void* y00, const void* y01, const void* y10, const void* y11, float* out,
                              int B, int C, int h, int w, int dtype, void* stream);

/* ---- channels-last bf16 backward (BASELINE config 4; csrc/bwd_nhwc.cu; math: SURVEY.md appendix C) -------------
 * grad_assemble: one pass over the output a [B,HW,C] of a styled layer that turns the gradient w.r.t. the modulated
 * input of the NEXT convolution (gxs, style s_next; either may be NULL) and the gradient of a ToRGB reading a
 * (g_rgb [B,3,HW] fp32 with w_rgb [3,C], s_rgb [B,C]; or NULL) into gz = dL/d(pre-activation) * demod (bf16, the
 * operand of this layer's dgrad; may be NULL), and reduces sums[b,0,c] = sum_p gxs*a (direct term of dL/ds_next),
 * sums[b,1,c] = sum_p (sum_o g_rgb[o] w_rgb[o,c]) * a (dL/ds_rgb) and, when demod != NULL,
 * sums[b,2,c] = sum_p dL/dy * (y - noise_w*noise - bias) (= demod * dL/ddemod).  act_kind: W2E_ACT_LRELU inverts
 * a = lrelu(y)*sqrt2 (op/fused_act.py:23-39).  workspace: w2e_grad_assemble_workspace(B,HW,C) floats.            */
int64_t w2e_grad_assemble_workspace(int B, int64_t HW, int C);
int w2e_grad_assemble_nhwc(const void* gxs, const float* s_next, const void* act, const float* g_rgb,
                           const float* w_rgb, const float* s_rgb, const float* noise, const float* noise_w,
                           int noise_batch, const float* bias, const float* demod, int act_kind, void* gz,
                           float* sums, float* workspace, int B, int64_t HW, int C, void* stream);
/* dot[b,c] = sum_p a[b,p,c] * b[b or 0,p,c] (bf16 channels-last; b_batch == 1 broadcasts b); same workspace.   */
int w2e_rowdot_nhwc(const void* a, const void* b, int b_batch, float* dot, float* workspace, int B, int64_t HW,
                    int C, void* stream);
/* out [B,h,w,C] = y00 [B,h+1,w+1,C] + y01 [B,h+1,w,C] + y10 [B,h,w+1,C] + y11 [B,h,w,C] over the h x w region.   */
int w2e_sum4_nhwc(const void* y00, const void* y01, const void* y10, const void* y11, void* out, int B, int h,
                  int w, int C, void* stream);

/* Gradient of the ToRGB skip upsample (upfirdn2d(skip, outer(k1,k1), up 2, pad (2,1)), model.py:31-49,358) for a
 * separable 4-tap kernel: g [planes,2h,2w] fp32 -> g_skip [planes,h,w]; host_taps1d = k1 (4 floats, with gain).  */
int w2e_skip_grad(const float* g, float* g_skip, const float* host_taps1d, int64_t planes, int h, int w,
                  void* stream);

/* ---- Blur + NoiseInjection + FusedLeakyReLU, channels-last (model.py:200-206,260,279-290) ---
 * z bf16 [B,in_h,in_w,C] --(4x4 separable FIR `host_taps` [16], unflipped; pad py0/px0 before)-->
 * [B,out_h,out_w,C]; then the same epilogue as w2e_modconv_tc2 (no demodulation).
 * variant: 0 = kernel with a run-time tile shape, 1 = templated channel tile with a predicate-free interior
 * body (C >= 32); both give bit-identical results.                                                */
int w2e_blur_act_nhwc(const void* z, const float* host_taps, const float* bias, const float* noise,
                      const float* noise_w, int noise_batch, const float* next_scale, void* out,
                      void* out_mod, int B, int C, int in_h, int in_w, int py0, int px0, int out_h,
                      int out_w, int act, int variant, void* stream);

/* ---- ToRGB, channels-last input (model.py:353-362): as w2e_torgb_fwd with x bf16 [B,H,W,C]. */
int w2e_torgb_nhwc(const void* x, const float* w, const float* style, const float* bias,
                   const float* skip, const float* host_taps1d, float* rgb, int B, int C, int H,
                   int W, void* stream);

/* ---- region-mask blend, channels-last bf16 (attention_model.py:548-549) ---------------------
 * out = m*edited + (1-m)*orig ; out_mod = out*next_scale[b,c].                                */
int w2e_blend_nhwc(const void* edited, const void* orig, const float* mask, const float* next_scale,
                   void* out, void* out_mod, int B, int C, int H, int W, int mh, int mw,
                   void* stream);

/* ---- region-mask construction of the cluster-style mapper -------
 * The step just before the blended Generator.forward: attention/run_attention.py:775-794 (cluster
 * assignment) and :852-884 (per-cluster mean attention, losses, threshold, gaussian blur).
 *
 * w2e_cluster_assign: feature f32 NCHW [B,C,h,h]; centres f32 [K, C + 2*pos_channels] (feature
 *   channels, then pos_channels copies of the x position, then of the y position, positions =
 *   i*2/(h-1)-1); ids int64 [B,S,S] = b*K + argmin_k squared distance (utils.py:244-263), nearest-
 *   resized from h to S.  low_ws: int32 workspace [B,h,h].
 * w2e_region_mask_fwd: each f32 [B,S,S] (the sigmoid attention), ids as above; host_taps25 = the
 *   5x5 blur kernel (HOST pointer, row-major).  Writes final_map [B,S,S] (the attention_map, add the
 *   channel axis in the caller), same [B,S,S] (per-cluster means painted back), stats [B,K,2]
 *   (mean, count), losses[2] = {loss_reg, loss_tv}; parts_ws: f32 workspace [B,2].  Pixels whose id
 *   lies outside [b*K, (b+1)*K) keep the value 1 (the reference's torch.ones initialisation).
 * w2e_region_mask_bwd: g_each = d/d each of <final_map, g_final> + g_losses[0]*loss_reg +
 *   g_losses[1]*loss_tv (g_final and g_losses are device pointers and may be null = zero).      */
int w2e_cluster_assign(const float* feature, const float* centres, int* low_ws, int64_t* ids, int B, int C,
                       int h, int K, int pos_channels, int S, void* stream);
int w2e_region_mask_fwd(const float* each, const int64_t* ids, const float* host_taps25, float* final_map,
                        float* same, float* stats, float* parts_ws, float* losses, int B, int S, int K,
                        float threshold, float margin, void* stream);
int w2e_region_mask_bwd(const float* g_final, const float* g_losses, const float* each, const int64_t* ids,
                        const float* same, const float* stats, const float* host_taps25, float* g_each, int B,
                        int S, int K, float margin, void* stream);

/* ---- CLIP / VGG pre-resample ------------------------------------
 * y = AvgPool2d(pool)(Upsample(scale_factor=up, mode="nearest")(x))  (criteria/clip_loss.py:10-14,
 * criteria/perceptual_loss.py:12-17; up = 7, pool = stylegan_size/32 -> 224x224) without the
 * upsampled intermediate.  x f32 [planes,H,W] -> y f32 [planes, H*up/pool, W*up/pool] (floor);
 * bwd: gy -> gx, the exact adjoint.                                                             */
int w2e_box_resample_fwd(const float* x, float* y, int64_t planes, int H, int W, int up, int pool,
                         void* stream);
int w2e_box_resample_bwd(const float* gy, float* gx, int64_t planes, int H, int W, int up, int pool,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* W2E_H_ */
