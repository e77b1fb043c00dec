// Fused elementwise kernels of the synthesis path (all HBM-bound):
//  * ToRGB: 1x1 modulated conv to 3 channels (no demodulation) + bias + polyphase x2 upsample of
//    the running skip image + accumulate -- models/stylegan2/model.py:353-362, 31-49.
//    Algorithmic bytes: B*H*W*(Cin*sizeof(T) + 3*4 (+3*4/4 skip)).
//  * Region-mask blend -- attention/attention_model.py:548-549: the nearest-resized mask is
//    indexed with integer arithmetic and never materialised.  Bytes: 3*B*C*H*W*sizeof(T).
//  * rowdot / style_demod: the small reductions around the modulated convolution.
#include "common.cuh"

namespace w2e {

// ------------------------------------------------------------------------------------------ ToRGB
template <typename T>
__global__ void __launch_bounds__(256)
torgb_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ style,
                 const float* __restrict__ bias, const float* __restrict__ skip, float4 kf, float* __restrict__ rgb,
                 int Cin, int H, int W) {
  extern __shared__ float wm[];  // [3][Cin] per-sample modulated weights
  const int b = blockIdx.y;
  for (int e = threadIdx.x; e < 3 * Cin; e += blockDim.x)
    wm[e] = __ldg(w + e) * __ldg(style + (int64_t)b * Cin + (e % Cin));
  __syncthreads();
  const int64_t HW = (int64_t)H * W;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const T* xp = x + (int64_t)b * Cin * HW + p;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 8
  for (int c = 0; c < Cin; ++c) {
    const float v = to_f32(xp[(int64_t)c * HW]);
    a0 = fmaf(v, wm[c], a0);
    a1 = fmaf(v, wm[Cin + c], a1);
    a2 = fmaf(v, wm[2 * Cin + c], a2);
  }
  float out[3] = {a0, a1, a2};
  if (bias) {
#pragma unroll
    for (int o = 0; o < 3; ++o) out[o] += __ldg(bias + o);
  }
  if (skip) {
    // upfirdn2d(skip, k, up=2, pad=(2,1)): polyphase, 2 taps per axis (flipped taps kf)
    const int oy = (int)(p / W), ox = (int)(p % W);
    const int h = H / 2, wd = W / 2;
    const float kfa[4] = {kf.x, kf.y, kf.z, kf.w};
    const int ya = (oy & 1) ? (oy - 1) / 2 : oy / 2 - 1, xa = (ox & 1) ? (ox - 1) / 2 : ox / 2 - 1;
    const float cy0 = kfa[(oy & 1)], cy1 = kfa[(oy & 1) + 2];
    const float cx0 = kfa[(ox & 1)], cx1 = kfa[(ox & 1) + 2];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      const float* sp = skip + ((int64_t)b * 3 + o) * h * wd;
      float acc = 0.f;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int iy = ya + dy;
        if (iy < 0 || iy >= h) continue;
        float row = 0.f;
        if (xa >= 0) row = cx0 * __ldg(sp + (int64_t)iy * wd + xa);
        if (xa + 1 < wd) row = fmaf(cx1, __ldg(sp + (int64_t)iy * wd + xa + 1), row);
        acc = fmaf(dy ? cy1 : cy0, row, acc);
      }
      out[o] += acc;
    }
  }
#pragma unroll
  for (int o = 0; o < 3; ++o) rgb[((int64_t)b * 3 + o) * HW + p] = out[o];
}

// one block per (c, b) plane: gx = t*style, gstyle = sum_p t*x with t = sum_o g[o]*w[o,c]
__global__ void __launch_bounds__(256)
torgb_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ w,
                 const float* __restrict__ style, float* __restrict__ gx, float* __restrict__ gstyle, int Cin,
                 int64_t HW) {
  __shared__ double red[32];
  const int c = blockIdx.x, b = blockIdx.y;
  const float w0 = __ldg(w + c), w1 = __ldg(w + Cin + c), w2 = __ldg(w + 2 * Cin + c);
  const float s = __ldg(style + (int64_t)b * Cin + c);
  const float* gp = g + (int64_t)b * 3 * HW;
  const float* xp = x + ((int64_t)b * Cin + c) * HW;
  float* gxp = gx + ((int64_t)b * Cin + c) * HW;
  double acc = 0.0;
  for (int64_t p = threadIdx.x; p < HW; p += blockDim.x) {
    const float t = fmaf(gp[p], w0, fmaf(gp[HW + p], w1, gp[2 * HW + p] * w2));
    gxp[p] = t * s;
    acc += (double)(t * xp[p]);
  }
  acc = block_sum_d(acc, red);
  if (threadIdx.x == 0) gstyle[(int64_t)b * Cin + c] = (float)acc;
}

// the same with every (c, b) plane split into nseg fixed segments (grid = nseg x Cin x B); the partial sums of
// gstyle are combined in segment order by rowdot_finish_kernel (deterministic)
__global__ void __launch_bounds__(256)
torgb_bwd_seg_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ style, float* __restrict__ gx, float* __restrict__ partial, int Cin,
                     int64_t HW, int64_t seg_len, int nseg) {
  __shared__ double red[32];
  const int sg = blockIdx.x, c = blockIdx.y, b = blockIdx.z;
  const float w0 = __ldg(w + c), w1 = __ldg(w + Cin + c), w2 = __ldg(w + 2 * Cin + c);
  const float s = __ldg(style + (int64_t)b * Cin + c);
  const float* gp = g + (int64_t)b * 3 * HW;
  const float* xp = x + ((int64_t)b * Cin + c) * HW;
  float* gxp = gx + ((int64_t)b * Cin + c) * HW;
  const int64_t p0 = (int64_t)sg * seg_len, p1 = min(HW, p0 + seg_len);
  double acc = 0.0;
  for (int64_t p = p0 + (int64_t)threadIdx.x * 4; p < p1; p += (int64_t)blockDim.x * 4) {
    const float4 g0 = *reinterpret_cast<const float4*>(gp + p);
    const float4 g1 = *reinterpret_cast<const float4*>(gp + HW + p);
    const float4 g2 = *reinterpret_cast<const float4*>(gp + 2 * HW + p);
    const float4 xv = *reinterpret_cast<const float4*>(xp + p);
    float4 t;
    t.x = fmaf(g0.x, w0, fmaf(g1.x, w1, g2.x * w2));
    t.y = fmaf(g0.y, w0, fmaf(g1.y, w1, g2.y * w2));
    t.z = fmaf(g0.z, w0, fmaf(g1.z, w1, g2.z * w2));
    t.w = fmaf(g0.w, w0, fmaf(g1.w, w1, g2.w * w2));
    *reinterpret_cast<float4*>(gxp + p) = make_float4(t.x * s, t.y * s, t.z * s, t.w * s);
    acc += (double)(t.x * xv.x) + (double)(t.y * xv.y) + (double)(t.z * xv.z) + (double)(t.w * xv.w);
  }
  acc = block_sum_d(acc, red);
  if (threadIdx.x == 0) partial[((int64_t)b * Cin + c) * nseg + sg] = (float)acc;
}

// ------------------------------------------------------------------------------------------ blend
template <typename T>
__global__ void __launch_bounds__(256)
mask_blend_fwd_kernel(const T* __restrict__ edited, const T* __restrict__ orig, const float* __restrict__ mask,
                      T* __restrict__ out, int64_t total, int C, int H, int W, int mh, int mw) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(e % W), yy = (int)((e / W) % H);
    const int64_t b = e / ((int64_t)W * H * C);
    const int my = (int)(((int64_t)yy * mh) / H), mx = (int)(((int64_t)xx * mw) / W);
    const float m = __ldg(mask + (b * mh + my) * mw + mx);
    // same operation order as the reference (m*out + (1-m)*orig), no FMA contraction
    const float v = __fadd_rn(__fmul_rn(m, to_f32(edited[e])), __fmul_rn(__fsub_rn(1.f, m), to_f32(orig[e])));
    out[e] = from_f32<T>(v);
  }
}

// thread per pixel: g_edited = m*g over all channels and d[b,y,x] = sum_c g*(edited-orig)
__global__ void __launch_bounds__(256)
mask_blend_bwd_kernel(const float* __restrict__ g, const float* __restrict__ edited, const float* __restrict__ orig,
                      const float* __restrict__ mask, float* __restrict__ g_edited, float* __restrict__ dsum, int B,
                      int C, int H, int W, int mh, int mw) {
  const int64_t HW = (int64_t)H * W;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * HW) return;
  const int64_t b = idx / HW, p = idx % HW;
  const int yy = (int)(p / W), xx = (int)(p % W);
  const int my = (int)(((int64_t)yy * mh) / H), mx = (int)(((int64_t)xx * mw) / W);
  const float m = __ldg(mask + (b * mh + my) * mw + mx);
  float acc = 0.f;
  for (int c = 0; c < C; ++c) {
    const int64_t e = (b * C + c) * HW + p;
    const float gv = g[e];
    g_edited[e] = m * gv;
    acc = fmaf(gv, edited[e] - orig[e], acc);
  }
  dsum[idx] = acc;
}

// thread per mask cell: sum d over the pixels that map to it (a rectangle, fixed order)
__global__ void mask_grad_gather_kernel(const float* __restrict__ dsum, float* __restrict__ g_mask, int B, int H,
                                        int W, int mh, int mw) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * mh * mw) return;
  const int mx = (int)(idx % mw), my = (int)((idx / mw) % mh);
  const int64_t b = idx / ((int64_t)mw * mh);
  // y maps to my  <=>  my*H <= y*mh < (my+1)*H
  const int y0 = (int)(((int64_t)my * H + mh - 1) / mh), y1 = (int)((((int64_t)my + 1) * H + mh - 1) / mh);
  const int x0 = (int)(((int64_t)mx * W + mw - 1) / mw), x1 = (int)((((int64_t)mx + 1) * W + mw - 1) / mw);
  float acc = 0.f;
  for (int y = y0; y < y1 && y < H; ++y)
    for (int x = x0; x < x1 && x < W; ++x) acc += dsum[(b * H + y) * W + x];
  g_mask[idx] = acc;
}

// ------------------------------------------------------------------------------------------ small reductions
// one block per row: dot[r] = sum_p a*b ; prod = a*scale[r]
__global__ void __launch_bounds__(256)
rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ scale,
              float* __restrict__ prod, float* __restrict__ dot, int64_t inner) {
  __shared__ double red[32];
  const int64_t r = blockIdx.x;
  const float* ar = a + r * inner;
  const float* br = b + r * inner;
  const float s = scale ? __ldg(scale + r) : 1.f;
  double acc = 0.0;   // fp64 accumulation: see block_sum_d
  for (int64_t p = threadIdx.x; p < inner; p += blockDim.x) {
    const float av = ar[p];
    acc += (double)(av * br[p]);
    if (prod) prod[r * inner + p] = av * s;
  }
  acc = block_sum_d(acc, red);
  if (threadIdx.x == 0 && dot) dot[r] = (float)acc;
}

// Same reduction with every row split into nseg segments (grid = nseg x rows): long rows of the >= 256^2 layers
// would otherwise run on as few blocks as there are (sample, channel) rows.  Deterministic: fixed segments, the
// partial sums are combined in segment order by rowdot_finish_kernel.
__global__ void __launch_bounds__(256)
rowdot_seg_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ scale,
                  float* __restrict__ prod, float* __restrict__ partial, int64_t inner, int64_t seg_len, int nseg) {
  __shared__ double red[32];
  const int64_t r = blockIdx.y;
  const int sg = blockIdx.x;
  const int64_t p0 = (int64_t)sg * seg_len, p1 = min(inner, p0 + seg_len);
  const float* ar = a + r * inner;
  const float* br = b + r * inner;
  const float s = scale ? __ldg(scale + r) : 1.f;
  double acc = 0.0;
  // seg_len and inner are multiples of 4 and the rows 16-byte aligned (checked by the launcher)
  for (int64_t p = p0 + (int64_t)threadIdx.x * 4; p < p1; p += (int64_t)blockDim.x * 4) {
    const float4 av = *reinterpret_cast<const float4*>(ar + p);
    const float4 bv = *reinterpret_cast<const float4*>(br + p);
    acc += (double)(av.x * bv.x) + (double)(av.y * bv.y) + (double)(av.z * bv.z) + (double)(av.w * bv.w);
    if (prod) *reinterpret_cast<float4*>(prod + r * inner + p) = make_float4(av.x * s, av.y * s, av.z * s, av.w * s);
  }
  acc = block_sum_d(acc, red);
  if (threadIdx.x == 0) partial[r * nseg + sg] = (float)acc;
}

__global__ void __launch_bounds__(256)
rowdot_finish_kernel(const float* __restrict__ partial, float* __restrict__ dot, int64_t rows, int nseg) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double acc = 0.0;
  for (int sg = 0; sg < nseg; ++sg) acc += (double)partial[r * nseg + sg];
  dot[r] = (float)acc;
}

// one warp per (b, o): demod = rsqrt(sum_i style^2 * wsq + 1e-8)   (models/stylegan2/model.py:242)
__global__ void __launch_bounds__(256)
style_demod_kernel(const float* __restrict__ style, const float* __restrict__ wsq, float* __restrict__ demod, int B,
                   int Cin, int Cout) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * Cout) return;
  const int b = warp / Cout, o = warp % Cout;
  float acc = 0.f;
  for (int i = lane; i < Cin; i += 32) {
    const float s = __ldg(style + (int64_t)b * Cin + i);
    acc = fmaf(s * s, __ldg(wsq + (int64_t)o * Cin + i), acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) demod[warp] = rsqrtf(acc + 1e-8f);
}

}  // namespace w2e

using namespace w2e;

extern "C" int w2e_torgb_fwd(const void* x, const float* w, const float* style, const float* bias, const float* skip,
                             const float* host_taps1d, float* rgb, int B, int Cin, int H, int W, int dtype,
                             void* stream) {
  W2E_CHECK_ARG(x && w && style && rgb, "torgb: null pointer");
  W2E_CHECK_ARG(B >= 0 && Cin > 0 && H > 0 && W > 0 && B <= 65535, "torgb: bad shape");
  W2E_CHECK_ARG(skip == nullptr || (host_taps1d && H % 2 == 0 && W % 2 == 0), "torgb: skip needs taps and even H,W");
  W2E_CHECK_ARG((size_t)3 * Cin * sizeof(float) <= 48 * 1024, "torgb: Cin %d too large", Cin);
  if (B == 0) return W2E_OK;
  float4 kf = make_float4(0, 0, 0, 0);
  if (skip) kf = make_float4(host_taps1d[3], host_taps1d[2], host_taps1d[1], host_taps1d[0]);
  dim3 grid((unsigned)ceil_div64((int64_t)H * W, 256), B);
  const size_t smem = (size_t)3 * Cin * sizeof(float);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == W2E_F32)
    torgb_fwd_kernel<float><<<grid, 256, smem, s>>>((const float*)x, w, style, bias, skip, kf, rgb, Cin, H, W);
  else if (dtype == W2E_BF16)
    torgb_fwd_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>((const __nv_bfloat16*)x, w, style, bias, skip, kf, rgb,
                                                            Cin, H, W);
  else
    return set_error(W2E_ERR_INVALID, "torgb: unknown dtype %d", dtype);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_torgb_bwd(const float* g, const float* x, const float* w, const float* style, float* gx,
                             float* gstyle, int B, int Cin, int H, int W, void* stream) {
  W2E_CHECK_ARG(g && x && w && style && gx && gstyle, "torgb_bwd: null pointer");
  W2E_CHECK_ARG(B >= 0 && Cin > 0 && H > 0 && W > 0 && B <= 65535, "torgb_bwd: bad shape");
  if (B == 0) return W2E_OK;
  torgb_bwd_kernel<<<dim3(Cin, B), 256, 0, (cudaStream_t)stream>>>(g, x, w, style, gx, gstyle, Cin, (int64_t)H * W);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_torgb_bwd_seg(const float* g, const float* x, const float* w, const float* style, float* gx,
                                 float* gstyle, float* partial, int B, int Cin, int H, int W, int nseg, void* stream) {
  W2E_CHECK_ARG(g && x && w && style && gx && gstyle && partial, "torgb_bwd_seg: null pointer");
  const int64_t HW = (int64_t)H * W;
  W2E_CHECK_ARG(B >= 0 && Cin > 0 && Cin <= 65535 && H > 0 && W > 0 && B <= 65535 && HW % 4 == 0 && nseg >= 1 && nseg <= 64,
                "torgb_bwd_seg: bad shape (H*W must be a multiple of 4)");
  W2E_CHECK_ARG((((uintptr_t)g | (uintptr_t)x | (uintptr_t)gx) & 15) == 0, "torgb_bwd_seg: operands must be 16-byte aligned");
  if (B == 0) return W2E_OK;
  const int64_t seg_len = ceil_div64(ceil_div64(HW, nseg), 4) * 4;
  cudaStream_t st = (cudaStream_t)stream;
  torgb_bwd_seg_kernel<<<dim3((unsigned)nseg, (unsigned)Cin, (unsigned)B), 256, 0, st>>>(g, x, w, style, gx, partial, Cin, HW,
                                                                                     seg_len, nseg);
  W2E_LAUNCH_OK();
  const int64_t rows = (int64_t)B * Cin;
  rowdot_finish_kernel<<<(unsigned)ceil_div64(rows, 256), 256, 0, st>>>(partial, gstyle, rows, nseg);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_mask_blend_fwd(const void* edited, const void* orig, const float* mask, void* out, int B, int C,
                                  int H, int W, int mh, int mw, int dtype, void* stream) {
  W2E_CHECK_ARG(edited && orig && mask && out, "mask_blend: null pointer");
  W2E_CHECK_ARG(B >= 0 && C > 0 && H > 0 && W > 0 && mh > 0 && mw > 0, "mask_blend: bad shape");
  const int64_t total = (int64_t)B * C * H * W;
  if (total == 0) return W2E_OK;
  const int64_t blocks = ceil_div64(total, 256), cap = (int64_t)sm_count() * 16;
  const unsigned nb = (unsigned)(blocks < cap ? blocks : cap);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == W2E_F32)
    mask_blend_fwd_kernel<float><<<nb, 256, 0, s>>>((const float*)edited, (const float*)orig, mask, (float*)out,
                                                    total, C, H, W, mh, mw);
  else if (dtype == W2E_BF16)
    mask_blend_fwd_kernel<__nv_bfloat16><<<nb, 256, 0, s>>>((const __nv_bfloat16*)edited,
                                                            (const __nv_bfloat16*)orig, mask, (__nv_bfloat16*)out,
                                                            total, C, H, W, mh, mw);
  else
    return set_error(W2E_ERR_INVALID, "mask_blend: unknown dtype %d", dtype);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_mask_blend_bwd(const float* g, const float* edited, const float* orig, const float* mask,
                                  float* g_edited, float* g_mask, float* workspace, int B, int C, int H, int W,
                                  int mh, int mw, void* stream) {
  W2E_CHECK_ARG(g && edited && orig && mask && g_edited && g_mask && workspace, "mask_blend_bwd: null pointer");
  W2E_CHECK_ARG(B >= 0 && C > 0 && H > 0 && W > 0 && mh > 0 && mw > 0, "mask_blend_bwd: bad shape");
  if (B == 0) return W2E_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t px = (int64_t)B * H * W;
  mask_blend_bwd_kernel<<<(unsigned)ceil_div64(px, 256), 256, 0, s>>>(g, edited, orig, mask, g_edited, workspace, B,
                                                                      C, H, W, mh, mw);
  W2E_LAUNCH_OK();
  const int64_t cells = (int64_t)B * mh * mw;
  mask_grad_gather_kernel<<<(unsigned)ceil_div64(cells, 128), 128, 0, s>>>(workspace, g_mask, B, H, W, mh, mw);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_rowdot_f32(const float* a, const float* b, const float* scale, float* prod, float* dot,
                              int64_t rows, int64_t inner, void* stream) {
  W2E_CHECK_ARG(a && b, "rowdot: null pointer");
  W2E_CHECK_ARG(rows >= 0 && inner > 0 && rows < (1ll << 31), "rowdot: bad shape");
  if (rows == 0) return W2E_OK;
  rowdot_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(a, b, scale, prod, dot, inner);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int64_t w2e_rowdot_segments(int64_t rows, int64_t inner) {
  // enough blocks to fill the machine (~8 per SM), at least 8192 elements per segment, at most 64 segments
  if (rows <= 0 || inner <= 0 || inner % 4 != 0) return 1;
  int64_t want = ceil_div64((int64_t)sm_count() * 8, rows);
  int64_t cap = inner / 8192;
  if (want > cap) want = cap;
  if (want > 64) want = 64;
  return want < 1 ? 1 : want;
}

extern "C" int w2e_rowdot_seg_f32(const float* a, const float* b, const float* scale, float* prod, float* dot,
                                  float* partial, int64_t rows, int64_t inner, int nseg, void* stream) {
  W2E_CHECK_ARG(a && b && dot && partial, "rowdot_seg: null pointer");
  W2E_CHECK_ARG(rows >= 0 && rows <= 65535 && inner > 0 && inner % 4 == 0 && nseg >= 1 && nseg <= 64,
                "rowdot_seg: bad shape (inner must be a multiple of 4, rows <= 65535)");
  W2E_CHECK_ARG((((uintptr_t)a | (uintptr_t)b | (uintptr_t)prod) & 15) == 0, "rowdot_seg: operands must be 16-byte aligned");
  if (rows == 0) return W2E_OK;
  int64_t seg_len = ceil_div64(ceil_div64(inner, nseg), 4) * 4;
  cudaStream_t st = (cudaStream_t)stream;
  rowdot_seg_kernel<<<dim3((unsigned)nseg, (unsigned)rows), 256, 0, st>>>(a, b, scale, prod, partial, inner, seg_len, nseg);
  W2E_LAUNCH_OK();
  rowdot_finish_kernel<<<(unsigned)ceil_div64(rows, 256), 256, 0, st>>>(partial, dot, rows, nseg);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_style_demod(const float* style, const float* wsq, float* demod, int B, int Cin, int Cout,
                               void* stream) {
  W2E_CHECK_ARG(style && wsq && demod, "style_demod: null pointer");
  W2E_CHECK_ARG(B >= 0 && Cin > 0 && Cout > 0, "style_demod: bad shape");
  if (B == 0) return W2E_OK;
  const int64_t warps = (int64_t)B * Cout;
  style_demod_kernel<<<(unsigned)ceil_div64(warps * 32, 256), 256, 0, (cudaStream_t)stream>>>(style, wsq, demod, B,
                                                                                             Cin, Cout);
  W2E_LAUNCH_OK();
  return W2E_OK;
}
