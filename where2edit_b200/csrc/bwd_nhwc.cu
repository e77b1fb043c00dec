// Channels-last bf16 backward of the synthesis path (BASELINE config 4: the gradient of a CLIP / ID loss w.r.t. the
// styles, consumers attention/run_attention.py:1419 and mapper/training/coach.py:91).  Math: SURVEY.md appendix C.
//
// The forward engine keeps, per styled layer, the UNMODULATED activation a [B,H,W,C] (bf16) next to the modulated
// copy the next convolution consumes.  The backward walks the layers in reverse; between two tensor-core dgrad
// launches (csrc/modconv_tc2.cu, flipped weights) everything elementwise and every per-(sample, channel) reduction of
// a layer boundary is ONE pass over the activation:
//
//   w2e_grad_assemble_nhwc   for the output a of layer l, consumed by the next 3x3 convolution (style s_next) and,
//                            possibly, by a ToRGB (style s_rgb, weight w_rgb):
//        g_a   = gxs_next * s_next  +  sum_o g_rgb[o] * w_rgb[o,c] * s_rgb[b,c]            (gradient w.r.t. a)
//        R1    = sum_p gxs_next * a                 -> direct term of dL/ds_next          (appendix C: gs = sum gxs*x)
//        R2    = sum_p (sum_o g_rgb[o] w_rgb[o,c]) * a   -> dL/ds_rgb
//        g_pre = g_a * lrelu'(a)                    (a = lrelu(y) * sqrt2, so sign(a) = sign(y); fused_act.py:23-39)
//        R3    = sum_p g_pre * (y - noise_w * noise - bias)   -> demodulation gradient: dL/dd = R3 / d
//        gz    = g_pre * d                          (what the dgrad of layer l consumes), bf16
//   w2e_rowdot_nhwc          sum_p a * b per (sample, channel): the two reductions that have no elementwise companion
//   w2e_sum4_nhwc            the four parity-class partial results of the transposed convolution's dgrad, summed
//
// Reductions are deterministic (fixed pixel chunks, partial sums combined in chunk order, fp64 in the second pass) and
// therefore batch-invariant.  Algorithmic bytes: grad_assemble 3 * B*H*W*C*2 (+ 12 B per pixel with a ToRGB),
// rowdot 2 * B*H*W*C*2, sum4 5 * B*h*w*C*2.
#include "common.cuh"

namespace w2e {

namespace {

constexpr int kGaThreads = 256;

struct GaParams {
  const __nv_bfloat16* gxs;     // [B,HW,C] or null
  const float* s_next;          // [B,C] or null (1)
  const __nv_bfloat16* act;     // [B,HW,C]
  const float* g_rgb;           // [B,3,HW] or null
  const float* w_rgb;           // [3,C]
  const float* s_rgb;           // [B,C]
  const float* noise;           // [noise_batch,HW] or null
  const float* noise_w;         // device scalar
  const float* bias;            // [C] or null
  const float* demod;           // [B,C] or null: plain-conv layer (demodulation after the conv, before the activation)
  __nv_bfloat16* gz;            // [B,HW,C] or null
  float* partial;               // [B,nchunks,3,C]
  int64_t HW;
  int C, nchunks, noise_per_sample, lrelu, want_r3;
  int64_t chunk;                // pixels per chunk
};

__device__ __forceinline__ void unpack8(const uint4 v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// grid (nchunks, B); thread = (pixel lane, 8-channel group).  Per-channel constants live in shared memory (read back
// with 128-bit loads) so that the register budget goes to loads in flight: two pixels per thread are fetched before
// either is consumed, three CTAs per SM.  Shared: 9 * C floats of constants + 3 * lanes * C floats for the reduction.
template <bool RGB, bool R3>
__global__ void __launch_bounds__(kGaThreads, RGB ? 2 : 3)
grad_assemble_nhwc_kernel(const GaParams P) {
  extern __shared__ float smem_f[];
  const int C = P.C;
  float* c_sn = smem_f;            // s_next
  float* c_dm = c_sn + C;          // demod
  float* c_bs = c_dm + C;          // bias
  float* c_sr = c_bs + C;          // s_rgb
  float* c_wr = c_sr + C;          // w_rgb [3][C]
  float* red = c_wr + 3 * C + 2 * C;   // (two spare rows keep the block 16-byte aligned for any C % 8 == 0)
  const int groups = C >> 3;
  const int lanes = kGaThreads / groups;
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
  const int b = blockIdx.y;
  const int c0 = g * 8;
  for (int c = threadIdx.x; c < C; c += kGaThreads) {
    c_sn[c] = P.s_next ? __ldg(P.s_next + (int64_t)b * C + c) : 1.f;
    c_dm[c] = P.demod ? __ldg(P.demod + (int64_t)b * C + c) : 1.f;
    c_bs[c] = P.bias ? __ldg(P.bias + c) : 0.f;
    if (RGB) {
      c_sr[c] = __ldg(P.s_rgb + (int64_t)b * C + c);
#pragma unroll
      for (int o = 0; o < 3; ++o) c_wr[o * C + c] = __ldg(P.w_rgb + o * C + c);
    }
  }
  __syncthreads();
  const int64_t p_begin = (int64_t)blockIdx.x * P.chunk;
  const int64_t p_end = min(P.HW, p_begin + P.chunk);
  const float nw = P.noise ? __ldg(P.noise_w) : 0.f;
  const float gain = P.lrelu ? 1.41421356237309515f : 1.f;
  const float neg = P.lrelu ? 0.2f : 1.f;
  const float dpos = gain, dneg = gain * neg, ipos = 1.f / gain, ineg = 1.f / (gain * neg);
  float r1[8], r2[8], r3[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r1[k] = r2[k] = r3[k] = 0.f;
  const int64_t base = (int64_t)b * P.HW;
  const float* nzp = P.noise ? P.noise + (P.noise_per_sample ? (int64_t)b * P.HW : 0) : nullptr;
  const float* grp = RGB ? P.g_rgb + (int64_t)b * 3 * P.HW : nullptr;

  auto consume = [&](int64_t p, const uint4 av, const uint4 gv, const float gr0, const float gr1, const float gr2, const float nz) {
    float a[8], gx[8], out[8];
    unpack8(av, a);
    unpack8(gv, gx);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 sn = *reinterpret_cast<const float4*>(c_sn + c0 + 4 * h);
      const float4 dm = *reinterpret_cast<const float4*>(c_dm + c0 + 4 * h);
      float4 bs = make_float4(0.f, 0.f, 0.f, 0.f), sr = bs, w0 = bs, w1 = bs, w2 = bs;
      if (R3) bs = *reinterpret_cast<const float4*>(c_bs + c0 + 4 * h);
      if (RGB) {
        sr = *reinterpret_cast<const float4*>(c_sr + c0 + 4 * h);
        w0 = *reinterpret_cast<const float4*>(c_wr + c0 + 4 * h);
        w1 = *reinterpret_cast<const float4*>(c_wr + C + c0 + 4 * h);
        w2 = *reinterpret_cast<const float4*>(c_wr + 2 * C + c0 + 4 * h);
      }
      const float snv[4] = {sn.x, sn.y, sn.z, sn.w}, dmv[4] = {dm.x, dm.y, dm.z, dm.w}, bsv[4] = {bs.x, bs.y, bs.z, bs.w};
      const float srv[4] = {sr.x, sr.y, sr.z, sr.w}, w0v[4] = {w0.x, w0.y, w0.z, w0.w}, w1v[4] = {w1.x, w1.y, w1.z, w1.w};
      const float w2v[4] = {w2.x, w2.y, w2.z, w2.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = 4 * h + q;
        float ga = gx[k] * snv[q];
        r1[k] = fmaf(gx[k], a[k], r1[k]);
        if (RGB) {
          const float t = fmaf(gr0, w0v[q], fmaf(gr1, w1v[q], gr2 * w2v[q]));   // unmodulated ToRGB pull-back
          ga = fmaf(t, srv[q], ga);
          r2[k] = fmaf(t, a[k], r2[k]);
        }
        const bool pos = a[k] > 0.f;
        const float gp = ga * (pos ? dpos : dneg);                             // gradient w.r.t. the pre-activation y
        if (R3) r3[k] = fmaf(gp, fmaf(a[k], pos ? ipos : ineg, -nz) - bsv[q], r3[k]);
        out[k] = gp * dmv[q];
      }
    }
    if (P.gz) *reinterpret_cast<uint4*>(P.gz + (base + p) * C + c0) = pack8(out);
  };

  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  int64_t p = p_begin + lane;
  for (; p + lanes < p_end; p += 2 * lanes) {      // two pixels in flight per thread
    const int64_t q = p + lanes;
    const uint4 a0 = *reinterpret_cast<const uint4*>(P.act + (base + p) * C + c0);
    const uint4 a1 = *reinterpret_cast<const uint4*>(P.act + (base + q) * C + c0);
    const uint4 g0 = P.gxs ? *reinterpret_cast<const uint4*>(P.gxs + (base + p) * C + c0) : zero4;
    const uint4 g1 = P.gxs ? *reinterpret_cast<const uint4*>(P.gxs + (base + q) * C + c0) : zero4;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f, y0 = 0.f, y1 = 0.f, y2 = 0.f;
    if (RGB) {
      x0 = __ldg(grp + p); x1 = __ldg(grp + P.HW + p); x2 = __ldg(grp + 2 * P.HW + p);
      y0 = __ldg(grp + q); y1 = __ldg(grp + P.HW + q); y2 = __ldg(grp + 2 * P.HW + q);
    }
    const float n0 = (R3 && nzp) ? nw * __ldg(nzp + p) : 0.f, n1 = (R3 && nzp) ? nw * __ldg(nzp + q) : 0.f;
    consume(p, a0, g0, x0, x1, x2, n0);
    consume(q, a1, g1, y0, y1, y2, n1);
  }
  if (p < p_end) {
    const uint4 a0 = *reinterpret_cast<const uint4*>(P.act + (base + p) * C + c0);
    const uint4 g0 = P.gxs ? *reinterpret_cast<const uint4*>(P.gxs + (base + p) * C + c0) : zero4;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    if (RGB) { x0 = __ldg(grp + p); x1 = __ldg(grp + P.HW + p); x2 = __ldg(grp + 2 * P.HW + p); }
    consume(p, a0, g0, x0, x1, x2, (R3 && nzp) ? nw * __ldg(nzp + p) : 0.f);
  }
  // deterministic reduction over the pixel lanes (lane order), per channel
  float* s1 = red;
  float* s2 = red + (size_t)lanes * C;
  float* s3 = red + (size_t)2 * lanes * C;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s1[lane * C + c0 + k] = r1[k];
    s2[lane * C + c0 + k] = r2[k];
    s3[lane * C + c0 + k] = r3[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += kGaThreads) {
    const int r = i / C, c = i - r * C;
    const float* src = red + (size_t)r * lanes * C + c;
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += src[(size_t)l * C];
    P.partial[(((int64_t)b * P.nchunks + blockIdx.x) * 3 + r) * C + c] = acc;
  }
}

// out[b, r, c] = sum over chunks (in chunk order, fp64) of partial[b, chunk, r, c]
__global__ void __launch_bounds__(256)
reduce_chunks_kernel(const float* __restrict__ partial, float* __restrict__ out, int nchunks, int rc, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over B * rc
  if (i >= total) return;
  const int64_t b = i / rc, j = i - b * rc;
  double acc = 0.0;
  for (int k = 0; k < nchunks; ++k) acc += (double)partial[(b * nchunks + k) * rc + j];
  out[i] = (float)acc;
}

// partial[b, chunk, c] = sum_p a * b over the chunk; bb broadcasts over the batch when b_batch == 1
__global__ void __launch_bounds__(kGaThreads)
rowdot_nhwc_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ bb, float* __restrict__ partial,
                   int64_t HW, int C, int nchunks, int64_t chunk, int b_batch) {
  extern __shared__ float red[];
  const int groups = C >> 3;
  const int lanes = kGaThreads / groups;
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
  const int b = blockIdx.y, c0 = g * 8;
  const int64_t p_begin = (int64_t)blockIdx.x * chunk, p_end = min(HW, p_begin + chunk);
  float r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = 0.f;
  for (int64_t p = p_begin + lane; p < p_end; p += lanes) {
    float x[8], y[8];
    unpack8(*reinterpret_cast<const uint4*>(a + ((int64_t)b * HW + p) * C + c0), x);
    unpack8(*reinterpret_cast<const uint4*>(bb + ((int64_t)(b_batch == 1 ? 0 : b) * HW + p) * C + c0), y);
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = fmaf(x[k], y[k], r[k]);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[lane * C + c0 + k] = r[k];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kGaThreads) {
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += red[(size_t)l * C + c];
    partial[((int64_t)b * nchunks + blockIdx.x) * C + c] = acc;
  }
}

// out[b,j,i,c] = y00[b,j,i,c] + y01 + y10 + y11 over the common h x w region; y_pq is [B, h+1-p, w+1-q, C]
__global__ void __launch_bounds__(256)
sum4_nhwc_kernel(const __nv_bfloat16* __restrict__ y00, const __nv_bfloat16* __restrict__ y01,
                 const __nv_bfloat16* __restrict__ y10, const __nv_bfloat16* __restrict__ y11,
                 __nv_bfloat16* __restrict__ out, int h, int w, int C, int64_t total) {
  const int groups = C >> 3;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(t % groups);
    int64_t pix = t / groups;
    const int i = (int)(pix % w);
    pix /= w;
    const int j = (int)(pix % h);
    const int64_t b = pix / h;
    const int c0 = g * 8;
    float f0[8], f1[8], f2[8], f3[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(y00 + ((b * (h + 1) + j) * (int64_t)(w + 1) + i) * C + c0), f0);
    unpack8(*reinterpret_cast<const uint4*>(y01 + ((b * (h + 1) + j) * (int64_t)w + i) * C + c0), f1);
    unpack8(*reinterpret_cast<const uint4*>(y10 + ((b * h + j) * (int64_t)(w + 1) + i) * C + c0), f2);
    unpack8(*reinterpret_cast<const uint4*>(y11 + ((b * h + j) * (int64_t)w + i) * C + c0), f3);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = (f0[k] + f1[k]) + (f2[k] + f3[k]);
    *reinterpret_cast<uint4*>(out + ((b * h + j) * (int64_t)w + i) * C + c0) = pack8(o);
  }
}

// Gradient of the ToRGB skip path: skip_up = upfirdn2d(skip, outer(k1, k1), up 2, pad (2,1)) (model.py:31-49, 358) is the
// 2x2-tap polyphase filter  up[2m] = f0 x[m-1] + f2 x[m],  up[2m+1] = f1 x[m] + f3 x[m+1]  per axis (f = flipped k1), so
//   g_skip[m] = f3 g[2m-1] + f2 g[2m] + f1 g[2m+1] + f0 g[2m+2]      (a 4-tap stride-2 filter per axis, zero outside)
// planes = B * 3 image planes of fp32 [2h, 2w] -> [h, w]; thread = 4 adjacent outputs of a row.
__global__ void __launch_bounds__(256)
skip_grad_kernel(const float* __restrict__ g, float* __restrict__ gx, int h, int w, int64_t total, float f0, float f1,
                 float f2, float f3) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int wq = (w + 3) >> 2;
  const int xq = (int)(t % wq);
  const int y = (int)((t / wq) % h);
  const int64_t plane = t / ((int64_t)wq * h);
  const int W2 = 2 * w, H2 = 2 * h;
  const float* gp = g + plane * (int64_t)H2 * W2;
  const float fy[4] = {f3, f2, f1, f0};
  const int x0 = xq * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int dy = 0; dy < 4; ++dy) {
    const int Y = 2 * y - 1 + dy;
    if (Y < 0 || Y >= H2) continue;
    const float* row = gp + (int64_t)Y * W2;
    float v[10];                                   // columns 2*x0-1 .. 2*x0+8
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const int X = 2 * x0 - 1 + i;
      v[i] = (X >= 0 && X < W2) ? __ldg(row + X) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      acc[k] = fmaf(fy[dy], fmaf(f3, v[2 * k], fmaf(f2, v[2 * k + 1], fmaf(f1, v[2 * k + 2], f0 * v[2 * k + 3]))), acc[k]);
  }
  float* out = gx + (plane * h + y) * (int64_t)w + x0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (x0 + k < w) out[k] = acc[k];
}

// pixel chunks: a function of the image size ONLY (never of the batch), so that a sample's partial sums -- and with
// them its gradients -- are bit-identical whatever batch it is part of: at most 128 chunks of at least 256 pixels
int64_t plan_chunks(int B, int64_t HW, int64_t* chunk_out) {
  (void)B;
  int64_t want = ceil_div64(HW, 256);
  if (want > 128) want = 128;
  if (want < 1) want = 1;
  const int64_t chunk = ceil_div64(HW, want);
  *chunk_out = chunk;
  return ceil_div64(HW, chunk);
}

bool good_channels(int C) { return C >= 8 && C % 8 == 0 && C / 8 <= kGaThreads && kGaThreads % (C / 8) == 0; }

}  // namespace
}  // namespace w2e

using namespace w2e;

extern "C" int64_t w2e_grad_assemble_workspace(int B, int64_t HW, int C) {
  if (B <= 0 || HW <= 0 || C <= 0) return 0;
  int64_t chunk;
  const int64_t n = plan_chunks(B, HW, &chunk);
  return (int64_t)B * n * 3 * C;   // floats
}

extern "C" int w2e_grad_assemble_nhwc(const void* gxs, const float* s_next, const void* act, const float* g_rgb,
                                      const float* w_rgb, const float* s_rgb, const float* noise, const float* noise_w,
                                      int noise_batch, const float* bias, const float* demod, int act_kind, void* gz,
                                      float* sums, float* workspace, int B, int64_t HW, int C, void* stream) {
  W2E_CHECK_ARG(act && sums && workspace && (gxs || g_rgb), "grad_assemble_nhwc: null pointer");
  W2E_CHECK_ARG(B >= 0 && HW > 0 && good_channels(C), "grad_assemble_nhwc: C must be a multiple of 8 dividing 2048 (got %d)", C);
  W2E_CHECK_ARG(g_rgb == nullptr || (w_rgb && s_rgb), "grad_assemble_nhwc: ToRGB gradient needs w_rgb and s_rgb");
  W2E_CHECK_ARG(noise == nullptr || (noise_w && (noise_batch == 1 || noise_batch == B)), "grad_assemble_nhwc: noise");
  W2E_CHECK_ARG(B <= 65535, "grad_assemble_nhwc: batch above 65535");
  W2E_CHECK_ARG((((uintptr_t)gxs | (uintptr_t)act | (uintptr_t)gz) & 15) == 0, "grad_assemble_nhwc: tensors must be 16-byte aligned");
  if (B == 0) return W2E_OK;
  GaParams P;
  P.gxs = (const __nv_bfloat16*)gxs; P.s_next = s_next; P.act = (const __nv_bfloat16*)act; P.g_rgb = g_rgb; P.w_rgb = w_rgb;
  P.s_rgb = s_rgb; P.noise = noise; P.noise_w = noise_w; P.bias = bias; P.demod = demod; P.gz = (__nv_bfloat16*)gz;
  P.partial = workspace; P.HW = HW; P.C = C; P.noise_per_sample = (noise && noise_batch != 1) ? 1 : 0;
  P.lrelu = act_kind == W2E_ACT_LRELU ? 1 : 0;
  P.want_r3 = demod != nullptr ? 1 : 0;
  P.nchunks = (int)plan_chunks(B, HW, &P.chunk);
  const int lanes = kGaThreads / (C / 8);
  const size_t smem = ((size_t)9 * C + (size_t)3 * lanes * C) * sizeof(float);   // constants + 24 KB of lane partials
  W2E_CHECK_ARG(smem <= 48 * 1024, "grad_assemble_nhwc: shared memory");
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((unsigned)P.nchunks, (unsigned)B);
  if (g_rgb && demod) grad_assemble_nhwc_kernel<true, true><<<grid, kGaThreads, smem, st>>>(P);
  else if (g_rgb) grad_assemble_nhwc_kernel<true, false><<<grid, kGaThreads, smem, st>>>(P);
  else if (demod) grad_assemble_nhwc_kernel<false, true><<<grid, kGaThreads, smem, st>>>(P);
  else grad_assemble_nhwc_kernel<false, false><<<grid, kGaThreads, smem, st>>>(P);
  W2E_LAUNCH_OK();
  const int64_t total = (int64_t)B * 3 * C;
  reduce_chunks_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(workspace, sums, P.nchunks, 3 * C, total);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_rowdot_nhwc(const void* a, const void* b, int b_batch, float* dot, float* workspace, int B, int64_t HW,
                               int C, void* stream) {
  W2E_CHECK_ARG(a && b && dot && workspace, "rowdot_nhwc: null pointer");
  W2E_CHECK_ARG(B >= 0 && HW > 0 && good_channels(C) && (b_batch == 1 || b_batch == B) && B <= 65535, "rowdot_nhwc: bad shape");
  W2E_CHECK_ARG((((uintptr_t)a | (uintptr_t)b) & 15) == 0, "rowdot_nhwc: tensors must be 16-byte aligned");
  if (B == 0) return W2E_OK;
  int64_t chunk;
  const int nchunks = (int)plan_chunks(B, HW, &chunk);
  const int lanes = kGaThreads / (C / 8);
  const size_t smem = (size_t)lanes * C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  rowdot_nhwc_kernel<<<dim3((unsigned)nchunks, (unsigned)B), kGaThreads, smem, st>>>(
      (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, workspace, HW, C, nchunks, chunk, b_batch);
  W2E_LAUNCH_OK();
  const int64_t total = (int64_t)B * C;
  reduce_chunks_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(workspace, dot, nchunks, C, total);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_sum4_nhwc(const void* y00, const void* y01, const void* y10, const void* y11, void* out, int B, int h,
                             int w, int C, void* stream) {
  W2E_CHECK_ARG(y00 && y01 && y10 && y11 && out, "sum4_nhwc: null pointer");
  W2E_CHECK_ARG(B >= 0 && h > 0 && w > 0 && C > 0 && C % 8 == 0, "sum4_nhwc: bad shape");
  if (B == 0) return W2E_OK;
  const int64_t total = (int64_t)B * h * w * (C / 8);
  int64_t blocks = ceil_div64(total, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  sum4_nhwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)y00, (const __nv_bfloat16*)y01, (const __nv_bfloat16*)y10, (const __nv_bfloat16*)y11,
      (__nv_bfloat16*)out, h, w, C, total);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_skip_grad(const float* g, float* g_skip, const float* host_taps1d, int64_t planes, int h, int w,
                             void* stream) {
  W2E_CHECK_ARG(g && g_skip && host_taps1d, "skip_grad: null pointer");
  W2E_CHECK_ARG(planes >= 0 && h > 0 && w > 0, "skip_grad: bad shape");
  if (planes == 0) return W2E_OK;
  const int64_t total = planes * h * ((w + 3) / 4);
  // flipped 1-D taps f[i] = k1[3 - i] (the forward's polyphase coefficients, csrc/torgb_blend.cu)
  skip_grad_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
      g, g_skip, h, w, total, host_taps1d[3], host_taps1d[2], host_taps1d[1], host_taps1d[0]);
  W2E_LAUNCH_OK();
  return W2E_OK;
}
