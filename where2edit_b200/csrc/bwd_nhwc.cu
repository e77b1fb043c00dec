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

// grid (nchunks, B); thread = (pixel lane, 8-channel group); shared: 3 * lanes * C floats for the lane reduction
__global__ void __launch_bounds__(kGaThreads)
grad_assemble_nhwc_kernel(const GaParams P) {
  extern __shared__ float red[];
  const int groups = P.C >> 3;
  const int lanes = kGaThreads / groups;          // pixel lanes per CTA (C <= 2048, C % 8 == 0, groups | 256)
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
  const int b = blockIdx.y;
  const int c0 = g * 8;
  const int64_t p_begin = (int64_t)blockIdx.x * P.chunk;
  const int64_t p_end = min(P.HW, p_begin + P.chunk);
  float sn[8], ws[3][8], wr[3][8], dm[8], bs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c0 + k;
    sn[k] = P.s_next ? __ldg(P.s_next + (int64_t)b * P.C + c) : 1.f;
    dm[k] = P.demod ? __ldg(P.demod + (int64_t)b * P.C + c) : 1.f;
    bs[k] = P.bias ? __ldg(P.bias + c) : 0.f;
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      wr[o][k] = P.g_rgb ? __ldg(P.w_rgb + o * P.C + c) : 0.f;
      ws[o][k] = P.g_rgb ? wr[o][k] * __ldg(P.s_rgb + (int64_t)b * P.C + c) : 0.f;
    }
  }
  const float nw = P.noise ? __ldg(P.noise_w) : 0.f;
  const float gain = P.lrelu ? 1.41421356237309515f : 1.f;
  const float neg = P.lrelu ? 0.2f : 1.f;
  float r1[8], r2[8], r3[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r1[k] = r2[k] = r3[k] = 0.f;
  const int64_t base = (int64_t)b * P.HW;
  for (int64_t p = p_begin + lane; p < p_end; p += lanes) {
    const int64_t e = (base + p) * P.C + c0;
    float a[8], gx[8];
    unpack8(*reinterpret_cast<const uint4*>(P.act + e), a);
    if (P.gxs) unpack8(*reinterpret_cast<const uint4*>(P.gxs + e), gx);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) gx[k] = 0.f;
    }
    float gr[3] = {0.f, 0.f, 0.f};
    if (P.g_rgb) {
#pragma unroll
      for (int o = 0; o < 3; ++o) gr[o] = __ldg(P.g_rgb + ((int64_t)b * 3 + o) * P.HW + p);
    }
    const float nz = P.noise ? nw * __ldg(P.noise + (P.noise_per_sample ? (int64_t)b * P.HW : 0) + p) : 0.f;
    float out[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float t = fmaf(gr[0], wr[0][k], fmaf(gr[1], wr[1][k], gr[2] * wr[2][k]));      // unmodulated ToRGB pull-back
      const float ga = fmaf(gx[k], sn[k], fmaf(gr[0], ws[0][k], fmaf(gr[1], ws[1][k], gr[2] * ws[2][k])));
      r1[k] = fmaf(gx[k], a[k], r1[k]);
      r2[k] = fmaf(t, a[k], r2[k]);
      const bool pos = a[k] > 0.f;
      const float gp = ga * (pos ? gain : gain * neg);                 // gradient w.r.t. the pre-activation y
      if (P.want_r3) {
        const float y = a[k] * (pos ? 1.f / gain : 1.f / (gain * neg));
        r3[k] = fmaf(gp, y - nz - bs[k], r3[k]);
      }
      out[k] = gp * dm[k];
    }
    if (P.gz) *reinterpret_cast<uint4*>(P.gz + e) = pack8(out);
  }
  // deterministic reduction over the pixel lanes (lane order), per channel
  float* s1 = red;
  float* s2 = red + (size_t)lanes * P.C;
  float* s3 = red + (size_t)2 * lanes * P.C;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s1[lane * P.C + c0 + k] = r1[k];
    s2[lane * P.C + c0 + k] = r2[k];
    s3[lane * P.C + c0 + k] = r3[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * P.C; i += kGaThreads) {
    const int r = i / P.C, c = i - r * P.C;
    const float* src = red + (size_t)r * lanes * P.C + c;
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += src[(size_t)l * P.C];
    P.partial[(((int64_t)b * P.nchunks + blockIdx.x) * 3 + r) * P.C + c] = acc;
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
  const size_t smem = (size_t)3 * lanes * C * sizeof(float);
  // (lanes * C == 2048 floats whatever C: 24 KB)
  cudaStream_t st = (cudaStream_t)stream;
  grad_assemble_nhwc_kernel<<<dim3((unsigned)P.nchunks, (unsigned)B), kGaThreads, smem, st>>>(P);
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
