// Grouped launch of the region-attention heads of the cluster-style mapper:
// attention/run_attention.py:803-806 and 829-839 -- per captured generator feature map F_h [B,C_h,H_h,H_h] one
// 1x1 StyledConv(C_h -> 32, stylespace input) followed by F.interpolate(nearest) to size x size, the 17..18 results
// concatenated along the channel axis (:841).  The reference evaluates every head on the full-resolution map (up to
// 1024^2 pixels) and keeps size^2 of them; a 1x1 modulated convolution, NoiseInjection, FusedLeakyReLU and a nearest
// resize all act per pixel, so here ONE launch computes, for all heads, only the pixels that survive the resize and
// writes them straight into the concatenated map:
//   out[b, 32 h + o, Y, X] = lrelu( d[b,o] * sum_c W_h[o,c] s_h[b,c] F_h[b,c,y,x] + nw_h * noise_h[b,.] + bias_h[o] ) * sqrt(2)
//   (y, x) = nearest source pixel of (Y, X);  d[b,o] = rsqrt( sum_c (W_h[o,c] s_h[b,c])^2 + 1e-8 )   (model.py:239-247)
// Noise (model.py:286-288 draws it at the head's own resolution) is indexed by the SOURCE pixel when the map is
// smaller than `size` (replicated pixels share their noise, as in the reference) and by the output pixel otherwise.
// Backward (the reference trains these heads, run_attention.py:725-735): two launches for all heads --
//   attn_heads_bwd_reduce  g_pre = g * lrelu', per (head, sample): d bias, d noise weight, d * dL/dd, and g_acc = g_pre * d
//   attn_heads_bwd_wgrad   M[b,o,c] = sum_p g_acc[b,o,p] F[b,c,p] in registers -> dW[o,c], ds[b,c] incl. the demodulation terms
// Every reduction runs in a fixed order (no atomics): results do not depend on the launch geometry.
#include "common.cuh"

namespace w2e {

constexpr int kHeadOut = 32;      // output channels of every head (run_attention.py:728)
constexpr int kMaxHeads = 24;
constexpr int kHeadPix = 128;     // output pixels per block (forward) / per shared-memory tile (wgrad)
constexpr int kHeadCk = 64;       // channels per shared-memory weight chunk (forward)
constexpr int kWgC = 32;          // channels per block (wgrad): 8 warps x 4 channels

struct HeadsParams {
  const float* feat[kMaxHeads];     // [B,C,H,H] fp32
  const float* weight[kMaxHeads];   // [32,C], equalised-lr scale folded in
  const float* style[kMaxHeads];    // [B,C]
  const float* bias[kMaxHeads];     // [32]
  const float* noise_w[kMaxHeads];  // device scalar
  const float* noise[kMaxHeads];    // [B,r,r], r = min(H, size), or null
  float* dW[kMaxHeads];             // backward: [32,C] gradient of `weight`
  float* ds[kMaxHeads];             // backward: [B,C] gradient of `style`
  int chans[kMaxHeads], res[kMaxHeads];
  int cum_blocks[kMaxHeads + 1];    // wgrad: first block of every head
  int nheads, B, size;
};

// F.interpolate(mode="nearest"): src = min(floor(dst * (float)in / out), in - 1)
__device__ __forceinline__ int nearest_src(int dst, float scale, int in) { return min((int)floorf((float)dst * scale), in - 1); }

__global__ void __launch_bounds__(kHeadPix)
attn_heads_fwd_kernel(const __grid_constant__ HeadsParams P, float* __restrict__ out, float* __restrict__ demod) {
  const int h = blockIdx.z, b = blockIdx.y, tid = threadIdx.x;
  const int C = P.chans[h], H = P.res[h], S = P.size, SS = S * S;
  const int p = blockIdx.x * kHeadPix + tid;
  const bool valid = p < SS;
  const int Y = valid ? p / S : 0, X = valid ? p - (p / S) * S : 0;
  const float scale = (float)H / (float)S;
  const int sy = nearest_src(Y, scale, H), sx = nearest_src(X, scale, H);
  const int64_t plane = (int64_t)H * H;
  const float* f = P.feat[h] + (int64_t)b * C * plane + (int64_t)sy * H + sx;
  const float* W = P.weight[h];
  const float* s = P.style[h] + (int64_t)b * C;

  __shared__ __align__(16) float w_s[kHeadCk][kHeadOut];   // (W * s)[c][o] of the current channel chunk
  __shared__ float d_part[4][kHeadOut];
  __shared__ float d_s[kHeadOut];
  float acc[kHeadOut];
#pragma unroll
  for (int o = 0; o < kHeadOut; ++o) acc[o] = 0.f;
  const int wo = tid & 31, wpart = tid >> 5;   // weight staging: output channel, one of four channel phases
  float dsum = 0.f;
  for (int c0 = 0; c0 < C; c0 += kHeadCk) {
    const int n = min(kHeadCk, C - c0);
    for (int cc = wpart; cc < kHeadCk; cc += 4) {
      float ws = 0.f;
      if (cc < n) ws = __ldg(W + (int64_t)wo * C + c0 + cc) * __ldg(s + c0 + cc);
      w_s[cc][wo] = ws;
      dsum = fmaf(ws, ws, dsum);
    }
    __syncthreads();
    if (valid) {
#pragma unroll 4
      for (int cc = 0; cc < n; ++cc) {
        const float v = __ldg(f + (int64_t)(c0 + cc) * plane);
        const float4* wr = reinterpret_cast<const float4*>(w_s[cc]);
#pragma unroll
        for (int o4 = 0; o4 < kHeadOut / 4; ++o4) {
          const float4 w4 = wr[o4];
          acc[4 * o4 + 0] = fmaf(w4.x, v, acc[4 * o4 + 0]);
          acc[4 * o4 + 1] = fmaf(w4.y, v, acc[4 * o4 + 1]);
          acc[4 * o4 + 2] = fmaf(w4.z, v, acc[4 * o4 + 2]);
          acc[4 * o4 + 3] = fmaf(w4.w, v, acc[4 * o4 + 3]);
        }
      }
    }
    __syncthreads();
  }
  d_part[wpart][wo] = dsum;
  __syncthreads();
  if (tid < kHeadOut) {
    const float d = rsqrtf(((d_part[0][tid] + d_part[1][tid]) + (d_part[2][tid] + d_part[3][tid])) + 1e-8f);
    d_s[tid] = d;
    if (blockIdx.x == 0) demod[((int64_t)h * P.B + b) * kHeadOut + tid] = d;
  }
  __syncthreads();
  if (!valid) return;
  float nz = 0.f;
  if (P.noise[h]) {
    const int r = min(H, S);
    const int ny = H < S ? sy : Y, nx = H < S ? sx : X;
    nz = __ldg(P.noise_w[h]) * __ldg(P.noise[h] + ((int64_t)b * r + ny) * r + nx);
  }
  const float* bias = P.bias[h];
  float* o_ptr = out + ((int64_t)b * P.nheads * kHeadOut + (int64_t)h * kHeadOut) * SS + p;
#pragma unroll
  for (int o = 0; o < kHeadOut; ++o) {
    const float pre = fmaf(acc[o], d_s[o], nz) + __ldg(bias + o);
    o_ptr[(int64_t)o * SS] = (pre > 0.f ? pre : 0.2f * pre) * 1.41421356237309515f;
  }
}

// One block per (sample, head): g_acc = g * lrelu'(pre) * sqrt2 * d  and the per-(sample, head) reductions
//   red[h][b][0][o] = sum_p g_pre                      (d bias)
//   red[h][b][1][o] = sum_p g_pre * (pre - nz - bias)  (= d * dL/dd: pre - nz - bias is d * acc)
//   red[h][b][2][0] = sum_{p,o} g_pre * noise          (d noise weight)
__global__ void __launch_bounds__(256)
attn_heads_bwd_reduce_kernel(const __grid_constant__ HeadsParams P, const float* __restrict__ gout,
                             const float* __restrict__ out, const float* __restrict__ demod, float* __restrict__ gacc,
                             float* __restrict__ red) {
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const int H = P.res[h], S = P.size, SS = S * S;
  const float scale = (float)H / (float)S;
  const int64_t base = ((int64_t)b * P.nheads * kHeadOut + (int64_t)h * kHeadOut) * SS;
  const float nw = P.noise[h] ? __ldg(P.noise_w[h]) : 0.f;
  const float* bias = P.bias[h];
  const float* d = demod + ((int64_t)h * P.B + b) * kHeadOut;
  constexpr float kS2 = 1.41421356237309515f, kInvPos = 0.70710678118654752f, kInvNeg = 3.53553390593273762f;
  __shared__ float s_red[8][2 * kHeadOut + 1];
  const int lane = tid & 31, warp = tid >> 5;
  // thread = pixel (coalesced over pixels), all 32 channels in registers; warp-shuffle then cross-warp sums in a fixed order
  float sb[kHeadOut], sd[kHeadOut], sn = 0.f;
#pragma unroll
  for (int o = 0; o < kHeadOut; ++o) { sb[o] = 0.f; sd[o] = 0.f; }
  for (int p = tid; p < SS; p += 256) {
    float nz = 0.f;
    if (P.noise[h]) {
      const int Y = p / S, X = p - Y * S, r = min(H, S);
      const int ny = H < S ? nearest_src(Y, scale, H) : Y, nx = H < S ? nearest_src(X, scale, H) : X;
      nz = __ldg(P.noise[h] + ((int64_t)b * r + ny) * r + nx);
    }
#pragma unroll
    for (int o = 0; o < kHeadOut; ++o) {
      const float g = __ldg(gout + base + (int64_t)o * SS + p), y = __ldg(out + base + (int64_t)o * SS + p);
      const float gp = g * (y > 0.f ? kS2 : 0.2f * kS2);
      const float pre = y > 0.f ? y * kInvPos : y * kInvNeg;
      sb[o] += gp;
      sd[o] = fmaf(gp, pre - nw * nz - __ldg(bias + o), sd[o]);
      sn = fmaf(gp, nz, sn);
      gacc[base + (int64_t)o * SS + p] = gp * __ldg(d + o);
    }
  }
#pragma unroll
  for (int o = 0; o < kHeadOut; ++o) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      sb[o] += __shfl_xor_sync(0xffffffffu, sb[o], off);
      sd[o] += __shfl_xor_sync(0xffffffffu, sd[o], off);
    }
    if (lane == 0) { s_red[warp][o] = sb[o]; s_red[warp][kHeadOut + o] = sd[o]; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sn += __shfl_xor_sync(0xffffffffu, sn, off);
  if (lane == 0) s_red[warp][2 * kHeadOut] = sn;
  __syncthreads();
  float* r = red + ((int64_t)h * P.B + b) * (3 * kHeadOut);
  if (tid < 2 * kHeadOut + 1) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w][tid];
    r[tid] = t;
  }
}

// Block = 32 channels of one head (8 warps x 4 channels, lane = output channel o), samples walked one after the other:
//   M[b,o,c] = sum_p g_acc[b,o,p] * F[b,c,src(p)]            (register accumulators, tiles of 128 pixels in shared memory)
//   dW[o,c]  = sum_b ( s[b,c] M - t[b,o] W[o,c] s[b,c]^2 )     t[b,o] = (d * dL/dd)[b,o] * d[b,o]^2
//   ds[b,c]  = sum_o ( W[o,c] M - t[b,o] s[b,c] W[o,c]^2 )
__global__ void __launch_bounds__(256)
attn_heads_bwd_wgrad_kernel(const __grid_constant__ HeadsParams P, const float* __restrict__ gacc,
                            const float* __restrict__ demod, const float* __restrict__ red) {
  int h = 0;
  while (h + 1 < P.nheads && (int)blockIdx.x >= P.cum_blocks[h + 1]) ++h;
  const int c0 = ((int)blockIdx.x - P.cum_blocks[h]) * kWgC;
  const int C = P.chans[h], H = P.res[h], S = P.size, SS = S * S, tid = threadIdx.x;
  const int o = tid & 31, warp = tid >> 5;
  const float scale = (float)H / (float)S;
  const int64_t plane = (int64_t)H * H;
  __shared__ float g_s[kHeadOut][kHeadPix + 1];
  __shared__ float f_s[kWgC][kHeadPix];
  float wv[4], dw[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + warp * 4 + k;
    wv[k] = c < C ? __ldg(P.weight[h] + (int64_t)o * C + c) : 0.f;
    dw[k] = 0.f;
  }
  for (int b = 0; b < P.B; ++b) {
    const int64_t gbase = ((int64_t)b * P.nheads * kHeadOut + (int64_t)h * kHeadOut) * SS;
    const float* fb = P.feat[h] + (int64_t)b * C * plane;
    float m[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p0 = 0; p0 < SS; p0 += kHeadPix) {
      for (int i = tid; i < kHeadOut * kHeadPix; i += 256) {
        const int oo = i / kHeadPix, pp = i - oo * kHeadPix;
        g_s[oo][pp] = (p0 + pp < SS) ? __ldg(gacc + gbase + (int64_t)oo * SS + p0 + pp) : 0.f;
      }
      for (int i = tid; i < kWgC * kHeadPix; i += 256) {
        const int cc = i / kHeadPix, pp = i - cc * kHeadPix;
        float v = 0.f;
        if (c0 + cc < C && p0 + pp < SS) {
          const int Y = (p0 + pp) / S, X = (p0 + pp) - Y * S;
          v = __ldg(fb + (int64_t)(c0 + cc) * plane + (int64_t)nearest_src(Y, scale, H) * H + nearest_src(X, scale, H));
        }
        f_s[cc][pp] = v;
      }
      __syncthreads();
#pragma unroll 8
      for (int pp = 0; pp < kHeadPix; ++pp) {
        const float g = g_s[o][pp];
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] = fmaf(g, f_s[warp * 4 + k][pp], m[k]);
      }
      __syncthreads();
    }
    const float dd = __ldg(demod + ((int64_t)h * P.B + b) * kHeadOut + o);
    const float t = __ldg(red + ((int64_t)h * P.B + b) * (3 * kHeadOut) + kHeadOut + o) * dd * dd;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + warp * 4 + k;
      const float sv = c < C ? __ldg(P.style[h] + (int64_t)b * C + c) : 0.f;
      dw[k] += sv * m[k] - t * wv[k] * sv * sv;
      float part = wv[k] * m[k] - t * sv * wv[k] * wv[k];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
      if (o == 0 && c < C) P.ds[h][(int64_t)b * C + c] = part;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + warp * 4 + k;
    if (c < C) P.dW[h][(int64_t)o * C + c] = dw[k];
  }
}

static int fill_params(HeadsParams& P, int nheads, const float* const* feat, const int* chans, const int* res,
                       const float* const* weight, const float* const* style, const float* const* bias,
                       const float* const* noise_w, const float* const* noise, int B, int size) {
  W2E_CHECK_ARG(nheads > 0 && nheads <= kMaxHeads, "attn_heads: 1..%d heads (got %d)", kMaxHeads, nheads);
  W2E_CHECK_ARG(B > 0 && size > 0 && feat && chans && res && weight && style && bias, "attn_heads: bad arguments");
  memset(&P, 0, sizeof(P));
  P.nheads = nheads; P.B = B; P.size = size;
  int blocks = 0;
  for (int h = 0; h < nheads; ++h) {
    W2E_CHECK_ARG(feat[h] && weight[h] && style[h] && bias[h] && chans[h] > 0 && res[h] > 0, "attn_heads: head %d", h);
    W2E_CHECK_ARG(!(noise && noise[h]) || (noise_w && noise_w[h]), "attn_heads: head %d has noise but no noise weight", h);
    P.feat[h] = feat[h]; P.weight[h] = weight[h]; P.style[h] = style[h]; P.bias[h] = bias[h];
    P.noise_w[h] = noise_w ? noise_w[h] : nullptr;
    P.noise[h] = noise ? noise[h] : nullptr;
    P.chans[h] = chans[h]; P.res[h] = res[h];
    P.cum_blocks[h] = blocks;
    blocks += ceil_div(chans[h], kWgC);
  }
  P.cum_blocks[nheads] = blocks;
  return W2E_OK;
}

}  // namespace w2e

using namespace w2e;

extern "C" int w2e_attn_heads_fwd(int nheads, const float* const* feat, const int* chans, const int* res,
                                  const float* const* weight, const float* const* style, const float* const* bias,
                                  const float* const* noise_w, const float* const* noise, float* out, float* demod, int B,
                                  int size, void* stream) {
  HeadsParams P;
  int rc = fill_params(P, nheads, feat, chans, res, weight, style, bias, noise_w, noise, B, size);
  if (rc) return rc;
  W2E_CHECK_ARG(out && demod, "attn_heads_fwd: null output");
  const dim3 grid((unsigned)ceil_div(size * size, kHeadPix), (unsigned)B, (unsigned)nheads);
  attn_heads_fwd_kernel<<<grid, kHeadPix, 0, (cudaStream_t)stream>>>(P, out, demod);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_attn_heads_bwd(int nheads, const float* const* feat, const int* chans, const int* res,
                                  const float* const* weight, const float* const* style, const float* const* bias,
                                  const float* const* noise_w, const float* const* noise, const float* gout,
                                  const float* out, const float* demod, float* gacc, float* red, float* const* dW,
                                  float* const* ds, int B, int size, void* stream) {
  HeadsParams P;
  int rc = fill_params(P, nheads, feat, chans, res, weight, style, bias, noise_w, noise, B, size);
  if (rc) return rc;
  W2E_CHECK_ARG(gout && out && demod && gacc && red && dW && ds, "attn_heads_bwd: null pointer");
  for (int h = 0; h < nheads; ++h) {
    W2E_CHECK_ARG(dW[h] && ds[h], "attn_heads_bwd: head %d has no gradient buffers", h);
    P.dW[h] = dW[h]; P.ds[h] = ds[h];
  }
  attn_heads_bwd_reduce_kernel<<<dim3((unsigned)B, (unsigned)nheads), 256, 0, (cudaStream_t)stream>>>(P, gout, out, demod,
                                                                                                      gacc, red);
  W2E_LAUNCH_OK();
  attn_heads_bwd_wgrad_kernel<<<(unsigned)P.cum_blocks[nheads], 256, 0, (cudaStream_t)stream>>>(P, gacc, demod, red);
  W2E_LAUNCH_OK();
  return W2E_OK;
}
