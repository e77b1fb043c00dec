// Exact fp32 convolution engine (CUDA cores, FFMA) for the shared-weight modulated convolution.
// Replaces the per-sample-weight grouped F.conv2d / F.conv_transpose2d of
// models/stylegan2/model.py:239-274 in "fp32 mode" (the <=1e-4 parity mode) and serves as the
// dgrad engine of the backward.  The tensor-core (tcgen05) path for the
// bf16 mode lives in modconv_tc.cu; this file is the exactness-first variant.
//
// One CTA computes a tile of 8x16 grid points x 64 output channels for one sample, looping over
// input channels in chunks of 8.  Per chunk the input patch (with halo, multiplied by the
// per-sample input scale = style, i.e. "modulation folded into the activation load") and the
// weight slice [taps][8][64] are staged in shared memory; each thread accumulates a 4 (pixels
// along x) x 8 (channels) register tile.  Epilogue: per-sample output scale (= demodulation),
// optional noise + bias + leaky-ReLU*sqrt(2).
#include "common.cuh"

namespace w2e {

constexpr int kTH = 8, kTW = 16, kBCO = 64, kBCI = 8, kMaxTaps = 9;
constexpr int kMaxIH = (kTH - 1) * 2 + 1 + 2, kMaxIW = (kTW - 1) * 2 + 1 + 2;  // stride-2 gather with a 3-tap span

struct ConvTaps {
  int n;
  int dy[kMaxTaps], dx[kMaxTaps], slot[kMaxTaps];
};

struct ConvParams {
  const float* x; const float* w; const float* in_scale; const float* out_scale;
  const float* bias; const float* noise; const float* noise_w; float* y;
  int noise_per_sample;
  int B, Cin, Cout, in_h, in_w, out_h, out_w, grid_h, grid_w, in_stride, out_stride, py, px;
  int dy_min, dx_min, ih, iw;  // staged input patch geometry
  int act, accumulate;
  ConvTaps taps;
};

__global__ void __launch_bounds__(256)
conv_engine_kernel(const __grid_constant__ ConvParams P) {
  __shared__ __align__(16) float s_in[kBCI * kMaxIH * kMaxIW];
  __shared__ __align__(16) float s_w[kMaxTaps * kBCI * kBCO];

  const int tid = threadIdx.x;
  const int tc = tid & 7;    // channel lane: channels co0 + tc*4 + {0..3} and co0 + 32 + tc*4 + {0..3}
  const int tp = tid >> 3;   // pixel group 0..31: row tp/4, columns (tp%4)*4 .. +3
  const int tiles_x = (P.grid_w + kTW - 1) / kTW;
  const int j0 = (blockIdx.x / tiles_x) * kTH, i0 = (blockIdx.x % tiles_x) * kTW;
  const int co0 = blockIdx.y * kBCO;
  const int b = blockIdx.z;
  const int iy0 = j0 * P.in_stride + P.dy_min, ix0 = i0 * P.in_stride + P.dx_min;
  const int patch = P.ih * P.iw;
  const int lr = tp >> 2, lc = (tp & 3) * 4;

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const float* xb = P.x + (int64_t)b * P.Cin * P.in_h * P.in_w;
  for (int ci0 = 0; ci0 < P.Cin; ci0 += kBCI) {
    for (int e = tid; e < kBCI * patch; e += 256) {
      const int ci = e / patch, r = (e % patch) / P.iw, c = e % P.iw;
      const int gy = iy0 + r, gx = ix0 + c, cg = ci0 + ci;
      float v = 0.f;
      if (cg < P.Cin && gy >= 0 && gy < P.in_h && gx >= 0 && gx < P.in_w) {
        v = __ldg(xb + ((int64_t)cg * P.in_h + gy) * P.in_w + gx);
        if (P.in_scale) v *= __ldg(P.in_scale + (int64_t)b * P.Cin + cg);
      }
      s_in[e] = v;
    }
    for (int e = tid; e < P.taps.n * kBCI * kBCO; e += 256) {
      const int t = e / (kBCI * kBCO), ci = (e / kBCO) % kBCI, co = e % kBCO;
      const int cg = ci0 + ci, og = co0 + co;
      float v = 0.f;
      if (cg < P.Cin && og < P.Cout) v = __ldg(P.w + ((int64_t)P.taps.slot[t] * P.Cin + cg) * P.Cout + og);
      s_w[e] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int t = 0; t < P.taps.n; ++t) {
      const int roff = (lr * P.in_stride + P.taps.dy[t] - P.dy_min) * P.iw + (P.taps.dx[t] - P.dx_min);
#pragma unroll
      for (int ci = 0; ci < kBCI; ++ci) {
        const float4 w0 = *reinterpret_cast<const float4*>(&s_w[(t * kBCI + ci) * kBCO + tc * 4]);
        const float4 w1 = *reinterpret_cast<const float4*>(&s_w[(t * kBCI + ci) * kBCO + 32 + tc * 4]);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const float* src = &s_in[ci * patch + roff];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float xv = src[(lc + i) * P.in_stride];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xv, wv[j], acc[i][j]);
        }
      }
    }
    __syncthreads();
  }

  const int j = j0 + lr;
  if (j >= P.grid_h) return;
  const int oy = j * P.out_stride + P.py;
  const float nw = (P.noise != nullptr) ? __ldg(P.noise_w) : 0.f;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int co = co0 + (jj < 4 ? tc * 4 + jj : 32 + tc * 4 + (jj - 4));
    if (co >= P.Cout) continue;
    const float os = P.out_scale ? __ldg(P.out_scale + (int64_t)b * P.Cout + co) : 1.f;
    const float bs = P.bias ? __ldg(P.bias + co) : 0.f;
    float* yrow = P.y + (((int64_t)b * P.Cout + co) * P.out_h + oy) * P.out_w;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gi = i0 + lc + i;
      if (gi >= P.grid_w) continue;
      const int ox = gi * P.out_stride + P.px;
      float v = acc[i][jj] * os;
      if (P.act == W2E_ACT_LRELU) {
        if (P.noise)
          v = fmaf(nw, __ldg(P.noise + (P.noise_per_sample ? (int64_t)b * P.out_h * P.out_w : 0) +
                             (int64_t)oy * P.out_w + ox), v);
        v = lrelu_gain(v + bs, 0.2f, 1.41421356237309515f);
      } else {
        v += bs;
      }
      if (P.accumulate) v += yrow[ox];
      yrow[ox] = v;
    }
  }
}

}  // namespace w2e

using namespace w2e;

extern "C" int w2e_conv_engine_f32(const float* x, const float* w, const float* in_scale, const float* out_scale,
                                   const float* bias, const float* noise, const float* noise_w, int noise_batch,
                                   float* y, int B, int Cin, int Cout, int in_h, int in_w, int out_h, int out_w,
                                   int grid_h, int grid_w, int in_stride, int out_stride, int py, int px,
                                   const int* host_taps, int ntaps, int act, int accumulate, void* stream) {
  W2E_CHECK_ARG(x && w && y && host_taps, "conv_engine: null pointer");
  W2E_CHECK_ARG(B >= 0 && Cin > 0 && Cout > 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, "conv_engine: bad shape");
  W2E_CHECK_ARG(ntaps > 0 && ntaps <= kMaxTaps, "conv_engine: %d taps (max %d)", ntaps, kMaxTaps);
  W2E_CHECK_ARG(in_stride == 1 || in_stride == 2, "conv_engine: in_stride must be 1 or 2");
  W2E_CHECK_ARG(out_stride >= 1 && py >= 0 && px >= 0, "conv_engine: bad output mapping");
  W2E_CHECK_ARG((grid_h - 1) * out_stride + py < out_h && (grid_w - 1) * out_stride + px < out_w,
                "conv_engine: output grid exceeds the output tensor");
  W2E_CHECK_ARG(noise == nullptr || noise_w != nullptr, "conv_engine: noise without weight");
  W2E_CHECK_ARG(noise == nullptr || noise_batch == 1 || noise_batch == B, "conv_engine: noise batch %d", noise_batch);
  W2E_CHECK_ARG(B <= 65535, "conv_engine: batch above 65535");
  if (B == 0 || grid_h <= 0 || grid_w <= 0) return W2E_OK;
  ConvParams P;
  P.x = x; P.w = w; P.in_scale = in_scale; P.out_scale = out_scale; P.bias = bias; P.noise = noise;
  P.noise_w = noise_w; P.y = y; P.noise_per_sample = (noise && noise_batch != 1) ? 1 : 0;
  P.B = B; P.Cin = Cin; P.Cout = Cout; P.in_h = in_h; P.in_w = in_w; P.out_h = out_h; P.out_w = out_w;
  P.grid_h = grid_h; P.grid_w = grid_w; P.in_stride = in_stride; P.out_stride = out_stride; P.py = py; P.px = px;
  P.act = act; P.accumulate = accumulate;
  int dy_min = 1 << 30, dy_max = -(1 << 30), dx_min = 1 << 30, dx_max = -(1 << 30);
  P.taps.n = ntaps;
  for (int t = 0; t < ntaps; ++t) {
    P.taps.dy[t] = host_taps[3 * t]; P.taps.dx[t] = host_taps[3 * t + 1]; P.taps.slot[t] = host_taps[3 * t + 2];
    W2E_CHECK_ARG(P.taps.slot[t] >= 0, "conv_engine: negative weight slot");
    dy_min = P.taps.dy[t] < dy_min ? P.taps.dy[t] : dy_min; dy_max = P.taps.dy[t] > dy_max ? P.taps.dy[t] : dy_max;
    dx_min = P.taps.dx[t] < dx_min ? P.taps.dx[t] : dx_min; dx_max = P.taps.dx[t] > dx_max ? P.taps.dx[t] : dx_max;
  }
  for (int t = ntaps; t < kMaxTaps; ++t) P.taps.dy[t] = P.taps.dx[t] = P.taps.slot[t] = 0;
  P.dy_min = dy_min; P.dx_min = dx_min;
  P.ih = (kTH - 1) * in_stride + 1 + (dy_max - dy_min);
  P.iw = (kTW - 1) * in_stride + 1 + (dx_max - dx_min);
  W2E_CHECK_ARG(P.ih <= kMaxIH && P.iw <= kMaxIW, "conv_engine: tap span too wide (%d x %d patch)", P.ih, P.iw);
  dim3 grid(ceil_div(grid_h, kTH) * ceil_div(grid_w, kTW), ceil_div(Cout, kBCO), B);
  conv_engine_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  W2E_LAUNCH_OK();
  return W2E_OK;
}
