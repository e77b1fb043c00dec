// upfirdn2d for NCHW planes -- replaces models/stylegan2/op/upfirdn2d.py:19-60.
//
// Two kernels:
//  * upfirdn2d_generic_kernel: any up/down/pad, any kernel up to 8x8, one output per thread.
//    Direct restatement of zero-stuff -> pad/crop -> true convolution -> decimate.
//  * blur4_sep_kernel: the hot shape (Blur after the up-convolution, model.py:206, and its
//    backward): up = down = 1, 4x4 separable kernel.  HBM-bound: every input element is read
//    from global memory once into a shared-memory tile (with its 3-pixel halo), the FIR runs
//    as a horizontal 4-tap pass in registers followed by a vertical 4-tap accumulation, and
//    each thread writes a 4x4 output patch with 128-bit stores.
//    Algorithmic bytes per call: planes * (in_h*in_w + out_h*out_w) * sizeof(T).
#include <cmath>

#include "common.cuh"

namespace w2e {

struct FirTaps {
  float k[64];  // flipped kernel, row-major [kh][kw]
};

template <typename T>
__global__ void __launch_bounds__(256)
upfirdn2d_generic_kernel(const T* __restrict__ x, T* __restrict__ y, const __grid_constant__ FirTaps taps,
                         int64_t planes, int in_h, int in_w, int out_h, int out_w, int kh, int kw,
                         int up_x, int up_y, int down_x, int down_y, int px0, int py0) {
  const int64_t total = planes * (int64_t)out_h * out_w;
  const int uh = in_h * up_y, uw = in_w * up_x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % out_w);
    const int oy = (int)((idx / out_w) % out_h);
    const int64_t p = idx / ((int64_t)out_w * out_h);
    const T* xp = x + p * (int64_t)in_h * in_w;
    const int by = oy * down_y - py0, bx = ox * down_x - px0;
    float acc = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
      const int uy = by + ky;
      if (uy < 0 || uy >= uh || (uy % up_y) != 0) continue;
      const int iy = uy / up_y;
      for (int kx = 0; kx < kw; ++kx) {
        const int ux = bx + kx;
        if (ux < 0 || ux >= uw || (ux % up_x) != 0) continue;
        acc = fmaf(taps.k[ky * kw + kx], to_f32(xp[(int64_t)iy * in_w + ux / up_x]), acc);
      }
    }
    y[idx] = from_f32<T>(acc);
  }
}

// Tile: (4*TPX) x (4*TPY) outputs per block of TPX*TPY threads; each thread a 4x4 patch.
template <typename T, int TPX, int TPY>
__global__ void __launch_bounds__(TPX* TPY)
blur4_sep_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t planes, int in_h, int in_w, int out_h,
                 int out_w, int py0, int px0, float4 fv, float4 fh) {
  constexpr int TW = 4 * TPX, TH = 4 * TPY;
  constexpr int SW = TW + 4;  // 3 halo columns + 1 pad: keeps every row 16-byte aligned
  constexpr int SH = TH + 3;
  __shared__ __align__(16) float tile[SH][SW];
  const int tid = threadIdx.x;
  const int tx = tid % TPX, ty = tid / TPX;
  const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
  const float fvv[4] = {fv.x, fv.y, fv.z, fv.w};
  const float fhh[4] = {fh.x, fh.y, fh.z, fh.w};

  for (int64_t p = blockIdx.z; p < planes; p += gridDim.z) {
    const T* xp = x + p * (int64_t)in_h * in_w;
    // stage the input tile; out-of-image samples are the zero padding of upfirdn2d.py:32-34
    for (int e = tid; e < SH * SW; e += TPX * TPY) {
      const int r = e / SW, c = e % SW;
      const int iy = oy0 - py0 + r, ix = ox0 - px0 + c;
      float v = 0.f;
      if (iy >= 0 && iy < in_h && ix >= 0 && ix < in_w) v = to_f32(xp[(int64_t)iy * in_w + ix]);
      tile[r][c] = v;
    }
    __syncthreads();

    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

#pragma unroll
    for (int r = 0; r < 7; ++r) {
      const float4 lo = *reinterpret_cast<const float4*>(&tile[ty * 4 + r][tx * 4]);
      const float4 hi = *reinterpret_cast<const float4*>(&tile[ty * 4 + r][tx * 4 + 4]);
      const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      float h[4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        h[c] = fhh[0] * v[c] + fhh[1] * v[c + 1] + fhh[2] * v[c + 2] + fhh[3] * v[c + 3];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int orow = r - j;  // output row inside the patch fed by input row r through tap j
        if (orow >= 0 && orow < 4) {
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[orow][c] = fmaf(fvv[j], h[c], acc[orow][c]);
        }
      }
    }
    __syncthreads();

    T* yp = y + p * (int64_t)out_h * out_w;
    const int ox = ox0 + tx * 4;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int oy = oy0 + ty * 4 + a;
      if (oy >= out_h || ox >= out_w) continue;
      T* dst = yp + (int64_t)oy * out_w + ox;
      if (sizeof(T) == 4 && ox + 3 < out_w && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (ox + c < out_w) dst[c] = from_f32<T>(acc[a][c]);
      }
    }
  }
}

static bool separable4(const float* k, float* kv, float* kh_) {
  int bi = 0, bj = 0;
  float best = 0.f;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (fabsf(k[i * 4 + j]) > best) best = fabsf(k[i * 4 + j]), bi = i, bj = j;
  if (best == 0.f) return false;
  for (int i = 0; i < 4; ++i) kv[i] = k[i * 4 + bj];
  for (int j = 0; j < 4; ++j) kh_[j] = k[bi * 4 + j] / k[bi * 4 + bj];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (fabsf(k[i * 4 + j] - kv[i] * kh_[j]) > 1e-6f * best) return false;
  return true;
}

template <typename T>
static int upfirdn2d_launch(const T* x, T* y, const float* k /*unflipped host*/, int64_t planes, int in_h,
                            int in_w, int kh, int kw, int up_x, int up_y, int down_x, int down_y, int px0,
                            int px1, int py0, int py1, cudaStream_t stream) {
  const int out_h = (in_h * up_y + py0 + py1 - kh) / down_y + 1;
  const int out_w = (in_w * up_x + px0 + px1 - kw) / down_x + 1;
  W2E_CHECK_ARG(out_h > 0 && out_w > 0, "upfirdn2d: empty output (%d x %d)", out_h, out_w);
  if (planes == 0) return W2E_OK;
  float kv[4], kh_[4];
  if (kh == 4 && kw == 4 && up_x == 1 && up_y == 1 && down_x == 1 && down_y == 1 && separable4(k, kv, kh_)) {
    // flipped taps: out[oy] = sum_ky P[oy+ky] * k[3-ky]
    const float4 fv = make_float4(kv[3], kv[2], kv[1], kv[0]);
    const float4 fh = make_float4(kh_[3], kh_[2], kh_[1], kh_[0]);
    if (out_w >= 48 && out_h >= 48) {
      dim3 grid(ceil_div(out_w, 64), ceil_div(out_h, 64), (unsigned)(planes < 32768 ? planes : 32768));
      blur4_sep_kernel<T, 16, 16><<<grid, 256, 0, stream>>>(x, y, planes, in_h, in_w, out_h, out_w, py0, px0, fv, fh);
    } else {
      dim3 grid(ceil_div(out_w, 32), ceil_div(out_h, 16), (unsigned)(planes < 32768 ? planes : 32768));
      blur4_sep_kernel<T, 8, 4><<<grid, 32, 0, stream>>>(x, y, planes, in_h, in_w, out_h, out_w, py0, px0, fv, fh);
    }
    W2E_LAUNCH_OK();
    return W2E_OK;
  }
  W2E_CHECK_ARG(kh * kw <= 64, "upfirdn2d: kernel %dx%d larger than the 64-tap limit", kh, kw);
  FirTaps taps;
  for (int i = 0; i < kh; ++i)
    for (int j = 0; j < kw; ++j) taps.k[i * kw + j] = k[(kh - 1 - i) * kw + (kw - 1 - j)];
  const int64_t total = planes * (int64_t)out_h * out_w;
  const int64_t blocks = ceil_div64(total, 256);
  const int64_t cap = (int64_t)sm_count() * 32;
  upfirdn2d_generic_kernel<T><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, stream>>>(
      x, y, taps, planes, in_h, in_w, out_h, out_w, kh, kw, up_x, up_y, down_x, down_y, px0, py0);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

}  // namespace w2e

using namespace w2e;

extern "C" int w2e_upfirdn2d_fwd(const void* x, void* y, const float* host_taps, int64_t planes, int in_h,
                                 int in_w, int kh, int kw, int up_x, int up_y, int down_x, int down_y, int px0,
                                 int px1, int py0, int py1, int dtype, void* stream) {
  W2E_CHECK_ARG(x && y && host_taps, "upfirdn2d: null pointer");
  W2E_CHECK_ARG(planes >= 0 && in_h > 0 && in_w > 0 && kh > 0 && kw > 0, "upfirdn2d: bad shape");
  W2E_CHECK_ARG(up_x > 0 && up_y > 0 && down_x > 0 && down_y > 0, "upfirdn2d: up/down must be positive");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == W2E_F32)
    return upfirdn2d_launch<float>((const float*)x, (float*)y, host_taps, planes, in_h, in_w, kh, kw, up_x, up_y,
                                   down_x, down_y, px0, px1, py0, py1, s);
  if (dtype == W2E_BF16)
    return upfirdn2d_launch<__nv_bfloat16>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, host_taps, planes, in_h,
                                           in_w, kh, kw, up_x, up_y, down_x, down_y, px0, px1, py0, py1, s);
  return set_error(W2E_ERR_INVALID, "upfirdn2d: unknown dtype %d", dtype);
}

extern "C" int w2e_upfirdn2d_bwd(const void* gy, void* gx, const float* host_taps, int64_t planes, int in_h,
                                 int in_w, int kh, int kw, int up_x, int up_y, int down_x, int down_y, int px0,
                                 int px1, int py0, int py1, int dtype, void* stream) {
  W2E_CHECK_ARG(gy && gx && host_taps, "upfirdn2d_bwd: null pointer");
  W2E_CHECK_ARG(kh > 0 && kw > 0 && kh * kw <= 64, "upfirdn2d_bwd: bad kernel size");
  const int out_h = (in_h * up_y + py0 + py1 - kh) / down_y + 1;
  const int out_w = (in_w * up_x + px0 + px1 - kw) / down_x + 1;
  float flipped[64];
  for (int i = 0; i < kh; ++i)
    for (int j = 0; j < kw; ++j) flipped[i * kw + j] = host_taps[(kh - 1 - i) * kw + (kw - 1 - j)];
  const int gpx0 = kw - px0 - 1, gpy0 = kh - py0 - 1;
  const int gpx1 = in_w * up_x - out_w * down_x + px0 - up_x + 1;
  const int gpy1 = in_h * up_y - out_h * down_y + py0 - up_y + 1;
  // the transposed operator: upsample by `down`, filter with the flipped kernel, decimate by `up`
  const int chk_h = (out_h * down_y + gpy0 + gpy1 - kh) / up_y + 1;
  const int chk_w = (out_w * down_x + gpx0 + gpx1 - kw) / up_x + 1;
  W2E_CHECK_ARG(chk_h == in_h && chk_w == in_w, "upfirdn2d_bwd: geometry mismatch (%d,%d) vs (%d,%d)", chk_h,
                chk_w, in_h, in_w);
  return w2e_upfirdn2d_fwd(gy, gx, flipped, planes, out_h, out_w, kh, kw, down_x, down_y, up_x, up_y, gpx0, gpx1,
                           gpy0, gpy1, dtype, stream);
}
