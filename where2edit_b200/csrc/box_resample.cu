// CLIP / VGG pre-resample: AvgPool2d(P)(Upsample(scale_factor=U, nearest)(image)),
// criteria/clip_loss.py:10-14, criteria/perceptual_loss.py:12-17, called on every generated image right after
// Generator.forward (attention/run_attention.py:1163,1259).  The reference materialises the upsampled
// [B,3,U*H,U*W] tensor (617 MB fp32 per 1024^2 image at U=7) to produce 224x224; composed, every output pixel is
// a small weighted box of source pixels with integer weights
//     w(o, j) = |[U*j, U*j+U-1]  intersect  [P*o, P*o+P-1]|          (rows and columns separately),
// so one pass reads the image once and writes the 224x224 result; the backward is the transposed gather.
#include "common.cuh"

namespace w2e {

// overlap of upsampled rows [U*j, U*j+U-1] with the pooling window [P*o, P*o+P-1]
__device__ __forceinline__ int box_weight(int j, int o, int U, int P) {
  return min(U * j + U - 1, P * o + P - 1) - max(U * j, P * o) + 1;
}

__global__ void __launch_bounds__(256)
box_resample_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total, int H, int W, int OH, int OW,
                        int U, int P) {
  const float denom = (float)P * (float)P;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW), oy = (int)((i / OW) % OH);
    const int64_t n = i / ((int64_t)OW * OH);
    const float* xp = x + n * H * W;
    const int j0 = (P * oy) / U, j1 = (P * oy + P - 1) / U;
    const int i0 = (P * ox) / U, i1 = (P * ox + P - 1) / U;
    float acc = 0.f;
    for (int j = j0; j <= j1; ++j) {
      const float wy = (float)box_weight(j, oy, U, P);
      float row = 0.f;
      for (int c = i0; c <= i1; ++c) row = fmaf((float)box_weight(c, ox, U, P), __ldg(xp + (int64_t)j * W + c), row);
      acc = fmaf(wy, row, acc);
    }
    y[i] = __fdiv_rn(acc, denom);
  }
}

// gx[n, j, c] = (1/P^2) * sum_{oy, ox} w(oy, j) * w(ox, c) * gy[n, oy, ox]; upsampled rows beyond OH*P (AvgPool's floor)
// receive nothing.  One thread writes 4 consecutive source columns (16-byte stores when W % 4 == 0).
__global__ void __launch_bounds__(256)
box_resample_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, int64_t total4, int H, int W, int W4, int OH,
                        int OW, int U, int P, int vec) {
  const float inv = 1.f / ((float)P * (float)P);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % W4), j = (int)((i / W4) % H);
    const int64_t n = i / ((int64_t)W4 * H);
    const float* gp = gy + n * OH * OW;
    const int oy0 = (U * j) / P, oy1 = min((U * j + U - 1) / P, OH - 1);
    float out[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = c4 * 4 + e;
      float acc = 0.f;
      if (c < W) {
        const int ox0 = (U * c) / P, ox1 = min((U * c + U - 1) / P, OW - 1);
        for (int oy = oy0; oy <= oy1; ++oy) {
          const float wy = (float)box_weight(j, oy, U, P);
          float row = 0.f;
          for (int ox = ox0; ox <= ox1; ++ox) row = fmaf((float)box_weight(c, ox, U, P), __ldg(gp + (int64_t)oy * OW + ox), row);
          acc = fmaf(wy, row, acc);
        }
      }
      out[e] = acc * inv;
    }
    float* dst = gx + (n * H + j) * W + c4 * 4;
    if (vec) {
      *reinterpret_cast<float4*>(dst) = make_float4(out[0], out[1], out[2], out[3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (c4 * 4 + e < W) dst[e] = out[e];
    }
  }
}

// ---- row kernels: the fast path (fully coalesced global traffic) ---------------------------------------------
// Forward: one CTA per (plane, output row).  The <= ceil(P/U)+1 source rows of the window are read with 16-byte
// loads (all issued before the first use) and reduced vertically into a shared-memory row; OW threads then
// take the horizontal boxes from shared memory.
constexpr int kRowChunk = 8;   // source rows in flight per thread

__global__ void __launch_bounds__(256)
box_resample_fwd_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int OH, int OW, int U, int P) {
  extern __shared__ float vs[];   // [W] vertically reduced row
  const int oy = blockIdx.x % OH;
  const int64_t n = blockIdx.x / OH;
  const int j0 = (P * oy) / U, j1 = (P * oy + P - 1) / U;
  const float* xp = x + n * H * W;
  const int W4 = W >> 2;
  for (int c4 = threadIdx.x; c4 < W4; c4 += 256) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int jb = j0; jb <= j1; jb += kRowChunk) {
      float4 v[kRowChunk];
#pragma unroll
      for (int r = 0; r < kRowChunk; ++r)
        if (jb + r <= j1) v[r] = __ldg(reinterpret_cast<const float4*>(xp + (int64_t)(jb + r) * W) + c4);
#pragma unroll
      for (int r = 0; r < kRowChunk; ++r)
        if (jb + r <= j1) {
          const float wy = (float)box_weight(jb + r, oy, U, P);
          acc.x = fmaf(wy, v[r].x, acc.x);
          acc.y = fmaf(wy, v[r].y, acc.y);
          acc.z = fmaf(wy, v[r].z, acc.z);
          acc.w = fmaf(wy, v[r].w, acc.w);
        }
    }
    *reinterpret_cast<float4*>(vs + 4 * c4) = acc;
  }
  __syncthreads();
  const float denom = (float)P * (float)P;
  float* yp = y + (n * OH + oy) * OW;
  for (int ox = threadIdx.x; ox < OW; ox += 256) {
    const int i0 = (P * ox) / U, i1 = (P * ox + P - 1) / U;
    float acc = 0.f;
    for (int c = i0; c <= i1; ++c) acc = fmaf((float)box_weight(c, ox, U, P), vs[c], acc);
    yp[ox] = __fdiv_rn(acc, denom);
  }
}

// Backward (U <= P: a source pixel's U upsampled copies span at most two pooling windows per axis): one CTA per
// (plane, group of kBwdRows source rows); a thread owns 4 consecutive source columns, whose two window indices
// and weights are computed once, then walks the rows: <= 2 x 2 gradient values per element straight from L1/L2
// (the gradient image is 19 MB), one 16-byte store per row.  No shared memory, no barriers.
constexpr int kBwdRows = 16;

__global__ void __launch_bounds__(256)
box_resample_bwd_rows_kernel(const float* __restrict__ gy, float* __restrict__ gx, int H, int W, int OH, int OW, int U, int P) {
  const int groups = (H + kBwdRows - 1) / kBwdRows;
  const int jg = blockIdx.x % groups;
  const int64_t n = blockIdx.x / groups;
  const float* gp = gy + n * OH * OW;
  const float inv = 1.f / ((float)P * (float)P);
  const int W4 = W >> 2;
  const int j_end = min(H, (jg + 1) * kBwdRows);
  for (int c4 = threadIdx.x; c4 < W4; c4 += 256) {
    int xa[4], xb[4];
    float wa[4], wb[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * c4 + e;
      const int o0 = (U * c) / P, o1 = (U * c + U - 1) / P;
      wa[e] = o0 < OW ? (float)box_weight(c, o0, U, P) * inv : 0.f;
      wb[e] = (o1 != o0 && o1 < OW) ? (float)box_weight(c, o1, U, P) * inv : 0.f;
      xa[e] = min(o0, OW - 1);
      xb[e] = min(o1, OW - 1);
    }
    for (int j = jg * kBwdRows; j < j_end; ++j) {
      const int o0 = (U * j) / P, o1 = (U * j + U - 1) / P;
      const float wy0 = o0 < OH ? (float)box_weight(j, o0, U, P) : 0.f;
      const float wy1 = (o1 != o0 && o1 < OH) ? (float)box_weight(j, o1, U, P) : 0.f;
      const float* g0 = gp + (int64_t)min(o0, OH - 1) * OW;
      float out[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) out[e] = wy0 * fmaf(wa[e], __ldg(g0 + xa[e]), wb[e] * __ldg(g0 + xb[e]));
      if (wy1 != 0.f) {   // uniform over the CTA: this source row straddles two pooling windows
        const float* g1 = gp + (int64_t)o1 * OW;
#pragma unroll
        for (int e = 0; e < 4; ++e) out[e] = fmaf(wy1, fmaf(wa[e], __ldg(g1 + xa[e]), wb[e] * __ldg(g1 + xb[e])), out[e]);
      }
      *reinterpret_cast<float4*>(gx + (n * H + j) * W + 4 * c4) = make_float4(out[0], out[1], out[2], out[3]);
    }
  }
}

static int resample_dims(int H, int W, int U, int P, int* OH, int* OW) {
  *OH = (int)(((int64_t)H * U) / P);
  *OW = (int)(((int64_t)W * U) / P);
  return *OH > 0 && *OW > 0;
}

}  // namespace w2e

using namespace w2e;

extern "C" int w2e_box_resample_fwd(const float* x, float* y, int64_t planes, int H, int W, int up, int pool, void* stream) {
  W2E_CHECK_ARG(x && y, "box_resample_fwd: null pointer");
  W2E_CHECK_ARG(planes >= 0 && H > 0 && W > 0 && up >= 1 && pool >= 1 && (int64_t)H * up < (1 << 30) && (int64_t)W * up < (1 << 30),
                "box_resample_fwd: bad shape");
  int OH, OW;
  W2E_CHECK_ARG(resample_dims(H, W, up, pool, &OH, &OW), "box_resample_fwd: pooling window larger than the upsampled image");
  const int64_t total = planes * OH * OW;
  if (total == 0) return W2E_OK;
  const bool rows_ok = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && W <= 8192 && planes * OH < (1ll << 31);
  if (rows_ok) {   // coalesced row kernel (every configuration of the reference)
    box_resample_fwd_rows_kernel<<<(unsigned)(planes * OH), 256, (size_t)W * sizeof(float), (cudaStream_t)stream>>>(
        x, y, H, W, OH, OW, up, pool);
  } else {         // any shape / alignment
    box_resample_fwd_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 148 * 32), 256, 0, (cudaStream_t)stream>>>(
        x, y, total, H, W, OH, OW, up, pool);
  }
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_box_resample_bwd(const float* gy, float* gx, int64_t planes, int H, int W, int up, int pool, void* stream) {
  W2E_CHECK_ARG(gy && gx, "box_resample_bwd: null pointer");
  W2E_CHECK_ARG(planes >= 0 && H > 0 && W > 0 && up >= 1 && pool >= 1 && (int64_t)H * up < (1 << 30) && (int64_t)W * up < (1 << 30),
                "box_resample_bwd: bad shape");
  int OH, OW;
  W2E_CHECK_ARG(resample_dims(H, W, up, pool, &OH, &OW), "box_resample_bwd: pooling window larger than the upsampled image");
  const int W4 = ceil_div(W, 4);
  const int64_t total4 = planes * H * W4;
  if (total4 == 0) return W2E_OK;
  const int vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(gx) & 15) == 0);
  const int64_t groups = ceil_div(H, kBwdRows);
  if (vec && up <= pool && planes * groups < (1ll << 31)) {
    box_resample_bwd_rows_kernel<<<(unsigned)(planes * groups), 256, 0, (cudaStream_t)stream>>>(gy, gx, H, W, OH, OW, up, pool);
  } else {
    box_resample_bwd_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total4, 256), 148 * 32), 256, 0, (cudaStream_t)stream>>>(
        gy, gx, total4, H, W, W4, OH, OW, up, pool, vec);
  }
  W2E_LAUNCH_OK();
  return W2E_OK;
}
