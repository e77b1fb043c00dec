// All style modulations and all demodulation coefficients of one Generator.forward in TWO launches
// (instead of ~3 torch kernels per EqualLinear and one demod kernel per conv):
//  * style_mod_all   s_l[b, :] = W_l * scale . latent[b, row_l, :] + bias_l * lr_mul for every
//                    modulated layer l (models/stylegan2/model.py:130-159 called from :238)
//  * style_demod_all d_l[b, o] = rsqrt(sum_i s_l[b, i]^2 * wsq_l[o, i] + 1e-8)      (model.py:242)
// The per-layer tensors are concatenated once by the host (cached until a weight changes); results
// are laid out per layer as contiguous [B, C_l] blocks so that the convolution kernels index them
// exactly like the per-layer tensors.
#include "common.cuh"

namespace w2e {

// meta[l] = {latent row, first concatenated channel, Cin}.  One block = 32 channels x kModB samples:
// the latent rows of the samples are staged in shared memory and every weight row is read once per
// kModB samples (the weights, 18 MB at 1024^2, are the only real traffic of this kernel).
constexpr int kModB = 8;
__global__ void __launch_bounds__(256)
style_mod_all_kernel(const float* __restrict__ latent, int64_t stride_b, int64_t stride_row, const float* __restrict__ w_all,
                     const float* __restrict__ b_all, const int* __restrict__ block_layer, const int4* __restrict__ meta,
                     float* __restrict__ s_all, int B, int D) {
  extern __shared__ float lat_s[];   // [kModB][D]
  const int blk = blockIdx.x, b0 = blockIdx.y * kModB;
  const int4 m = __ldg(meta + __ldg(block_layer + blk));   // x = latent row, y = channel offset, z = Cin
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < kModB * D; e += blockDim.x) {
    const int bb = e / D, i = e - bb * D;
    lat_s[e] = (b0 + bb < B) ? __ldg(latent + (int64_t)(b0 + bb) * stride_b + (int64_t)m.x * stride_row + i) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int gc = blk * 32 + warp * 4 + j;     // concatenated channel
    const float* w = w_all + (int64_t)gc * D;
    float acc[kModB];
#pragma unroll
    for (int bb = 0; bb < kModB; ++bb) acc[bb] = 0.f;
    for (int i = lane; i < D; i += 32) {
      const float wv = __ldg(w + i);
#pragma unroll
      for (int bb = 0; bb < kModB; ++bb) acc[bb] = fmaf(wv, lat_s[bb * D + i], acc[bb]);
    }
    const float bias = __ldg(b_all + gc);
#pragma unroll
    for (int bb = 0; bb < kModB; ++bb) {
      const float v = warp_sum(acc[bb]);
      if (lane == 0 && b0 + bb < B) s_all[(int64_t)B * m.y + (int64_t)(b0 + bb) * m.z + (gc - m.y)] = v + bias;
    }
  }
}

// meta[l] = {s offset of the layer (floats, batch 1), Cin, Cout, wsq offset}; dmeta[l] = d channel offset
__global__ void __launch_bounds__(256)
style_demod_all_kernel(const float* __restrict__ s_all, const float* __restrict__ wsq_all, const int* __restrict__ block_layer,
                       const int4* __restrict__ meta, const int* __restrict__ d_off, float* __restrict__ d_all, int B) {
  const int blk = blockIdx.x, b = blockIdx.y;
  const int l = __ldg(block_layer + blk);
  const int4 m = __ldg(meta + l);
  const int doff = __ldg(d_off + l);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blk * 8 + warp - doff;            // output channel inside the layer
  const float* s = s_all + (int64_t)B * m.x + (int64_t)b * m.y;
  const float* w = wsq_all + m.w + (int64_t)o * m.y;
  float acc = 0.f;
  for (int i = lane; i < m.y; i += 32) {
    const float v = __ldg(s + i);
    acc = fmaf(v * v, __ldg(w + i), acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) d_all[(int64_t)B * doff + (int64_t)b * m.z + o] = rsqrtf(acc + 1e-8f);
}

}  // namespace w2e

using namespace w2e;

extern "C" int w2e_style_mod_all(const float* latent, int64_t stride_b, int64_t stride_row, const float* w_all,
                                 const float* b_all, const int* block_layer, const int* meta4, float* s_all, int B, int D,
                                 int nblocks, void* stream) {
  W2E_CHECK_ARG(latent && w_all && b_all && block_layer && meta4 && s_all, "style_mod_all: null pointer");
  W2E_CHECK_ARG(B >= 0 && D > 0 && D <= 1024 && nblocks >= 0, "style_mod_all: bad shape (style_dim <= 1024)");
  if (B == 0 || nblocks == 0) return W2E_OK;
  style_mod_all_kernel<<<dim3((unsigned)nblocks, (unsigned)ceil_div(B, kModB)), 256, (size_t)kModB * D * sizeof(float),
                         (cudaStream_t)stream>>>(
      latent, stride_b, stride_row, w_all, b_all, block_layer, reinterpret_cast<const int4*>(meta4), s_all, B, D);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_style_demod_all(const float* s_all, const float* wsq_all, const int* block_layer, const int* meta4,
                                   const int* d_off, float* d_all, int B, int nblocks, void* stream) {
  W2E_CHECK_ARG(s_all && wsq_all && block_layer && meta4 && d_off && d_all, "style_demod_all: null pointer");
  W2E_CHECK_ARG(B >= 0 && B <= 65535 && nblocks >= 0, "style_demod_all: bad shape");
  if (B == 0 || nblocks == 0) return W2E_OK;
  style_demod_all_kernel<<<dim3((unsigned)nblocks, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(
      s_all, wsq_all, block_layer, reinterpret_cast<const int4*>(meta4), d_off, d_all, B);
  W2E_LAUNCH_OK();
  return W2E_OK;
}
