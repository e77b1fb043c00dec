// bias + (noise) + leaky-ReLU * gain -- replaces models/stylegan2/op/fused_act.py:23-39 and, with
// the optional noise operand, NoiseInjection (models/stylegan2/model.py:279-290) in one pass.
// HBM-bound elementwise: algorithmic bytes = 2 * numel * sizeof(T).
#include "common.cuh"

namespace w2e {

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
bias_act_fwd_kernel(const T* __restrict__ x, const float* __restrict__ bias, const float* __restrict__ noise,
                    const float* __restrict__ noise_w, int noise_per_outer, T* __restrict__ y, int64_t outer,
                    int C, int64_t inner, float slope, float gain) {
  const int64_t total = outer * C * inner / VEC;
  const float nw = (noise != nullptr) ? __ldg(noise_w) : 0.f;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e0 = v * VEC;
    T in[VEC], out[VEC];
    if (VEC == 4 && sizeof(T) == 4) {
      *reinterpret_cast<float4*>(in) = *reinterpret_cast<const float4*>(x + e0);
    } else if (VEC == 8 && sizeof(T) == 2) {
      *reinterpret_cast<uint4*>(in) = *reinterpret_cast<const uint4*>(x + e0);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) in[i] = x[e0 + i];
    }
    // VEC > 1 is only used when inner % VEC == 0 (one channel per vector) or inner == 1 && C % VEC == 0
    if (inner == 1) {
      const int c0 = (int)(e0 % C);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float val = to_f32(in[i]) + (bias ? __ldg(bias + c0 + i) : 0.f);
        if (noise) val = fmaf(nw, __ldg(noise + (noise_per_outer ? (e0 / C) : 0)), val);
        out[i] = from_f32<T>(lrelu_gain(val, slope, gain));
      }
    } else {
      const int64_t pix = e0 % inner;
      const int64_t oc = e0 / inner;  // outer * C + c
      const int c = (int)(oc % C);
      const float b = bias ? __ldg(bias + c) : 0.f;
      const float* nrow = noise ? noise + (noise_per_outer ? (oc / C) * inner : 0) + pix : nullptr;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float val = to_f32(in[i]) + b;
        if (noise) val = fmaf(nw, __ldg(nrow + i), val);
        out[i] = from_f32<T>(lrelu_gain(val, slope, gain));
      }
    }
    if (VEC == 4 && sizeof(T) == 4) {
      *reinterpret_cast<float4*>(y + e0) = *reinterpret_cast<float4*>(out);
    } else if (VEC == 8 && sizeof(T) == 2) {
      *reinterpret_cast<uint4*>(y + e0) = *reinterpret_cast<uint4*>(out);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) y[e0 + i] = out[i];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bias_act_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ y, T* __restrict__ gx, int64_t total,
                    float slope, float gain) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const float g = to_f32(gy[e]);
    gx[e] = from_f32<T>(to_f32(y[e]) > 0.f ? g * gain : g * (gain * slope));
  }
}

// gbias stage 1: partial[chunk][c] = sum over a slab of `outer` rows (and all inner) of gx[.., c, ..]
// grid = (C, chunks); fixed order inside a block => deterministic.
template <typename T>
__global__ void __launch_bounds__(256)
bias_grad_partial_kernel(const T* __restrict__ gx, float* __restrict__ partial, int64_t outer, int C, int64_t inner,
                         int64_t rows_per_chunk) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  const int64_t o0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t o1 = (o0 + rows_per_chunk < outer) ? o0 + rows_per_chunk : outer;
  float acc = 0.f;
  const int64_t span = (o1 - o0) * inner;
  for (int64_t t = threadIdx.x; t < span; t += blockDim.x) {
    const int64_t o = o0 + t / inner, p = t % inner;
    acc += to_f32(gx[(o * C + c) * inner + p]);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[(int64_t)blockIdx.y * C + c] = acc;
}

__global__ void bias_grad_final_kernel(const float* __restrict__ partial, float* __restrict__ gbias, int C,
                                       int chunks) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int k = 0; k < chunks; ++k) acc += partial[(int64_t)k * C + c];
  gbias[c] = acc;
}

static int bias_chunks(int64_t outer, int C, int64_t inner) {
  // enough blocks to fill the machine, never more than `outer` slabs
  int64_t want = ceil_div64((int64_t)sm_count() * 8, C);
  if (want < 1) want = 1;
  if (want > outer) want = outer;
  if (want > 1024) want = 1024;
  (void)inner;
  return (int)want;
}

template <typename T>
static int fwd_launch(const T* x, const float* bias, const float* noise, const float* noise_w, int noise_batch,
                      T* y, int64_t outer, int C, int64_t inner, float slope, float gain, cudaStream_t s) {
  const int64_t total = outer * C * inner;
  if (total == 0) return W2E_OK;
  const int per_outer = (noise && noise_batch != 1) ? 1 : 0;
  constexpr int V = 16 / sizeof(T);
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  const bool vec_ok = aligned && ((inner == 1) ? (C % V == 0) : (inner % V == 0));
  const int64_t cap = (int64_t)sm_count() * 16;
  if (vec_ok) {
    const int64_t blocks = ceil_div64(total / V, 256);
    bias_act_fwd_kernel<T, V><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(
        x, bias, noise, noise_w, per_outer, y, outer, C, inner, slope, gain);
  } else {
    const int64_t blocks = ceil_div64(total, 256);
    bias_act_fwd_kernel<T, 1><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(
        x, bias, noise, noise_w, per_outer, y, outer, C, inner, slope, gain);
  }
  W2E_LAUNCH_OK();
  return W2E_OK;
}

template <typename T>
static int bwd_launch(const T* gy, const T* y, T* gx, float* gbias, void* workspace, int64_t outer, int C,
                      int64_t inner, float slope, float gain, cudaStream_t s) {
  const int64_t total = outer * C * inner;
  if (total == 0) return W2E_OK;
  const int64_t cap = (int64_t)sm_count() * 16;
  const int64_t blocks = ceil_div64(total, 256);
  bias_act_bwd_kernel<T><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(gy, y, gx, total, slope, gain);
  W2E_LAUNCH_OK();
  if (gbias) {
    W2E_CHECK_ARG(workspace != nullptr, "bias_act_bwd: workspace required for the bias gradient");
    const int chunks = bias_chunks(outer, C, inner);
    const int64_t rows = ceil_div64(outer, chunks);
    float* partial = (float*)workspace;
    bias_grad_partial_kernel<T><<<dim3(C, chunks), 256, 0, s>>>(gx, partial, outer, C, inner, rows);
    W2E_LAUNCH_OK();
    bias_grad_final_kernel<<<ceil_div(C, 128), 128, 0, s>>>(partial, gbias, C, chunks);
    W2E_LAUNCH_OK();
  }
  return W2E_OK;
}

}  // namespace w2e

using namespace w2e;

extern "C" int w2e_bias_act_fwd(const void* x, const float* bias, const float* noise, const float* noise_w,
                                int noise_batch, void* y, int64_t outer, int C, int64_t inner, float slope,
                                float scale, int dtype, void* stream) {
  W2E_CHECK_ARG(x && y, "bias_act: null pointer");
  W2E_CHECK_ARG(outer >= 0 && C > 0 && inner > 0, "bias_act: bad shape");
  W2E_CHECK_ARG(noise == nullptr || noise_w != nullptr, "bias_act: noise given without its weight");
  W2E_CHECK_ARG(noise == nullptr || noise_batch == 1 || noise_batch == outer, "bias_act: noise batch %d", noise_batch);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == W2E_F32)
    return fwd_launch<float>((const float*)x, bias, noise, noise_w, noise_batch, (float*)y, outer, C, inner, slope,
                             scale, s);
  if (dtype == W2E_BF16)
    return fwd_launch<__nv_bfloat16>((const __nv_bfloat16*)x, bias, noise, noise_w, noise_batch, (__nv_bfloat16*)y,
                                     outer, C, inner, slope, scale, s);
  return set_error(W2E_ERR_INVALID, "bias_act: unknown dtype %d", dtype);
}

extern "C" int64_t w2e_bias_act_bwd_workspace(int64_t outer, int C, int64_t inner) {
  return (int64_t)bias_chunks(outer, C, inner) * C * (int64_t)sizeof(float);
}

extern "C" int w2e_bias_act_bwd(const void* gy, const void* y, void* gx, float* gbias, void* workspace,
                                int64_t outer, int C, int64_t inner, float slope, float scale, int dtype,
                                void* stream) {
  W2E_CHECK_ARG(gy && y && gx, "bias_act_bwd: null pointer");
  W2E_CHECK_ARG(outer >= 0 && C > 0 && inner > 0, "bias_act_bwd: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == W2E_F32)
    return bwd_launch<float>((const float*)gy, (const float*)y, (float*)gx, gbias, workspace, outer, C, inner,
                             slope, scale, s);
  if (dtype == W2E_BF16)
    return bwd_launch<__nv_bfloat16>((const __nv_bfloat16*)gy, (const __nv_bfloat16*)y, (__nv_bfloat16*)gx, gbias,
                                     workspace, outer, C, inner, slope, scale, s);
  return set_error(W2E_ERR_INVALID, "bias_act_bwd: unknown dtype %d", dtype);
}
