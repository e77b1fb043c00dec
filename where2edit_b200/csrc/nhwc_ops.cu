// Channels-last bf16 kernels around the tensor-core modulated convolution (modconv_tc.cu).
// All are HBM-bound; every activation is touched once per kernel with 128-bit accesses.
//
//  * nchw_to_nhwc_mod   fp32 NCHW (constant input / caller feature maps) -> bf16 NHWC, times the
//                       consumer's style (modulation folded into the producer, model.py:238-239)
//  * nhwc_to_nchw_f32   bf16 NHWC -> fp32 NCHW (feature capture, attention_model.py:542-543)
//  * blur_act_nhwc      Blur after the up-convolution (model.py:200-206,260 = upfirdn2d with the
//                       4x4 separable kernel, pad (1,1)) fused with NoiseInjection + FusedLeakyReLU
//                       (model.py:279-290, op/fused_act.py:23-39) and the next layer's modulation.
//                       Algorithmic bytes: B*C*((H+1)*(W+1) + H*W*n_out)*2.
//  * torgb_nhwc         ToRGB (model.py:353-362): 1x1 modulated conv to 3 channels + bias +
//                       polyphase x2 skip upsample + add, reading bf16 NHWC, writing fp32 NCHW.
//  * blend_nhwc         region-mask blend (attention_model.py:548-549) on bf16 NHWC.
#include "tc_ptx.cuh"

namespace w2e {

struct __align__(16) bf16x8 {
  __nv_bfloat162 v[4];
};

// 128-bit global accesses (a struct copy of four bfloat162 compiles to four 32-bit loads)
__device__ __forceinline__ bf16x8 ld8(const __nv_bfloat16* p) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  bf16x8 r;
  *reinterpret_cast<uint4*>(&r) = u;
  return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const bf16x8& v) {
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v);
}

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}

// Operand element of the tensor-core convolution: bf16 (round to nearest even), or fp32 rounded to tf32 with
// cvt.rna (the tf32 MMA TRUNCATES the low 13 mantissa bits of its operands; truncation is a biased error that grows
// coherently through the 17-layer chain -- measured: un-rounded tf32 was no closer to the reference than bf16)
template <typename T>
__device__ __forceinline__ T to_operand(float v);
template <>
__device__ __forceinline__ __nv_bfloat16 to_operand<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ float to_operand<float>(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

// ------------------------------------------------------------------------------ layout transforms
// x [Bx, C, HW] fp32 (Bx = 1 broadcasts over the batch) -> y [B, HW, C] bf16 * style[b, c]
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_mod_kernel(const float* __restrict__ x, const float* __restrict__ style, T* __restrict__ y,
                        int Bx, int C, int64_t HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const float* xb = x + (Bx == 1 ? 0 : (int64_t)b * C * HW);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;
    const int64_t p = p0 + tx;
    tile[r][tx] = (c < C && p < HW) ? __ldg(xb + (int64_t)c * HW + p) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t p = p0 + r;
    const int c = c0 + tx;
    if (p < HW && c < C) {
      float v = tile[tx][r];
      if (style) v *= __ldg(style + (int64_t)b * C + c);
      y[((int64_t)b * HW + p) * C + c] = to_operand<T>(v);
    }
  }
}

// x [B, HW, C] bf16 (or fp32) -> y [B, C, HW] fp32
template <typename T>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_f32_kernel(const T* __restrict__ x, float* __restrict__ y, int C, int64_t HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t p = p0 + r;
    const int c = c0 + tx;
    tile[r][tx] = (p < HW && c < C) ? to_f32<T>(x[((int64_t)b * HW + p) * C + c]) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;
    const int64_t p = p0 + tx;
    if (c < C && p < HW) y[((int64_t)b * C + c) * HW + p] = tile[tx][r];
  }
}

// The same transpose with 16-byte loads and 64 x 64 tiles (C % 64 == 0 or C == 32 -> CH = 32; bf16 input): every captured
// feature map of a forward goes through it when the caller asks for the reference's fp32 NCHW maps (25 GB of traffic at
// 1024^2, batch 32: as much time as the forward itself), and the 32 x 32 tile kernel above moves 64-byte rows.
template <int CH>
__global__ void __launch_bounds__(256)
nhwc_bf16_to_nchw_f32_v2_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int C, int64_t HW) {
  constexpr int PIX = 64, CG = CH / 8;
  __shared__ float tile[CH][PIX + 1];
  const int b = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * PIX;
  const int c0 = blockIdx.y * CH;
  const int tid = threadIdx.x;
#pragma unroll
  for (int q = tid; q < PIX * CG; q += 256) {
    const int px = q / CG, cg = q - px * CG;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (p0 + px < HW) v = __ldg(reinterpret_cast<const uint4*>(x + ((int64_t)b * HW + p0 + px) * C + c0 + cg * 8));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      tile[cg * 8 + 2 * j][px] = __uint_as_float(w[j] << 16);
      tile[cg * 8 + 2 * j + 1][px] = __uint_as_float(w[j] & 0xffff0000u);
    }
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  const bool full = p0 + PIX <= HW && (HW & 1) == 0;
#pragma unroll
  for (int c = warp; c < CH; c += 8) {
    float* dst = y + ((int64_t)b * C + c0 + c) * HW + p0;
    if (full) {
      *reinterpret_cast<float2*>(dst + 2 * lane) = make_float2(tile[c][2 * lane], tile[c][2 * lane + 1]);
    } else {
      if (p0 + 2 * lane < HW) dst[2 * lane] = tile[c][2 * lane];
      if (p0 + 2 * lane + 1 < HW) dst[2 * lane + 1] = tile[c][2 * lane + 1];
    }
  }
}

// Output-parity class (py, px) of a fp32 NCHW tensor x [B, C, H, W] as bf16 NHWC [B, hc, wc, C], times scale[b, c]:
// y[b, j, i, c] = x[b, c, 2j+py, 2i+px] * scale[b, c].  (The dgrad of the transposed x2 convolution consumes the
// upstream gradient class by class.)
template <typename T>
__global__ void __launch_bounds__(256)
nchw_class_to_nhwc_mod_kernel(const float* __restrict__ x, const float* __restrict__ scale, T* __restrict__ y,
                              int C, int H, int W, int hc, int wc, int py, int px) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int64_t HWc = (int64_t)hc * wc;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;
    const int64_t p = p0 + tx;
    float v = 0.f;
    if (c < C && p < HWc) {
      const int j = (int)(p / wc), i = (int)(p - (int64_t)j * wc);
      v = __ldg(x + (((int64_t)b * C + c) * H + (2 * j + py)) * W + (2 * i + px));
    }
    tile[r][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t p = p0 + r;
    const int c = c0 + tx;
    if (p < HWc && c < C) {
      float v = tile[tx][r];
      if (scale) v *= __ldg(scale + (int64_t)b * C + c);
      y[((int64_t)b * HWc + p) * C + c] = to_operand<T>(v);
    }
  }
}

// out[b, c, j, i] = sum over the four class results y_pq [B, h+1-p, w+1-q, C] (bf16 NHWC) at (j, i), fp32 NCHW [B, C, h, w]
template <typename T>
__global__ void __launch_bounds__(256)
nhwc_sum4_to_nchw_kernel(const T* __restrict__ y00, const T* __restrict__ y01, const T* __restrict__ y10,
                         const T* __restrict__ y11, float* __restrict__ out, int C, int h, int w) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int64_t HW = (int64_t)h * w;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t p = p0 + r;
    const int c = c0 + tx;
    float v = 0.f;
    if (p < HW && c < C) {
      const int j = (int)(p / w), i = (int)(p - (int64_t)j * w);
      v = to_f32<T>(y00[(((int64_t)b * (h + 1) + j) * (w + 1) + i) * C + c]) +
          to_f32<T>(y01[(((int64_t)b * (h + 1) + j) * w + i) * C + c]) +
          to_f32<T>(y10[(((int64_t)b * h + j) * (w + 1) + i) * C + c]) +
          to_f32<T>(y11[(((int64_t)b * h + j) * w + i) * C + c]);
    }
    tile[r][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;
    const int64_t p = p0 + tx;
    if (c < C && p < HW) out[((int64_t)b * C + c) * HW + p] = tile[tx][r];
  }
}

// ------------------------------------------------------------------------------ blur + noise + bias + act
// One thread = 8 channels of kBlurPx adjacent output columns, walking down kBlurRows output rows.
// Per input row it loads kBlurPx+3 pixels (128-bit each), filters them horizontally in registers
// and keeps the last four filtered rows in a register ring (compile-time slots, no moves); each
// new row completes one output row.  HBM sees every input element once (the 3-row / 3-column halo
// re-reads hit L1/L2).
struct BlurParams {
  const __nv_bfloat16* z;   // [B, IH, IW, C]
  const float* bias;        // [C] or null
  const float* noise;       // [noise_batch, H*W] or null
  const float* noise_w;     // device scalar
  const float* next_scale;  // [B, C] or null
  __nv_bfloat16* out;       // [B, H, W, C] or null
  __nv_bfloat16* out_mod;   // [B, H, W, C] or null
  int noise_per_sample;
  int B, C, IH, IW, H, W, py0, px0;
  int act;
  float fv[4], fh[4];       // flipped 1-D taps (vertical, horizontal)
};

// Tile = kBlurTY rows x TX columns x CT channels of the output, TX = 1024 / CT (CT = min(C, 128)):
// 128 threads = 2 row groups x (TX/2 column pairs) x (CT/8 channel groups).  The haloed input tile
// (kBlurTY+3) x (TX+3) x CT arrives in shared memory through ONE TMA box per CTA; out-of-image
// coordinates are zero-filled by the TMA unit, which is exactly upfirdn2d's padding
// (op/upfirdn2d.py:32-34), so the compute loop has no boundary logic.  Several CTAs per SM overlap
// one CTA's load with another's arithmetic.
constexpr int kBlurTY = 16;
constexpr int kBlurRG = 2;                    // row groups per tile
constexpr int kBlurRows = kBlurTY / kBlurRG;  // output rows walked by one thread
constexpr int kBlurPx = 2;

struct BlurTile {
  int ct, tx, tiles_x, tiles_y, tiles_c;
};

__global__ void __launch_bounds__(128, 4)
blur_act_nhwc_kernel(const __grid_constant__ CUtensorMap map_z, const __grid_constant__ BlurParams P,
                     const __grid_constant__ BlurTile T) {
  extern __shared__ uint8_t blur_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(blur_smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t bar;
  __shared__ float s_noise[kBlurTY][128];  // noise_w * gain * noise of the tile (TX <= 128)
  const int tid = threadIdx.x;
  int tile = blockIdx.x;
  const int tcx = tile % T.tiles_c; tile /= T.tiles_c;
  const int tx_i = tile % T.tiles_x; tile /= T.tiles_x;
  const int ty_i = tile % T.tiles_y;
  const int b = tile / T.tiles_y;
  const int ox_t = tx_i * T.tx, oy_t = ty_i * kBlurTY, c_t = tcx * T.ct;
  const int sw = T.tx + 3;                 // smem tile width in pixels
  const int pix_bytes = T.ct * 2;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(&bar, (uint32_t)((kBlurTY + 3) * sw * pix_bytes));
    tma_load_4d(smem, &map_z, &bar, c_t, ox_t - P.px0, oy_t - P.py0, b);
  }
  const bool lrelu = P.act == W2E_ACT_LRELU;
  const float gain = lrelu ? 1.41421356237309515f : 1.f;
  // while the TMA box is in flight: the tile's noise patch and this thread's per-channel constants
  if (P.noise) {
    const float nw = __ldg(P.noise_w) * gain;
    const float* noise = P.noise + (P.noise_per_sample ? (int64_t)b * P.H * P.W : 0);
    for (int e = tid; e < kBlurTY * T.tx; e += 128) {
      const int yy = e / T.tx, xx = e - yy * T.tx;
      const int oy = oy_t + yy, ox = ox_t + xx;
      s_noise[yy][xx] = (oy < P.H && ox < P.W) ? nw * __ldg(noise + (int64_t)oy * P.W + ox) : 0.f;
    }
  }
  const int cg = T.ct >> 3;
  const int g = tid % cg;
  const int xb = (tid / cg) % (T.tx / kBlurPx);
  const int rg = tid / (cg * (T.tx / kBlurPx));
  const int c0 = c_t + g * 8;
  const int ox0 = ox_t + xb * kBlurPx;
  const int oy0 = oy_t + rg * kBlurRows;
  // per-channel constants as fp32 pairs: all arithmetic below is packed fp32x2 (FFMA2/FMUL2/FADD2), which
  // halves the instruction count of this otherwise issue-bound kernel
  uint64_t bias2[4], nsc2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float b0 = (P.bias ? __ldg(P.bias + c0 + 2 * e) : 0.f) * gain, b1 = (P.bias ? __ldg(P.bias + c0 + 2 * e + 1) : 0.f) * gain;
    bias2[e] = pack2(b0, b1);
    const float n0 = P.next_scale ? __ldg(P.next_scale + (int64_t)b * P.C + c0 + 2 * e) : 1.f;
    const float n1 = P.next_scale ? __ldg(P.next_scale + (int64_t)b * P.C + c0 + 2 * e + 1) : 1.f;
    nsc2[e] = pack2(n0, n1);
  }
  const uint64_t fv0 = pack2(P.fv[0] * gain, P.fv[0] * gain), fv1 = pack2(P.fv[1] * gain, P.fv[1] * gain);
  const uint64_t fv2 = pack2(P.fv[2] * gain, P.fv[2] * gain), fv3 = pack2(P.fv[3] * gain, P.fv[3] * gain);
  const uint64_t fh0 = pack2(P.fh[0], P.fh[0]), fh1 = pack2(P.fh[1], P.fh[1]);
  const uint64_t fh2 = pack2(P.fh[2], P.fh[2]), fh3 = pack2(P.fh[3], P.fh[3]);
  const float slope = lrelu ? 0.2f : 1.f;
  const uint64_t slope2 = pack2(slope, slope);
  const bool has_noise = P.noise != nullptr;
  // output pointers of this thread's first row; advanced by one image row per output row
  const int64_t o_first = (((int64_t)b * P.H + oy0) * P.W + ox0) * P.C + c0;
  __nv_bfloat16* o_ptr = P.out ? P.out + o_first : nullptr;
  __nv_bfloat16* m_ptr = P.out_mod ? P.out_mod + o_first : nullptr;
  const int64_t row_stride = (int64_t)P.W * P.C;
  const int rows_ok = min(kBlurRows, P.H - oy0);          // may be <= 0 on a ragged bottom edge
  bool px_ok[kBlurPx];
#pragma unroll
  for (int px = 0; px < kBlurPx; ++px) px_ok[px] = ox0 + px < P.W;
  __syncthreads();  // barrier initialised, noise patch visible
  {
    uint32_t spin = 0;
    while (!mbar_try_wait(&bar, 0))
      if (++spin > (1u << 26)) __trap();  // a lost TMA completion becomes a launch failure, not a hang
  }
  const uint32_t col = smem_u32(smem) + (uint32_t)(((rg * kBlurRows) * sw + xb * kBlurPx) * pix_bytes + g * 16);
  const uint32_t row_bytes = (uint32_t)(sw * pix_bytes);

  uint64_t win[4][kBlurPx][4];   // ring of the last four horizontally filtered rows (compile-time slots)
#pragma unroll
  for (int t = 0; t < kBlurRows + 3; ++t) {
    const int u = t & 3;
    uint64_t f[kBlurPx + 3][4];
#pragma unroll
    for (int k = 0; k < kBlurPx + 3; ++k) {
      uint32_t w0, w1, w2, w3;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                   : "r"(col + (uint32_t)t * row_bytes + (uint32_t)(k * pix_bytes)));
      // bf16 pair -> fp32 pair: low element = w << 16, high element = w & 0xffff0000
      f[k][0] = pack2u(w0 << 16, w0 & 0xffff0000u);
      f[k][1] = pack2u(w1 << 16, w1 & 0xffff0000u);
      f[k][2] = pack2u(w2 << 16, w2 & 0xffff0000u);
      f[k][3] = pack2u(w3 << 16, w3 & 0xffff0000u);
    }
#pragma unroll
    for (int px = 0; px < kBlurPx; ++px)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        win[u][px][e] = fma2(fh3, f[px + 3][e], fma2(fh2, f[px + 2][e], fma2(fh1, f[px + 1][e], mul2(fh0, f[px][e]))));
    if (t >= 3) {
      const int r = t - 3;
      if (r < rows_ok) {
#pragma unroll
        for (int px = 0; px < kBlurPx; ++px) {
          if (px_ok[px]) {
            const float nz = has_noise ? s_noise[rg * kBlurRows + r][xb * kBlurPx + px] : 0.f;
            const uint64_t nz2 = pack2(nz, nz);
            uint32_t po[4], pm[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // rows t-3 .. t live in slots (u+1)&3, (u+2)&3, (u+3)&3, u
              const uint64_t a = fma2(fv3, win[u][px][e],
                                      fma2(fv2, win[(u + 3) & 3][px][e],
                                           fma2(fv1, win[(u + 2) & 3][px][e],
                                                fma2(fv0, win[(u + 1) & 3][px][e], add2(bias2[e], nz2)))));
              const uint64_t a_s = mul2(a, slope2);
              float x0, x1, y0, y1;
              unpack2(a, x0, x1);
              unpack2(a_s, y0, y1);
              x0 = fmaxf(x0, y0);
              x1 = fmaxf(x1, y1);
              if (o_ptr) po[e] = cvt_bf16x2(x0, x1);
              if (m_ptr) {
                float m0, m1;
                unpack2(mul2(pack2(x0, x1), nsc2[e]), m0, m1);
                pm[e] = cvt_bf16x2(m0, m1);
              }
            }
            if (o_ptr) *reinterpret_cast<uint4*>(o_ptr + px * P.C) = make_uint4(po[0], po[1], po[2], po[3]);
            if (m_ptr) *reinterpret_cast<uint4*>(m_ptr + px * P.C) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
          }
        }
      }
      if (o_ptr) o_ptr += row_stride;
      if (m_ptr) m_ptr += row_stride;
    }
  }
}

// ---- blur v2 (opt-in: W2E_BLUR_V2=1; NOT yet measured on hardware) -----------------------------------------
// ncu on the 1024^2 launch of blur_act_nhwc_kernel: issue-bound (69 % issue utilisation, DRAM 59 %), ~2 500
// instructions per thread of which ~40 % are integer / predicate overhead: shared-memory addresses built from
// the run-time tile shape, per-store edge predicates, null checks of the two outputs.  v2 keeps the tile, the
// thread mapping and the arithmetic order of v1 (so results are bit-identical) and makes the channel tile and
// the output set template parameters (shared-memory offsets become immediates) with a predicate-free body for
// tiles that lie fully inside the image.
template <int CT, bool HAS_OUT, bool HAS_MOD, bool INTERIOR>
__device__ __forceinline__ void blur_v2_body(const uint8_t* __restrict__ tile, const float* __restrict__ nz_row,
                                             bool has_noise, const uint64_t (&bias2)[4], const uint64_t (&nsc2)[4],
                                             const uint64_t (&fv)[4], const uint64_t (&fh)[4], uint64_t slope2,
                                             __nv_bfloat16* o_ptr, __nv_bfloat16* m_ptr, int64_t row_stride, int chan_stride,
                                             int rows_ok, bool px_ok0, bool px_ok1) {
  constexpr int TX = 1024 / CT, SW = TX + 3, PIX = CT * 2, ROW = SW * PIX;
  uint64_t win[4][kBlurPx][4];
#pragma unroll
  for (int t = 0; t < kBlurRows + 3; ++t) {
    const int u = t & 3;
    uint64_t f[kBlurPx + 3][4];
#pragma unroll
    for (int k = 0; k < kBlurPx + 3; ++k) {
      const uint4 w = *reinterpret_cast<const uint4*>(tile + t * ROW + k * PIX);
      f[k][0] = pack2u(w.x << 16, w.x & 0xffff0000u);
      f[k][1] = pack2u(w.y << 16, w.y & 0xffff0000u);
      f[k][2] = pack2u(w.z << 16, w.z & 0xffff0000u);
      f[k][3] = pack2u(w.w << 16, w.w & 0xffff0000u);
    }
#pragma unroll
    for (int px = 0; px < kBlurPx; ++px)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        win[u][px][e] = fma2(fh[3], f[px + 3][e], fma2(fh[2], f[px + 2][e], fma2(fh[1], f[px + 1][e], mul2(fh[0], f[px][e]))));
    if (t >= 3) {
      const int r = t - 3;
      if (INTERIOR || r < rows_ok) {
#pragma unroll
        for (int px = 0; px < kBlurPx; ++px) {
          if (INTERIOR || (px == 0 ? px_ok0 : px_ok1)) {
            const float nz = has_noise ? nz_row[r * 128 + px] : 0.f;
            const uint64_t nz2 = pack2(nz, nz);
            uint32_t po[4], pm[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint64_t a = fma2(fv[3], win[u][px][e],
                                      fma2(fv[2], win[(u + 3) & 3][px][e],
                                           fma2(fv[1], win[(u + 2) & 3][px][e],
                                                fma2(fv[0], win[(u + 1) & 3][px][e], add2(bias2[e], nz2)))));
              const uint64_t a_s = mul2(a, slope2);
              float x0, x1, y0, y1;
              unpack2(a, x0, x1);
              unpack2(a_s, y0, y1);
              x0 = fmaxf(x0, y0);
              x1 = fmaxf(x1, y1);
              if (HAS_OUT) po[e] = cvt_bf16x2(x0, x1);
              if (HAS_MOD) {
                float m0, m1;
                unpack2(mul2(pack2(x0, x1), nsc2[e]), m0, m1);
                pm[e] = cvt_bf16x2(m0, m1);
              }
            }
            if (HAS_OUT) *reinterpret_cast<uint4*>(o_ptr + r * row_stride + px * chan_stride) = make_uint4(po[0], po[1], po[2], po[3]);
            if (HAS_MOD) *reinterpret_cast<uint4*>(m_ptr + r * row_stride + px * chan_stride) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
          }
        }
      }
    }
  }
}

template <int CT, bool HAS_OUT, bool HAS_MOD>
__global__ void __launch_bounds__(128, 4)
blur_act_nhwc_v2_kernel(const __grid_constant__ CUtensorMap map_z, const __grid_constant__ BlurParams P,
                        const __grid_constant__ BlurTile T) {
  constexpr int TX = 1024 / CT, SW = TX + 3, PIX = CT * 2, CG = CT / 8, XB = TX / kBlurPx;
  static_assert(CG * XB * kBlurRG == 128, "128 threads = 2 row groups x column pairs x channel groups");
  extern __shared__ __align__(128) uint8_t blur_v2_smem[];
  __shared__ uint64_t bar;
  __shared__ float s_noise[kBlurTY][128];
  const int tid = threadIdx.x;
  int tile = blockIdx.x;
  const int tcx = tile % T.tiles_c; tile /= T.tiles_c;
  const int tx_i = tile % T.tiles_x; tile /= T.tiles_x;
  const int ty_i = tile % T.tiles_y;
  const int b = tile / T.tiles_y;
  const int ox_t = tx_i * TX, oy_t = ty_i * kBlurTY, c_t = tcx * CT;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(&bar, (uint32_t)((kBlurTY + 3) * SW * PIX));
    tma_load_4d(blur_v2_smem, &map_z, &bar, c_t, ox_t - P.px0, oy_t - P.py0, b);
  }
  const bool lrelu = P.act == W2E_ACT_LRELU;
  const float gain = lrelu ? 1.41421356237309515f : 1.f;
  if (P.noise) {
    const float nw = __ldg(P.noise_w) * gain;
    const float* noise = P.noise + (P.noise_per_sample ? (int64_t)b * P.H * P.W : 0);
    for (int e = tid; e < kBlurTY * TX; e += 128) {
      const int yy = e / TX, xx = e % TX;
      const int oy = oy_t + yy, ox = ox_t + xx;
      s_noise[yy][xx] = (oy < P.H && ox < P.W) ? nw * __ldg(noise + (int64_t)oy * P.W + ox) : 0.f;
    }
  }
  const int g = tid % CG;
  const int xb = (tid / CG) % XB;
  const int rg = tid / (CG * XB);
  const int c0 = c_t + g * 8;
  const int ox0 = ox_t + xb * kBlurPx;
  const int oy0 = oy_t + rg * kBlurRows;
  uint64_t bias2[4], nsc2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float b0 = (P.bias ? __ldg(P.bias + c0 + 2 * e) : 0.f) * gain, b1 = (P.bias ? __ldg(P.bias + c0 + 2 * e + 1) : 0.f) * gain;
    bias2[e] = pack2(b0, b1);
    const float n0 = HAS_MOD ? __ldg(P.next_scale + (int64_t)b * P.C + c0 + 2 * e) : 1.f;
    const float n1 = HAS_MOD ? __ldg(P.next_scale + (int64_t)b * P.C + c0 + 2 * e + 1) : 1.f;
    nsc2[e] = pack2(n0, n1);
  }
  uint64_t fv[4], fh[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    fv[i] = pack2(P.fv[i] * gain, P.fv[i] * gain);
    fh[i] = pack2(P.fh[i], P.fh[i]);
  }
  const float slope = lrelu ? 0.2f : 1.f;
  const uint64_t slope2 = pack2(slope, slope);
  const int64_t o_first = (((int64_t)b * P.H + oy0) * P.W + ox0) * P.C + c0;
  __nv_bfloat16* o_ptr = HAS_OUT ? P.out + o_first : nullptr;
  __nv_bfloat16* m_ptr = HAS_MOD ? P.out_mod + o_first : nullptr;
  const int64_t row_stride = (int64_t)P.W * P.C;
  const int rows_ok = min(kBlurRows, P.H - oy0);
  const bool px_ok0 = ox0 < P.W, px_ok1 = ox0 + 1 < P.W;
  const bool interior = (oy_t + kBlurTY <= P.H) && (ox_t + TX <= P.W);   // uniform over the CTA
  __syncthreads();
  {
    uint32_t spin = 0;
    while (!mbar_try_wait(&bar, 0))
      if (++spin > (1u << 26)) __trap();
  }
  const uint8_t* tile_ptr = blur_v2_smem + ((rg * kBlurRows) * SW + xb * kBlurPx) * PIX + g * 16;
  const float* nz_row = &s_noise[rg * kBlurRows][xb * kBlurPx];
  const bool has_noise = P.noise != nullptr;
  if (interior)
    blur_v2_body<CT, HAS_OUT, HAS_MOD, true>(tile_ptr, nz_row, has_noise, bias2, nsc2, fv, fh, slope2, o_ptr, m_ptr, row_stride,
                                             P.C, rows_ok, px_ok0, px_ok1);
  else
    blur_v2_body<CT, HAS_OUT, HAS_MOD, false>(tile_ptr, nz_row, has_noise, bias2, nsc2, fv, fh, slope2, o_ptr, m_ptr, row_stride,
                                              P.C, rows_ok, px_ok0, px_ok1);
}

template <int CT>
static int launch_blur_v2(const CUtensorMap& mz, const BlurParams& P, const BlurTile& T, int64_t blocks, int smem_bytes,
                          cudaStream_t stream) {
#define W2E_BLUR_V2_CASE(O, M)                                                                                        \
  {                                                                                                                    \
    static PerDeviceOnce configured;                                                                                    \
    if (!configured.done()) {                                                                                                 \
      W2E_CUDA_OK(cudaFuncSetAttribute(blur_act_nhwc_v2_kernel<CT, O, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       100 * 1024));                                                                   \
      configured.mark();                                                                                               \
    }                                                                                                                  \
    blur_act_nhwc_v2_kernel<CT, O, M><<<(unsigned)blocks, 128, smem_bytes, stream>>>(mz, P, T);                        \
  }
  if (P.out && P.out_mod) W2E_BLUR_V2_CASE(true, true)
  else if (P.out) W2E_BLUR_V2_CASE(true, false)
  else W2E_BLUR_V2_CASE(false, true)
#undef W2E_BLUR_V2_CASE
  W2E_LAUNCH_OK();
  return W2E_OK;
}

// ------------------------------------------------------------------------------ ToRGB (NHWC in, NCHW out)
// LP lanes share one pixel (8 channels per lane per step); a warp covers 32/LP pixels per step.
template <int LP>
__global__ void __launch_bounds__(256)
torgb_nhwc_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ style,
                  const float* __restrict__ bias, const float* __restrict__ skip, float4 kf, float* __restrict__ rgb,
                  int C, int H, int W, int pix_per_block) {
  extern __shared__ float wm[];  // [3][C] per-sample modulated weights
  const int b = blockIdx.y;
  for (int e = threadIdx.x; e < 3 * C; e += blockDim.x)
    wm[e] = __ldg(w + e) * __ldg(style + (int64_t)b * C + (e % C));
  __syncthreads();
  constexpr int PPW = 32 / LP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane % LP, grp = lane / LP;
  const int64_t HW = (int64_t)H * W;
  const int64_t p_begin = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p_end = min(HW, p_begin + pix_per_block);
  const __nv_bfloat16* xb = x + (int64_t)b * HW * C;
  const float kfa[4] = {kf.x, kf.y, kf.z, kf.w};
  constexpr int U = 4;  // independent pixels in flight per lane group
  for (int64_t p0 = p_begin + (int64_t)warp * PPW * U; p0 < p_end; p0 += (int64_t)nwarps * PPW * U) {
    float a0[U], a1[U], a2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a0[u] = a1[u] = a2[u] = 0.f;
    for (int c = sub * 8; c < C; c += LP * 8) {
      bf16x8 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t p = p0 + u * PPW + grp;
        if (p < p_end) v[u] = ld8(xb + p * C + c);
      }
      float w0[8], w1[8], w2[8];
#pragma unroll
      for (int e4 = 0; e4 < 2; ++e4) {
        const float4 t0 = *reinterpret_cast<const float4*>(wm + c + 4 * e4);
        const float4 t1 = *reinterpret_cast<const float4*>(wm + C + c + 4 * e4);
        const float4 t2 = *reinterpret_cast<const float4*>(wm + 2 * C + c + 4 * e4);
        w0[4 * e4] = t0.x; w0[4 * e4 + 1] = t0.y; w0[4 * e4 + 2] = t0.z; w0[4 * e4 + 3] = t0.w;
        w1[4 * e4] = t1.x; w1[4 * e4 + 1] = t1.y; w1[4 * e4 + 2] = t1.z; w1[4 * e4 + 3] = t1.w;
        w2[4 * e4] = t2.x; w2[4 * e4 + 1] = t2.y; w2[4 * e4 + 2] = t2.z; w2[4 * e4 + 3] = t2.w;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t p = p0 + u * PPW + grp;
        if (p < p_end) {
          float f[8];
          unpack8(v[u], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            a0[u] = fmaf(f[e], w0[e], a0[u]);
            a1[u] = fmaf(f[e], w1[e], a1[u]);
            a2[u] = fmaf(f[e], w2[e], a2[u]);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int o = LP >> 1; o > 0; o >>= 1) {
        a0[u] += __shfl_xor_sync(0xffffffffu, a0[u], o);
        a1[u] += __shfl_xor_sync(0xffffffffu, a1[u], o);
        a2[u] += __shfl_xor_sync(0xffffffffu, a2[u], o);
      }
      const int64_t p = p0 + u * PPW + grp;
      if (p < p_end && sub < 3) {
        float v = sub == 0 ? a0[u] : (sub == 1 ? a1[u] : a2[u]);
        if (bias) v += __ldg(bias + sub);
        if (skip) {
          // upfirdn2d(skip, k, up=2, pad=(2,1)) as a polyphase filter: 2 taps per axis
          const int oy = (int)(p / W), ox = (int)(p % W);
          const int h = H / 2, wd = W / 2;
          const int ya = (oy & 1) ? (oy - 1) / 2 : oy / 2 - 1, xa = (ox & 1) ? (ox - 1) / 2 : ox / 2 - 1;
          const float cy0 = kfa[(oy & 1)], cy1 = kfa[(oy & 1) + 2];
          const float cx0 = kfa[(ox & 1)], cx1 = kfa[(ox & 1) + 2];
          const float* sp = skip + ((int64_t)b * 3 + sub) * h * wd;
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
          v += acc;
        }
        rgb[((int64_t)b * 3 + sub) * HW + p] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------ blend (NHWC bf16)
// out = m*edited + (1-m)*orig ; out_mod = out * next_scale.  8 channels per thread.
__global__ void __launch_bounds__(256)
blend_nhwc_kernel(const __nv_bfloat16* __restrict__ edited, const __nv_bfloat16* __restrict__ orig,
                  const float* __restrict__ mask, const float* __restrict__ next_scale,
                  __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out_mod, int64_t total8, int C, int H,
                  int W, int mh, int mw) {
  const int cg = C >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    const int64_t pix = i / cg;
    const int xx = (int)(pix % W), yy = (int)((pix / W) % H);
    const int64_t b = pix / ((int64_t)W * H);
    const int my = (int)(((int64_t)yy * mh) / H), mx = (int)(((int64_t)xx * mw) / W);
    const float m = __ldg(mask + (b * mh + my) * mw + mx);
    float e[8], o[8], v[8];
    unpack8(ld8(edited + i * 8), e);
    unpack8(ld8(orig + i * 8), o);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __fadd_rn(__fmul_rn(m, e[k]), __fmul_rn(__fsub_rn(1.f, m), o[k]));
    if (out) st8(out + i * 8, pack8(v));
    if (out_mod) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= __ldg(next_scale + b * C + g * 8 + k);
      st8(out_mod + i * 8, pack8(v));
    }
  }
}

}  // namespace w2e

using namespace w2e;

extern "C" int w2e_nchw_to_nhwc_mod(const float* x, const float* style, void* y, int B, int Bx, int C, int64_t HW,
                                    int dtype, void* stream) {
  W2E_CHECK_ARG(dtype == W2E_BF16 || dtype == W2E_F32, "nchw_to_nhwc_mod: dtype");
  W2E_CHECK_ARG(x && y, "nchw_to_nhwc_mod: null pointer");
  W2E_CHECK_ARG(B >= 0 && C > 0 && HW > 0 && (Bx == 1 || Bx == B), "nchw_to_nhwc_mod: bad shape");
  W2E_CHECK_ARG(B <= 65535 && ceil_div(C, 32) <= 65535, "nchw_to_nhwc_mod: batch/channels above 65535");
  if (B == 0) return W2E_OK;
  dim3 grid((unsigned)ceil_div64(HW, 32), (unsigned)ceil_div(C, 32), (unsigned)B);
  if (dtype == W2E_F32) nchw_to_nhwc_mod_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(x, style, (float*)y, Bx, C, HW);
  else nchw_to_nhwc_mod_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(x, style, (__nv_bfloat16*)y, Bx, C, HW);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_nhwc_to_nchw_f32(const void* x, float* y, int B, int C, int64_t HW, int dtype, void* stream) {
  W2E_CHECK_ARG(x && y, "nhwc_to_nchw_f32: null pointer");
  W2E_CHECK_ARG(dtype == W2E_BF16 || dtype == W2E_F32, "nhwc_to_nchw_f32: dtype");
  W2E_CHECK_ARG(B >= 0 && C > 0 && HW > 0 && B <= 65535, "nhwc_to_nchw_f32: bad shape");
  if (B == 0) return W2E_OK;
  if (dtype == W2E_BF16 && (C % 64 == 0 || C == 32) && HW >= 64 && (((uintptr_t)x & 15) == 0) && (((uintptr_t)y & 7) == 0)) {
    const int ch = C % 64 == 0 ? 64 : 32;
    dim3 grid2((unsigned)ceil_div64(HW, 64), (unsigned)(C / ch), (unsigned)B);
    if (ch == 64) nhwc_bf16_to_nchw_f32_v2_kernel<64><<<grid2, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, y, C, HW);
    else nhwc_bf16_to_nchw_f32_v2_kernel<32><<<grid2, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, y, C, HW);
    W2E_LAUNCH_OK();
    return W2E_OK;
  }
  dim3 grid((unsigned)ceil_div64(HW, 32), (unsigned)ceil_div(C, 32), (unsigned)B);
  if (dtype == W2E_F32) nhwc_to_nchw_f32_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, y, C, HW);
  else nhwc_to_nchw_f32_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, y, C, HW);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_nchw_class_to_nhwc_mod(const float* x, const float* scale, void* y, int B, int C, int H, int W, int py,
                                          int px, int dtype, void* stream) {
  W2E_CHECK_ARG(x && y, "nchw_class_to_nhwc_mod: null pointer");
  W2E_CHECK_ARG(dtype == W2E_BF16 || dtype == W2E_F32, "nchw_class_to_nhwc_mod: dtype");
  W2E_CHECK_ARG(B >= 0 && C > 0 && H > 0 && W > 0 && (py == 0 || py == 1) && (px == 0 || px == 1) && B <= 65535,
                "nchw_class_to_nhwc_mod: bad shape");
  const int hc = (H - py + 1) / 2, wc = (W - px + 1) / 2;
  if (B == 0 || hc == 0 || wc == 0) return W2E_OK;
  dim3 grid((unsigned)ceil_div64((int64_t)hc * wc, 32), (unsigned)ceil_div(C, 32), (unsigned)B);
  if (dtype == W2E_F32)
    nchw_class_to_nhwc_mod_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(x, scale, (float*)y, C, H, W, hc, wc, py, px);
  else
    nchw_class_to_nhwc_mod_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(x, scale, (__nv_bfloat16*)y, C, H, W, hc, wc, py, px);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_nhwc_sum4_to_nchw_f32(const void* y00, const void* y01, const void* y10, const void* y11, float* out,
                                         int B, int C, int h, int w, int dtype, void* stream) {
  W2E_CHECK_ARG(y00 && y01 && y10 && y11 && out, "nhwc_sum4_to_nchw_f32: null pointer");
  W2E_CHECK_ARG(dtype == W2E_BF16 || dtype == W2E_F32, "nhwc_sum4_to_nchw_f32: dtype");
  W2E_CHECK_ARG(B >= 0 && C > 0 && h > 0 && w > 0 && B <= 65535, "nhwc_sum4_to_nchw_f32: bad shape");
  if (B == 0) return W2E_OK;
  dim3 grid((unsigned)ceil_div64((int64_t)h * w, 32), (unsigned)ceil_div(C, 32), (unsigned)B);
  if (dtype == W2E_F32)
    nhwc_sum4_to_nchw_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)y00, (const float*)y01, (const float*)y10,
                                                                          (const float*)y11, out, C, h, w);
  else
    nhwc_sum4_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)y00, (const __nv_bfloat16*)y01, (const __nv_bfloat16*)y10, (const __nv_bfloat16*)y11, out, C, h, w);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_blur_act_nhwc(const void* z, const float* host_taps, const float* bias, const float* noise,
                                 const float* noise_w, int noise_batch, const float* next_scale, void* out,
                                 void* out_mod, int B, int C, int in_h, int in_w, int py0, int px0, int out_h,
                                 int out_w, int act, int variant, void* stream) {
  W2E_CHECK_ARG(z && host_taps && (out || out_mod), "blur_act_nhwc: null pointer");
  W2E_CHECK_ARG(B >= 0 && C > 0 && C % 8 == 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0,
                "blur_act_nhwc: bad shape (C must be a multiple of 8)");
  W2E_CHECK_ARG(out_mod == nullptr || next_scale != nullptr, "blur_act_nhwc: out_mod needs next_scale");
  W2E_CHECK_ARG(noise == nullptr || (noise_w != nullptr && (noise_batch == 1 || noise_batch == B)),
                "blur_act_nhwc: noise");
  // separable factorisation of the 4x4 kernel (rank-1 check as in upfirdn2d.cu)
  int bi = 0, bj = 0;
  float best = 0.f;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (fabsf(host_taps[i * 4 + j]) > best) best = fabsf(host_taps[i * 4 + j]), bi = i, bj = j;
  W2E_CHECK_ARG(best > 0.f, "blur_act_nhwc: zero kernel");
  float kv[4], kh[4];
  for (int i = 0; i < 4; ++i) kv[i] = host_taps[i * 4 + bj];
  for (int j = 0; j < 4; ++j) kh[j] = host_taps[bi * 4 + j] / host_taps[bi * 4 + bj];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      W2E_CHECK_ARG(fabsf(host_taps[i * 4 + j] - kv[i] * kh[j]) <= 1e-6f * best,
                    "blur_act_nhwc: the 4x4 kernel is not separable");
  if (B == 0) return W2E_OK;
  BlurParams P;
  P.z = (const __nv_bfloat16*)z; P.bias = bias; P.noise = noise; P.noise_w = noise_w; P.next_scale = next_scale;
  P.out = (__nv_bfloat16*)out; P.out_mod = (__nv_bfloat16*)out_mod;
  P.noise_per_sample = (noise && noise_batch != 1) ? 1 : 0;
  P.B = B; P.C = C; P.IH = in_h; P.IW = in_w; P.H = out_h; P.W = out_w; P.py0 = py0; P.px0 = px0; P.act = act;
  for (int i = 0; i < 4; ++i) { P.fv[i] = kv[3 - i]; P.fh[i] = kh[3 - i]; }
  BlurTile T;
  T.ct = C >= 128 ? 128 : C;
  W2E_CHECK_ARG((T.ct & (T.ct - 1)) == 0 && T.ct >= 8 && C % T.ct == 0,
                "blur_act_nhwc: C must be 8, 16, 32, 64 or a multiple of 128 (got %d)", C);
  T.tx = 1024 / T.ct;
  T.tiles_x = ceil_div(out_w, T.tx); T.tiles_y = ceil_div(out_h, kBlurTY); T.tiles_c = C / T.ct;
  const int64_t blocks = (int64_t)T.tiles_x * T.tiles_y * T.tiles_c * B;
  W2E_CHECK_ARG(blocks < (1ll << 31), "blur_act_nhwc: too many tiles");
  W2E_CHECK_ARG(((uintptr_t)z & 15) == 0, "blur_act_nhwc: input must be 16-byte aligned");
  CUtensorMap mz;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)in_w, (uint64_t)in_h, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)in_w * C * 2, (uint64_t)in_h * in_w * C * 2};
    const uint32_t box[4] = {(uint32_t)T.ct, (uint32_t)(T.tx + 3), (uint32_t)(kBlurTY + 3), 1u};
    int rc = make_bf16_map(&mz, z, 4, dims, strides, box, 0 /* no swizzle */);
    if (rc) return rc;
  }
  const int smem_bytes = (kBlurTY + 3) * (T.tx + 3) * T.ct * 2 + 128;
  const bool use_v2 = variant == 1;
  if (use_v2 && T.ct >= 32) {   // the two kernels are bit-identical (tests/test_variants_gpu.py); `variant` picks one per call
    if (T.ct == 32) return launch_blur_v2<32>(mz, P, T, blocks, smem_bytes, (cudaStream_t)stream);
    if (T.ct == 64) return launch_blur_v2<64>(mz, P, T, blocks, smem_bytes, (cudaStream_t)stream);
    return launch_blur_v2<128>(mz, P, T, blocks, smem_bytes, (cudaStream_t)stream);
  }
  static PerDeviceOnce configured;
  if (!configured.done()) {
    W2E_CUDA_OK(cudaFuncSetAttribute(blur_act_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured.mark();
  }
  blur_act_nhwc_kernel<<<(unsigned)blocks, 128, smem_bytes, (cudaStream_t)stream>>>(mz, P, T);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_torgb_nhwc(const void* x, const float* w, const float* style, const float* bias,
                              const float* skip, const float* host_taps1d, float* rgb, int B, int C, int H, int W,
                              void* stream) {
  W2E_CHECK_ARG(x && w && style && rgb, "torgb_nhwc: null pointer");
  W2E_CHECK_ARG(B >= 0 && C >= 32 && C % 8 == 0 && H > 0 && W > 0 && B <= 65535, "torgb_nhwc: bad shape");
  W2E_CHECK_ARG(skip == nullptr || (host_taps1d != nullptr && H % 2 == 0 && W % 2 == 0), "torgb_nhwc: skip needs taps and even H, W");
  if (B == 0) return W2E_OK;
  float4 kf = make_float4(0, 0, 0, 0);
  if (skip) kf = make_float4(host_taps1d[3], host_taps1d[2], host_taps1d[1], host_taps1d[0]);
  const int64_t HW = (int64_t)H * W;
  // enough blocks to fill the machine, at least 256 pixels per block
  int64_t want_blocks = (int64_t)sm_count() * 8 / (B > 0 ? B : 1) + 1;
  int64_t ppb = ceil_div64(HW, want_blocks);
  if (ppb < 256) ppb = 256;
  ppb = ceil_div64(ppb, 32) * 32;
  dim3 grid((unsigned)ceil_div64(HW, ppb), (unsigned)B);
  const size_t smem = (size_t)3 * C * sizeof(float);
  cudaStream_t s = (cudaStream_t)stream;
  const __nv_bfloat16* xp = (const __nv_bfloat16*)x;
#define W2E_RGB_CASE(lp) torgb_nhwc_kernel<lp><<<grid, 256, smem, s>>>(xp, w, style, bias, skip, kf, rgb, C, H, W, (int)ppb)
  if (C >= 256) W2E_RGB_CASE(32);
  else if (C >= 128) W2E_RGB_CASE(16);
  else if (C >= 64) W2E_RGB_CASE(8);
  else W2E_RGB_CASE(4);
#undef W2E_RGB_CASE
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_blend_nhwc(const void* edited, const void* orig, const float* mask, const float* next_scale,
                              void* out, void* out_mod, int B, int C, int H, int W, int mh, int mw, void* stream) {
  W2E_CHECK_ARG(edited && orig && mask && (out || out_mod), "blend_nhwc: null pointer");
  W2E_CHECK_ARG(B >= 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0 && mh > 0 && mw > 0, "blend_nhwc: bad shape");
  W2E_CHECK_ARG(out_mod == nullptr || next_scale != nullptr, "blend_nhwc: out_mod needs next_scale");
  const int64_t total8 = (int64_t)B * H * W * (C / 8);
  if (total8 == 0) return W2E_OK;
  const int64_t blocks = ceil_div64(total8, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  blend_nhwc_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)edited, (const __nv_bfloat16*)orig, mask, next_scale, (__nv_bfloat16*)out,
      (__nv_bfloat16*)out_mod, total8, C, H, W, mh, mw);
  W2E_LAUNCH_OK();
  return W2E_OK;
}
