// Tensor-core modulated convolution for sm_100a: TMA-fed implicit GEMM, tcgen05.mma with the
// accumulator in TMEM, fused demodulation / noise / bias / leaky-ReLU epilogue.
// Replaces the grouped F.conv2d / F.conv_transpose2d of models/stylegan2/model.py:249-274 in
// bf16 mode.  SURVEY.md section 2.2 kernels K1a (plain 3x3) and K1b (transposed x2, polyphase).
//
// GEMM view (one launch = one "class" of output pixels):
//   D[m, n] = sum_{tap, c} A_tap[m, c] * W[slot(tap)][n, c]
//   m = output grid point (b, j, i)   -- 128 per CTA: a box of NB images x TH rows x TW cols
//   n = output channel                -- BN per CTA (<= 256 TMEM columns)
//   A_tap[m, c] = xs[b, j + dy(tap), i + dx(tap), c]   (NHWC bf16, ALREADY multiplied by the
//                 style: the producer of xs folds the modulation into its epilogue, so the
//                 whole batch shares ONE weight tensor)
// A tiles are fetched by 4-D TMA boxes {BK, TW, TH, NB} at shifted coordinates; TMA's
// out-of-bounds zero fill IS the convolution padding.  B tiles are 3-D boxes {BK, BN, 1} of the
// weight tensor [slots][Cout][Cin].  Both land in 128B(64B)-swizzled K-major shared memory, the
// canonical UMMA operand layout, so no thread ever touches the operands.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one
// elected lane), warps 2..5 = epilogue (TMEM -> registers -> global, one accumulator row = one
// output pixel per thread).  smem ring of kStages {A,B} stages with full/empty mbarriers;
// tcgen05.commit releases stages and publishes the accumulator.
//
// Algorithmic FLOPs per launch: 2 * ntaps * Cin * Cout * B * grid_h * grid_w.
#include "tc_ptx.cuh"

namespace w2e {

// ----------------------------------------------------------------------------------------- kernel
constexpr int kTcMaxTaps = 9;
constexpr int kTcThreads = 192;

struct TcParams {
  const float* out_scale;   // [B,Cout] demodulation or null
  const float* bias;        // [Cout] or null
  const float* noise;       // [noise_batch, OH*OW] or null
  const float* noise_w;     // device scalar
  const float* next_scale;  // [B,Cout] style of the consumer layer (for out_mod) or null
  __nv_bfloat16* out;       // [B,OH,OW,Cout] unscaled result or null
  __nv_bfloat16* out_mod;   // [B,OH,OW,Cout] result * next_scale or null
  int* error_flag;          // set to 1 if a pipeline wait timed out
  int noise_per_sample;
  int B, Cin, Cout, OH, OW;
  int grid_h, grid_w, out_stride, py, px;
  int nb, th, tw, tiles_x, tiles_y;
  int act;
  int ntaps;
  int dy[kTcMaxTaps], dx[kTcMaxTaps], slot[kTcMaxTaps];
};

template <int BN, int BK, int STAGES>
struct TcSmem {
  static constexpr int kABytes = 128 * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // barriers + alignment slack
};

template <int BN, int BK, int STAGES>
__global__ void __launch_bounds__(kTcThreads, 1)
modconv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ TcParams P) {
  using L = TcSmem<BN, BK, STAGES>;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates
  const int tile = blockIdx.x;
  const int tx = tile % P.tiles_x, ty = (tile / P.tiles_x) % P.tiles_y, tn = tile / (P.tiles_x * P.tiles_y);
  const int i0 = tx * P.tw, j0 = ty * P.th, n0 = tn * P.nb;
  const int co0 = blockIdx.y * BN;
  const int kchunks = P.Cin / BK;
  const int kiters = P.ntaps * kchunks;

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(accum_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer
      for (int it = 0; it < kiters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        if (!mbar_wait(&empty_bar[s], ph ^ 1u, abort_flag)) break;
        const int tap = it / kchunks, kc = it % kchunks;
        uint8_t* a_dst = smem + s * L::kStageBytes;
        uint8_t* b_dst = a_dst + L::kABytes;
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)L::kStageBytes);
        tma_load_4d(a_dst, &map_a, &full_bar[s], kc * BK, i0 + P.dx[tap], j0 + P.dy[tap], n0);
        tma_load_3d(b_dst, &map_b, &full_bar[s], kc * BK, co0, P.slot[tap]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer
      // instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major both, N>>3 @17, M>>4 @24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
      bool ok = true;
      for (int it = 0; it < kiters && ok; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        ok = mbar_wait(&full_bar[s], ph, abort_flag);
        if (!ok) break;
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
        const uint32_t b_addr = a_addr + L::kABytes;
        const uint64_t adesc = make_kmajor_desc(a_addr, BK * 2);
        const uint64_t bdesc = make_kmajor_desc(b_addr, BK * 2);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in 16-byte units
          umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem stage when these MMAs retire
      }
      umma_commit(accum_bar);  // accumulator complete
    }
  } else {
    // ---------------- epilogue: warps 2..5, TMEM lane quarter = warp % 4
    const int q = warp & 3;
    const int r = q * 32 + lane;  // accumulator row == pixel inside the tile
    const bool ready = mbar_wait(accum_bar, 0, abort_flag);
    tc_fence_after();
    const int n_in = r / (P.th * P.tw), yy = (r / P.tw) % P.th, xx = r % P.tw;
    const int b = n0 + n_in, j = j0 + yy, i = i0 + xx;
    const bool valid = ready && b < P.B && j < P.grid_h && i < P.grid_w;
    const int oy = j * P.out_stride + P.py, ox = i * P.out_stride + P.px;
    const int64_t pix = valid ? ((int64_t)b * P.OH + oy) * P.OW + ox : 0;
    float nz = 0.f;
    if (valid && P.noise)
      nz = __ldg(P.noise_w) * __ldg(P.noise + (P.noise_per_sample ? (int64_t)b * P.OH * P.OW : 0) + (int64_t)oy * P.OW + ox);
    const int bb = valid ? b : 0;
    const float* dsc = P.out_scale ? P.out_scale + (int64_t)bb * P.Cout + co0 : nullptr;
    const float* nsc = P.next_scale ? P.next_scale + (int64_t)bb * P.Cout + co0 : nullptr;
    const float* bia = P.bias ? P.bias + co0 : nullptr;
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      tmem_ld_wait();
      if (valid) {
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float val = __uint_as_float(v[e]);
          if (dsc) val *= __ldg(dsc + c + e);
          if (P.act == W2E_ACT_LRELU) {
            val += nz;
            if (bia) val += __ldg(bia + c + e);
            val = lrelu_gain(val, 0.2f, 1.41421356237309515f);
          } else if (bia) {
            val += __ldg(bia + c + e);
          }
          f[e] = val;
        }
        if (P.out) {
          uint4 pk[2];
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
          for (int e = 0; e < 8; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
          uint4* dst = reinterpret_cast<uint4*>(P.out + pix * P.Cout + co0 + c);
          dst[0] = pk[0];
          dst[1] = pk[1];
        }
        if (P.out_mod) {
          uint4 pk[2];
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            h[e] = __floats2bfloat162_rn(f[2 * e] * __ldg(nsc + c + 2 * e), f[2 * e + 1] * __ldg(nsc + c + 2 * e + 1));
          uint4* dst = reinterpret_cast<uint4*>(P.out_mod + pix * P.Cout + co0 + c);
          dst[0] = pk[0];
          dst[1] = pk[1];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && *abort_flag && P.error_flag) *P.error_flag = 1;
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ----------------------------------------------------------------------------------------- host
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_bf16_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int row_bytes) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) return set_error(W2E_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                  reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_bytes),
                  reinterpret_cast<const cuuint32_t*>(box), estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(W2E_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return W2E_OK;
}

int make_f32_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) return set_error(W2E_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base),
                  reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_bytes),
                  reinterpret_cast<const cuuint32_t*>(box), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(W2E_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed with CUresult %d", (int)r);
  return W2E_OK;
}

template <int BN, int BK, int STAGES>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& P, dim3 grid, cudaStream_t s) {
  using L = TcSmem<BN, BK, STAGES>;
  static bool configured = false;
  if (!configured) {
    W2E_CUDA_OK(cudaFuncSetAttribute(modconv_tc_kernel<BN, BK, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     L::kTotal));
    configured = true;
  }
  modconv_tc_kernel<BN, BK, STAGES><<<grid, kTcThreads, L::kTotal, s>>>(ma, mb, P);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

}  // namespace w2e

using namespace w2e;

extern "C" int w2e_modconv_tc_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return (major == 10 && tensor_map_encoder() != nullptr) ? 1 : 0;
}

extern "C" int w2e_modconv_tc(const void* xs, const void* w, const float* out_scale, const float* bias,
                              const float* noise, const float* noise_w, int noise_batch, const float* next_scale,
                              void* out, void* out_mod, int* error_flag, int B, int Cin, int Cout, int in_h, int in_w,
                              int out_h, int out_w, int grid_h, int grid_w, int out_stride, int py, int px,
                              const int* host_taps, int ntaps, int nslots, int act, void* stream) {
  W2E_CHECK_ARG(xs && w && (out || out_mod), "modconv_tc: null pointer");
  W2E_CHECK_ARG(out_mod == nullptr || next_scale != nullptr, "modconv_tc: out_mod needs next_scale");
  W2E_CHECK_ARG(B > 0 && in_h > 0 && in_w > 0 && grid_h > 0 && grid_w > 0, "modconv_tc: bad shape");
  W2E_CHECK_ARG(ntaps > 0 && ntaps <= kTcMaxTaps && nslots > 0, "modconv_tc: bad tap count %d", ntaps);
  W2E_CHECK_ARG(Cin % 32 == 0 && Cout % 16 == 0, "modconv_tc: needs Cin %% 32 == 0 and Cout %% 16 == 0 (got %d, %d)", Cin, Cout);
  W2E_CHECK_ARG((grid_h - 1) * out_stride + py < out_h && (grid_w - 1) * out_stride + px < out_w,
                "modconv_tc: output grid exceeds the output tensor");
  W2E_CHECK_ARG(noise == nullptr || (noise_w != nullptr && (noise_batch == 1 || noise_batch == B)), "modconv_tc: noise");
  W2E_CHECK_ARG(((uintptr_t)xs & 15) == 0 && ((uintptr_t)w & 15) == 0, "modconv_tc: operands must be 16-byte aligned");
  if (!w2e_modconv_tc_supported()) return set_error(W2E_ERR_UNSUPPORTED, "modconv_tc: device is not sm_100");

  TcParams P;
  P.out_scale = out_scale; P.bias = bias; P.noise = noise; P.noise_w = noise_w; P.next_scale = next_scale;
  P.out = (__nv_bfloat16*)out; P.out_mod = (__nv_bfloat16*)out_mod; P.error_flag = error_flag;
  P.noise_per_sample = (noise && noise_batch != 1) ? 1 : 0;
  P.B = B; P.Cin = Cin; P.Cout = Cout; P.OH = out_h; P.OW = out_w;
  P.grid_h = grid_h; P.grid_w = grid_w; P.out_stride = out_stride; P.py = py; P.px = px; P.act = act; P.ntaps = ntaps;
  for (int t = 0; t < kTcMaxTaps; ++t) {
    P.dy[t] = t < ntaps ? host_taps[3 * t] : 0;
    P.dx[t] = t < ntaps ? host_taps[3 * t + 1] : 0;
    P.slot[t] = t < ntaps ? host_taps[3 * t + 2] : 0;
    W2E_CHECK_ARG(P.slot[t] >= 0 && P.slot[t] < nslots, "modconv_tc: weight slot out of range");
  }
  // 128 grid points per CTA: NB images x TH rows x TW columns
  if (grid_w >= 16) { P.nb = 1; P.th = 8; P.tw = 16; }
  else if (grid_w >= 8) { P.nb = 2; P.th = 8; P.tw = 8; }
  else { P.nb = 8; P.th = 4; P.tw = 4; }
  P.tiles_x = ceil_div(grid_w, P.tw); P.tiles_y = ceil_div(grid_h, P.th);
  const int tiles_n = ceil_div(B, P.nb);

  const int BK = (Cin % 64 == 0) ? 64 : 32;
  int BN = 16;
  for (int cand : {256, 128, 64, 32, 16})
    if (Cout % cand == 0) { BN = cand; break; }

  CUtensorMap ma, mb;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)in_w, (uint64_t)in_h, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)in_w * Cin * 2, (uint64_t)in_h * in_w * Cin * 2};
    const uint32_t box[4] = {(uint32_t)BK, (uint32_t)P.tw, (uint32_t)P.th, (uint32_t)P.nb};
    int rc = make_bf16_map(&ma, xs, 4, dims, strides, box, BK * 2);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)nslots};
    const uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
    const uint32_t box[3] = {(uint32_t)BK, (uint32_t)BN, 1u};
    int rc = make_bf16_map(&mb, w, 3, dims, strides, box, BK * 2);
    if (rc) return rc;
  }
  dim3 grid((unsigned)(P.tiles_x * P.tiles_y * tiles_n), (unsigned)(Cout / BN));
  cudaStream_t s = (cudaStream_t)stream;
#define W2E_TC_CASE(bn, bk, st) \
  if (BN == bn && BK == bk) return launch_tc<bn, bk, st>(ma, mb, P, grid, s);
  W2E_TC_CASE(256, 64, 4)
  W2E_TC_CASE(128, 64, 6)
  W2E_TC_CASE(64, 64, 8)
  W2E_TC_CASE(32, 64, 8)
  W2E_TC_CASE(16, 64, 8)
  W2E_TC_CASE(256, 32, 6)
  W2E_TC_CASE(128, 32, 8)
  W2E_TC_CASE(64, 32, 8)
  W2E_TC_CASE(32, 32, 8)
  W2E_TC_CASE(16, 32, 8)
#undef W2E_TC_CASE
  return set_error(W2E_ERR_UNSUPPORTED, "modconv_tc: no kernel for BN=%d BK=%d", BN, BK);
}
