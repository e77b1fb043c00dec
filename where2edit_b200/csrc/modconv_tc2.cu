// Persistent, warp-specialised tensor-core modulated convolution for sm_100a (v2 of modconv_tc.cu).
// Replaces the grouped F.conv2d / F.conv_transpose2d of models/stylegan2/model.py:249-274 in bf16
// mode.  SURVEY.md section 2.2 kernels K1a (plain 3x3) and K1b (transposed x2).
//
// What changed against v1 (one 128-pixel tile per CTA, 9 shifted TMA fetches of the input):
//  * HALOED INPUT TILE.  One TMA box {BK channels, P pixels, TH+2 rows} of the (already
//    style-modulated) NHWC input is loaded ONCE per K-chunk; all 9 filter taps read it in place
//    through shifted UMMA shared-memory descriptors: an output tile is 8 pixels wide, so one
//    8-row core-matrix group = one image-row segment, the group stride (SBO) is the pitch of the
//    haloed tile and a tap (dy,dx) is a start-address offset of (dy*P+dx) rows.  L2->SM traffic
//    for the activation drops from 9x to ~1.4x.
//  * PERSISTENT CTAs with a static tile schedule, an smem ring for the weight blocks (or the
//    whole weight resident in smem when it is small), and double-buffered TMEM accumulators, so
//    TMA, MMA and the epilogue of consecutive tiles overlap.
//  * Up to 256 pixels x 256 channels per tile (two M=128 accumulators share every weight block).
//  * The transposed (x2) convolution computes its four output-parity classes in ONE pass: the 9
//    taps accumulate into 4 TMEM accumulators (class = parity of (ky,kx)) from the same input tile.
//
// Warp roles (320 threads): warps 0..7 = epilogue, warp 8 lane 0 = TMA producer, warp 9 = MMA issuer
// (also owns the TMEM allocation); epilogue (TMEM -> registers -> global; one accumulator
// row = one output pixel per thread, two warps per TMEM lane quarter splitting the columns):
// * demodulation, + noise, + bias, leaky-ReLU * sqrt(2), and the next layer's style (out_mod) -- so
// the whole batch shares one weight tensor.  Per-channel constants are staged in shared memory
// once per tile.
//
// TS epilogue (template flag TS; tiles with <= 64-channel staging units, i.e. the >= 256^2 layers
// whose time is the epilogue, not the MMA): the bf16 result tile is staged in 128B/64B-swizzled
// shared memory (conflict-free 128-bit writes, one accumulator row = one pixel per thread) and
// leaves through ONE TMA store per 128-pixel x 64-channel unit (the four output-parity classes of
// the transposed conv through four strided tensor maps); the noise patch and the 2x2-tap patch of
// the ToRGB skip image arrive through TMA boxes issued by the producer warp one tile ahead
// (out-of-bounds zero fill = the image borders), so the epilogue threads do no global-memory
// address arithmetic at all; per-element math is packed fp32x2 (FFMA2/FMUL2/FADD2).
//
// Algorithmic FLOPs per launch: 2 * taps * Cin * Cout * B * (pixels of the class grids).
#include <type_traits>

#include "tc_ptx.cuh"

namespace w2e {

constexpr int kT2Threads = 320;    // TMA warp, MMA warp, 8 epilogue warps
constexpr int kT2ThreadsTS = 640;  // TS flavour: 2 x 8 epilogue warps (one group per accumulator buffer) + 4 service warps
constexpr int kT2ThreadsTS2 = 640; // (see the warp-role table in the kernel)
constexpr int kT2EpiThreads = 256;
constexpr int kT2MaxA = 8, kT2MaxB = 16, kT2MaxAcc = 4;
constexpr int kEStages = 4;   // epilogue-input ring (TS flavour): deep enough not to throttle the A-tile prefetch
constexpr int kTileW = 8, kSubTileH = 16;  // one M=128 sub-tile = 16 rows x 8 pixels

struct Tc2Params {
  const float* out_scale;   // [B,Cout] demodulation or null
  const float* bias;        // [Cout] or null
  const float* noise;       // [noise_batch, OH*OW] or null
  const float* noise_w;     // device scalar
  const float* next_scale;  // [B,Cout] or null
  __nv_bfloat16* out;       // [B,OH,OW,Cout] or null
  __nv_bfloat16* out_mod;   // [B,OH,OW,Cout] or null
  int* error_flag;
  // fused up-convolution + Blur (w2e_modconv_tc2_upblur): out/out_mod/noise/bias/next_scale then describe the
  // FINAL [B, OH2, OW2, Cout] activation; fbv/fbh = flipped vertical/horizontal taps of the separable 4x4 FIR
  int fb, OH2, OW2;
  float fbv[4], fbh[4];
  int tile_dy, tile_dx, tile_o;   // tile (ty, tx) starts at input position (ty*tile_dy + tile_o, tx*tile_dx + tile_o)
  int cluster;              // log2 of the cluster size (0 = no clusters): multicast weight blocks
  int pair;                 // x-pair mode of the 32-channel RGB-only layer (see run_tc2): accumulator row = TWO adjacent pixels
  int OW_real;              // pair mode: width of the image in pixels (OW counts pixel pairs)
  int skip_box_bytes;       // bytes of one sub-tile's skip patch in the epilogue-input stage
  int dgk;                  // fused dgrad of the transposed convolution, K-LOOP form (weights in the ring): K chunks per parity class
  long long dg_masks;       // (0 = off); the K loop walks class 0..3 x chunks, dg_masks = the four 9-bit tap masks (class c at bit 9c)
  int dg4, cls16;           // fused dgrad of the transposed convolution (w2e_modconv_tc2_dgrad_up): an A stage holds the FOUR
                            // parity-class tiles of the upstream gradient, cls16 = bytes of one class tile >> 4
  int percls;               // transposed conv, one accumulator set (4 classes x MT x bn = 512 TMEM columns): the classes
                            // are handed over one by one (see "per-class hand-over" in the kernel)
  int reduce_add;           // TS epilogue: `out` boxes are ADDED to global memory (TMA reduce) instead of stored
  int tap_mask;             // plain conv: bit t set = filter tap t is used (0x1ff = all); class convolutions of the
                            // transposed convolution's dgrad use 4 / 2 / 2 / 1 of the 9 taps
  int bgroup;               // weight-ring kernels, TS flavour: filter taps per weight REQUEST (1, or 3 = one tap row per TMA box)
  int flags;                // A/B switches (w2e_modconv_tc2_flags): 1 = no edge-tile tap masking, 2 = one MMA issuer
  long long* dbg;           // optional timeline of CTA 0 (tools/tc2_timeline.py): [tile][8] clock64 stamps
  // fused ToRGB (models/stylegan2/model.py:353-362), RGB variants only
  const float* rgb_w;       // [3,Cout] 1x1 weight * 1/sqrt(Cout)
  const float* rgb_style;   // [B,Cout]
  const float* rgb_bias;    // [3] or null
  const float* rgb_skip;    // [B,3,OH/2,OW/2] fp32 NCHW or null
  void* rgb;                // [B,3,OH,OW] NCHW, fp32, (rgb_bf16 == 1) bf16 or (rgb_bf16 == 2) uint8: the image in the dtype
  int rgb_bf16;             // the caller wants; uint8 = the quantisation of torchvision.utils.save_image(normalize=True,
                            // range=(-1, 1)) that the reference applies to its results (run_attention.py:1470, 1535)
  float kf[4];              // flipped 1-D taps of the skip upsample filter
  int noise_per_sample;
  int B, Cin, Cout, OH, OW;
  int grid_h, grid_w, out_stride;
  int tiles_x, tiles_y, tiles_n, ntiles;
  int bn, bk, mt, ng, wres, nbuf;
  int pitch, box_rows, a_stages, b_stages;
  int a_stage_bytes, a_box_bytes, b_block_bytes;
  int tmem_cols;
  int act;
  int ntaps;
  // TS epilogue: staging units of 128 pixels x ts_unit_ch channels, ts_slots per epilogue half
  int ts_unit_ch, ts_unit_bytes, ts_slots, ts_off;
  int e_off, e_stage_bytes, e_noise_bytes, e_bytes, e_info_off, use_e;   // epilogue-input stages (tile info, noise, skip)
  int bars_off;
};

// tensor maps of the TS epilogue
struct alignas(64) Tc2Maps {
  CUtensorMap noise;   // fp32 {OW, OH, noise_batch}, box {8, 16*MT, 1}
  CUtensorMap skip;    // fp32 {OW/2, OH/2, B*3}, box {12, 10, 3}
  CUtensorMap bh;      // cluster mode: weight map with a half-block box {BK, bn/2, 1}
  CUtensorMap b3;      // grouped weight requests: box {BK, bn, 3} = the three taps of one kernel row
  CUtensorMap a4[3];   // fused dgrad of the transposed convolution: parity classes (0,1), (1,0), (1,1) of the gradient (map_a = (0,0))
  CUtensorMap st[4];   // bf16 stores, box {unit_ch, 8, 16, 1}: plain [0] = out, [1] = out_mod; transposed [g] = class g of out
};

struct Tc2Bars {
  uint64_t a_full[kT2MaxA], a_empty[kT2MaxA];
  uint64_t b_full[kT2MaxB], b_empty[kT2MaxB];
  uint64_t acc_full[kT2MaxAcc], acc_empty[kT2MaxAcc];
  uint64_t w_full;
  uint64_t e_full[kEStages], e_empty[kEStages];
  uint64_t mma_turn[2];   // issue token of the two MMA-issuing warps (they must alternate, not interleave)
  uint32_t tmem_slot;
  int abort_flag;
  // per-tile epilogue constants, double-buffered by tile parity: scale (demod*gain), shift (bias*gain), next style
  alignas(16) float ep_scale[2][256];
  alignas(16) float ep_shift[2][256];
  alignas(16) float ep_next[2][256];
  alignas(16) float ep_rgb[2][3][256];  // per-sample modulated ToRGB weights
  // TS flavour: the same 3072 floats viewed as [group][parity][6 = scale, shift, next, rgb0..2][128]
  __device__ __forceinline__ float* ts_consts(int group, int cb) { return &ep_scale[0][0] + (group * 2 + cb) * 768; }
};
static_assert(offsetof(Tc2Bars, ep_rgb) - offsetof(Tc2Bars, ep_scale) == 3 * 512 * sizeof(float), "ep arrays must be contiguous");

// uint8 image: exactly the fp32 operation sequence of torchvision's make_grid(normalize=True, value_range=(-1, 1)) followed
// by save_image's mul(255).add_(0.5).clamp_(0, 255).to(uint8): clamp, - low, / (high - low), * 255, + 0.5, clamp, truncate.
__device__ __forceinline__ uint8_t quant_u8(float v) {
  float t = fminf(fmaxf(v, -1.f), 1.f);
  t = __fmul_rn(__fadd_rn(t, 1.f), 0.5f);   // (division by 2 is exact)
  t = __fadd_rn(__fmul_rn(t, 255.f), 0.5f);
  return (uint8_t)fminf(fmaxf(t, 0.f), 255.f);
}

// Tile coordinates advanced without divisions: tile = ((tn*B + b)*tiles_y + ty)*tiles_x + tx and
// the per-CTA stride is decomposed once in the same mixed radix.
struct TileWalk {
  int tx, ty, b, tn, dtx, dty, db, dtn;
  __device__ __forceinline__ void init(int tile, int stride, const Tc2Params& P) {
    tx = tile % P.tiles_x; tile /= P.tiles_x;
    ty = tile % P.tiles_y; tile /= P.tiles_y;
    b = tile % P.B; tn = tile / P.B;
    dtx = stride % P.tiles_x; stride /= P.tiles_x;
    dty = stride % P.tiles_y; stride /= P.tiles_y;
    db = stride % P.B; dtn = stride / P.B;
  }
  __device__ __forceinline__ void next(const Tc2Params& P) {
    tx += dtx;
    if (tx >= P.tiles_x) { tx -= P.tiles_x; ++ty; }
    ty += dty;
    if (ty >= P.tiles_y) { ty -= P.tiles_y; ++b; }
    b += db;
    if (b >= P.B) { b -= P.B; ++tn; }
    tn += dtn;
  }
};

constexpr int kPitch = 10;  // pixels per row of the haloed tile (8 + 2 halo columns)
// skip-image patch of one sub-tile (fused ToRGB, TS epilogue): 3 planes x 10 rows x 12 columns fp32
constexpr int kSkipBoxW = 12, kSkipBoxBytes = 1536;

// Upper 32 bits of a K-major swizzled UMMA shared-memory descriptor: SBO (bytes >> 4) @32,
// version 1 @46, layout @61 (2 = SWIZZLE_128B, 4 = SWIZZLE_64B); the lower word is the start
// address >> 4.  Hardware fact validated by tools/tc2_probe.py: the swizzle is an XOR on absolute
// shared-memory address bits, so ANY 16-byte aligned start inside a TMA-written swizzled tile and
// any SBO are legal, with the base-offset field left at 0.
__device__ __forceinline__ uint32_t desc_hi(uint32_t row_bytes, uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | ((row_bytes == 128 ? 2u : 4u) << 29);
}
__device__ __forceinline__ uint64_t desc64(uint32_t hi, uint32_t addr_bytes) {
  return ((uint64_t)hi << 32) | (uint64_t)((addr_bytes & 0x3FFFFu) >> 4);
}

// tap t = ky*3+kx.  Plain 3x3: reads haloed pixel (y+ky, x+kx), one accumulator.  Transposed x2:
// out[2j+py, 2i+px] += x[j-(ky==2), i-(kx==2)] * W[ky,kx] with class (py,px) = (ky&1, kx&1).
template <bool TR>
__device__ __forceinline__ constexpr int tap_rows(int t) {
  const int ky = t / 3, kx = t % 3;
  return TR ? ((ky == 2 ? 0 : 1) * kPitch + (kx == 2 ? 0 : 1)) : (ky * kPitch + kx);
}
template <bool TR>
__device__ __forceinline__ constexpr int tap_group(int t) {
  return TR ? ((t / 3) & 1) * 2 + ((t % 3) & 1) : 0;
}
template <bool TR>
__device__ __forceinline__ constexpr bool tap_first(int t) {  // first tap (in issue order) of its accumulator
  return TR ? (t == 0 || t == 1 || t == 3 || t == 4) : (t == 0);
}

// Fused dgrad of the transposed x2 convolution: gx[j,i] = sum_{ky,kx} gz[2j+ky, 2i+kx] . W[ky,kx]^T.  With the gradient
// split into its four parity classes gzc_(py,px)[r,s] = gz[2r+py, 2s+px] (strided TMA views), tap (ky,kx) reads class
// (ky & 1, kx & 1) at offset (ky == 2, kx == 2): nine shifted reads of four class tiles, ONE accumulator.
__device__ __forceinline__ constexpr int dg4_class(int t) { return ((t / 3) & 1) * 2 + ((t % 3) & 1); }
__device__ __forceinline__ constexpr int dg4_rows(int t) { return ((t / 3) == 2 ? kPitch : 0) + ((t % 3) == 2 ? 1 : 0); }

__device__ __forceinline__ uint32_t e_base_fb(uint8_t* smem, const Tc2Params& P) { return smem_u32(smem + P.e_off); }

struct Ring {
  uint32_t idx = 0, phase = 0;
  __device__ __forceinline__ void advance(uint32_t n) {
    if (++idx == n) { idx = 0; phase ^= 1u; }
  }
  __device__ __forceinline__ void advance_by(uint32_t step, uint32_t n) {   // n is a multiple of step
    idx += step;
    if (idx >= n) { idx = 0; phase ^= 1u; }
  }
};

// TF32: fp32 activations / weights / results in HBM, tcgen05.mma kind::tf32 (the "tf32 mode" of the north star; only
// with the direct-store epilogue).  The shared-memory BYTE geometry is that of the bf16 kernel: a 128-byte row holds
// 32 fp32 channels instead of 64 bf16 ones and one MMA consumes 8 of them (32 bytes) instead of 16.
template <bool TR, int MT, int KSTEPS, bool WRES, bool RGB, bool TS, bool TF32 = false>
__global__ void __launch_bounds__(TS ? (WRES ? kT2ThreadsTS2 : kT2ThreadsTS) : kT2Threads, 1)
modconv_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ Tc2Params P, const __grid_constant__ Tc2Maps M) {
  constexpr int NG = TR ? 4 : 1;
  static_assert(!RGB || (!TR && (MT == 2 || (TS && MT == 4))), "fused ToRGB needs the plain conv with 2 (or 4) sub-tiles");
  static_assert(!TF32 || (!TS && !RGB), "the tf32 mode uses the direct-store epilogue without the fused ToRGB");
  constexpr int kRowBytes = KSTEPS * 32;  // BK * element size: 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
  constexpr int kBK = KSTEPS * (TF32 ? 8 : 16);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_base = smem;
  uint8_t* b_base = smem + (size_t)P.a_stages * P.a_stage_bytes;
  const int kchunks = P.Cin / kBK;
  const int n_bblocks = WRES ? 9 * kchunks : P.b_stages;
  Tc2Bars* bars = reinterpret_cast<Tc2Bars*>(smem + P.bars_off);
  (void)n_bblocks;
  volatile int* abort_flag = &bars->abort_flag;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 2-CTA cluster mode (weight-ring kernels): the two CTAs of a cluster work on the SAME tile position of two
  // consecutive samples, so they consume identical weight blocks: each loads half of every block and TMA-
  // multicasts it to both (half the L2->SM weight traffic).  P.B / P.ntiles then count sample PAIRS.
  const int cl = P.cluster;
  const uint32_t crank = cl ? cluster_ctarank() : 0u;
  const int cta_slot = (int)blockIdx.x >> cl, cta_slots = (int)gridDim.x >> cl;
  // Warp roles: the epilogue warps come FIRST and the TMA producer / MMA issuer LAST: the SM's warp
  // arbiter favours the highest warp id of a scheduler, and a starved MMA issuer stalls everyone
  // (measured: with the issuer as warp 1 it got an issue slot every ~9 cycles next to busy epilogue warps).
  constexpr int kEpiWarps = TS ? 16 : 8;
  // TS flavour, 20 warps: resident weights -> [16 epilogue][A producer][epilogue-input producer][2 MMA issuers];
  // weight ring -> [16 epilogue][A + epilogue-input producer][weight-ring producer][2 MMA issuers]
  constexpr int kProducerWarp = kEpiWarps, kEProducerWarp = kEpiWarps + 1, kBProducerWarp = kEpiWarps + 1;
  constexpr int kMmaWarp = kEpiWarps + (TS ? 2 : 1);
  // TS flavour with resident weights: TWO MMA-issuing warps take alternate tiles.  tcgen05.mma issue
  // blocks at the pipe's execution rate (shallow queue), so a single issuer leaves the tensor pipe idle
  // for its whole per-tile bookkeeping (barrier waits, fences, commits: ~800 cycles measured against
  // ~1600 cycles of MMAs per tile of the 32-channel layer); with two issuers one's bookkeeping hides
  // behind the other's MMAs.  Accumulator buffers and A stages are consumed in tile order by both.
  constexpr int kMmaWarps = TS ? 2 : 1;   // resident weights: alternate TILES; weight ring: alternate weight BLOCKS
  const int tiles_xy = P.tiles_x * P.tiles_y;
  const int tiles_per_n = tiles_xy * P.B;

  if (threadIdx.x == 0) {
    bars->abort_flag = 0;
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
  }
  if (warp == kMmaWarp) {
    if (lane == 0) {
      // two block-alternating MMA issuers (weight-ring TS kernels): BOTH commit every A stage and every accumulator
      // set, so no ordering between the two issuers' MMAs has to be assumed
      const uint32_t n_issuers = (TS && !WRES && !(P.flags & 2)) ? 2u : 1u;
      for (int s = 0; s < P.a_stages; ++s) { mbar_init(&bars->a_full[s], 1); mbar_init(&bars->a_empty[s], n_issuers); }
      for (int s = 0; s < P.b_stages; ++s) { mbar_init(&bars->b_full[s], 1); mbar_init(&bars->b_empty[s], 1u << cl); }
      // TS flavour with a single accumulator buffer: both epilogue groups drain every tile (unit-split mode)
      const uint32_t epi_arrivals = (TS && (P.nbuf == 1 || P.fb)) ? 2 * kT2EpiThreads : kT2EpiThreads;
      // (per-class hand-over: the four entries are the four parity classes of the single accumulator set)
      for (int s = 0; s < (P.percls ? 4 : P.nbuf); ++s) { mbar_init(&bars->acc_full[s], n_issuers); mbar_init(&bars->acc_empty[s], epi_arrivals); }
      mbar_init(&bars->w_full, 1);
      mbar_init(&bars->mma_turn[0], 1);
      mbar_init(&bars->mma_turn[1], 1);
      for (int s = 0; s < kEStages; ++s) { mbar_init(&bars->e_full[s], 1); mbar_init(&bars->e_empty[s], epi_arrivals); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&bars->tmem_slot, (uint32_t)P.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  if (cl) cluster_sync();   // the peer's mbarriers are initialised before any multicast / remote arrive targets them
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;

  if (warp == kProducerWarp) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      if (WRES) {
        mbar_arrive_expect_tx(&bars->w_full, (uint32_t)(n_bblocks * P.b_block_bytes));
        for (int kc = 0; kc < kchunks; ++kc)
          for (int t = 0; t < 9; ++t)
            tma_load_3d(b_base + (size_t)(kc * 9 + t) * P.b_block_bytes, &map_b, &bars->w_full, kc * kBK, 0, t);
      }
      Ring ar, br, er;
      bool ok = true;
      TileWalk wk;
      wk.init(cta_slot, cta_slots, P);
      for (int tile = cta_slot; tile < P.ntiles && ok; tile += cta_slots, wk.next(P)) {
        const int b = (wk.b << cl) + (int)crank;
        const int j0 = wk.ty * P.tile_dy + P.tile_o, i0 = wk.tx * P.tile_dx + P.tile_o, co0 = wk.tn * P.bn;
        if (TS && !WRES) {
          // weight-ring kernels: this warp also produces the epilogue inputs (tile coordinates, noise, skip patches)
          ok = mbar_wait(&bars->e_empty[er.idx], er.phase ^ 1u, abort_flag);
          if (!ok) break;
          uint8_t* eb = smem + P.e_off + (size_t)er.idx * P.e_stage_bytes;
          *reinterpret_cast<int4*>(eb + P.e_info_off) = make_int4(b, j0, i0, wk.tn);
          mbar_arrive_expect_tx(&bars->e_full[er.idx], (uint32_t)P.e_bytes);   // release: orders the info store
          if (P.noise && !P.fb) tma_load_3d(eb, &M.noise, &bars->e_full[er.idx], i0, j0, P.noise_per_sample ? b : 0);
          if (RGB && P.rgb_skip) {
#pragma unroll
            for (int m = 0; m < MT; ++m)
              tma_load_3d(eb + P.e_noise_bytes + m * P.skip_box_bytes, &M.skip, &bars->e_full[er.idx], (i0 >> 1) - 4,
                          ((j0 + m * kSubTileH) >> 1) - 1, b * 3);
          }
          er.advance(kEStages);
        }
        for (int kc = 0; kc < kchunks && ok; ++kc) {
          ok = mbar_wait(&bars->a_empty[ar.idx], ar.phase ^ 1u, abort_flag);
          if (!ok) break;
          if (TS && P.dbg && blockIdx.x == 0 && kc == 0 && tile / (int)gridDim.x < 64) P.dbg[(tile / gridDim.x) * 8 + 1] = clock64();
          mbar_arrive_expect_tx(&bars->a_full[ar.idx], (uint32_t)P.a_box_bytes);
          if (!TR && !TS && MT == 1 && !RGB && !TF32 && P.dg4) {   // (compile-time gate: only this instantiation has the mode)
            // the four parity-class tiles of the gradient (rows j0 .. j0+16*MT, columns i0 .. i0+8 of every class view)
            uint8_t* dst = a_base + (size_t)ar.idx * P.a_stage_bytes;
            tma_load_4d(dst, &map_a, &bars->a_full[ar.idx], kc * kBK, i0, j0, b);
#pragma unroll
            for (int c = 1; c < 4; ++c)
              tma_load_4d(dst + ((size_t)(c * P.cls16) << 4), &M.a4[c - 1], &bars->a_full[ar.idx], kc * kBK, i0, j0, b);
          } else if (!TR && P.dgk) {
            // K-loop form: chunk kc belongs to parity class kc / dgk of the gradient (its own strided view)
            const int cls = kc / P.dgk;
            tma_load_4d(a_base + (size_t)ar.idx * P.a_stage_bytes, cls == 0 ? &map_a : &M.a4[cls - 1], &bars->a_full[ar.idx],
                        (kc - cls * P.dgk) * kBK, i0 - 1, j0 - 1, b);
          } else {
            tma_load_4d(a_base + (size_t)ar.idx * P.a_stage_bytes, &map_a, &bars->a_full[ar.idx], kc * kBK, i0 - 1,
                        j0 - 1, b);
          }
          ar.advance(P.a_stages);
          if (!WRES && !TS) {
            const bool edge_y = TR && !(P.flags & 1) && j0 >= P.grid_h - 1;   // see the MMA issuer
            const bool edge_x = TR && !(P.flags & 1) && i0 >= P.grid_w - 1;
            for (int t = 0; t < 9; ++t) {
              if ((edge_y && t / 3 != 2) || (edge_x && t % 3 != 2)) continue;
              if (!TR && !((P.tap_mask >> t) & 1)) continue;
              ok = mbar_wait(&bars->b_empty[br.idx], br.phase ^ 1u, abort_flag);
              if (!ok) break;
              mbar_arrive_expect_tx(&bars->b_full[br.idx], (uint32_t)P.b_block_bytes);
              tma_load_3d(b_base + (size_t)br.idx * P.b_block_bytes, &map_b, &bars->b_full[br.idx], kc * kBK, co0, t);
              br.advance(P.b_stages);
            }
          }
        }
      }
    }
  } else if (TS && !WRES && warp == kBProducerWarp) {
    if (lane == 0) {
      // ------------------------------------------------------------------ weight-ring producer (TS flavour)
      // its own warp, so that the A tiles / epilogue inputs and the weight blocks are issued concurrently
      const int pj = 0;
      Ring br;
      bool ok = true;
      uint32_t nblk = 0;
      TileWalk wk;
      wk.init(cta_slot, cta_slots, P);
      for (int tile = cta_slot; tile < P.ntiles && ok; tile += cta_slots, wk.next(P)) {
        const int j0 = wk.ty * P.tile_dy + P.tile_o, i0 = wk.tx * P.tile_dx + P.tile_o, co0 = wk.tn * P.bn;
        const bool edge_y = TR && !(P.flags & 1) && !P.fb && j0 >= P.grid_h - 1;   // see the MMA issuer
        const bool edge_x = TR && !(P.flags & 1) && !P.fb && i0 >= P.grid_w - 1;
        if (P.bgroup == 3) {
          // One request per TAP ROW: a box {BK, bn, 3} fills three consecutive ring stages and signals the first
          // stage's barrier.  A single thread issues a TMA request every ~330...400 cycles whatever its size
          // (tools/tma_ingest_bench.cu), and a 16 KB block feeds only 256 cycles of MMAs: with per-tap requests the
          // ring was bound by the REQUEST RATE, not by bytes.  (Edge-column tiles use only kx == 2 of each row: the
          // whole row is still fetched.)
          for (int kc = 0; kc < kchunks && ok; ++kc) {
            const int kmask = (!TR && P.dgk) ? (int)((P.dg_masks >> (9 * (kc / P.dgk))) & 0x1ff) : P.tap_mask;
            for (int gi = 0; gi < 3; ++gi) {
              const int g = TR ? (gi == 0 ? 1 : (gi == 1 ? 0 : 2)) : gi;   // transposed: tap rows in the order 1, 0, 2 (see the issuer)
              if (edge_y && g != 2) continue;
              if (!TR && !((kmask >> (3 * g)) & 7)) continue;   // no tap of this row is used
              ok = mbar_wait(&bars->b_empty[br.idx], br.phase ^ 1u, abort_flag);
              if (!ok) break;
              mbar_arrive_expect_tx(&bars->b_full[br.idx], 3u * (uint32_t)P.b_block_bytes);
              tma_load_3d(b_base + (size_t)br.idx * P.b_block_bytes, &M.b3, &bars->b_full[br.idx], kc * kBK, co0, 3 * g);
              br.advance_by(3, P.b_stages);
            }
          }
          continue;
        }
        for (int kc = 0; kc < kchunks && ok; ++kc) {
          const int kmask = (!TR && P.dgk) ? (int)((P.dg_masks >> (9 * (kc / P.dgk))) & 0x1ff) : P.tap_mask;
          for (int t = 0; t < 9; ++t) {
            if ((edge_y && t / 3 != 2) || (edge_x && t % 3 != 2)) continue;
            if (!TR && !((kmask >> t) & 1)) continue;
            if (pj == 0) {
              ok = mbar_wait(&bars->b_empty[br.idx], br.phase ^ 1u, abort_flag);
              if (!ok) break;
              mbar_arrive_expect_tx(&bars->b_full[br.idx], (uint32_t)P.b_block_bytes);
              if (cl)   // my share of the block, delivered to every CTA of the cluster (each signals its own b_full)
                tma_load_3d_mc(b_base + (size_t)br.idx * P.b_block_bytes + (size_t)crank * (P.b_block_bytes >> cl), &M.bh,
                               &bars->b_full[br.idx], kc * kBK, co0 + (int)crank * (P.bn >> cl), t,
                               (uint16_t)((1u << (1 << cl)) - 1u));
              else
                tma_load_3d(b_base + (size_t)br.idx * P.b_block_bytes, &map_b, &bars->b_full[br.idx], kc * kBK, co0, t);
            }
            ++nblk;
            br.advance(P.b_stages);
          }
        }
      }
    }
  } else if (TS && WRES && warp == kEProducerWarp) {
    if (lane == 0) {
      // ------------------------------------------------------------------ epilogue-input producer
      // Its own warp: a single thread issuing every TMA of a tile took ~2200 cycles per tile (measured),
      // more than the MMAs of the 32-channel layers.  Per tile: the tile coordinates, the noise patch and
      // (fused ToRGB) the skip-image patch of each sub-tile.
      Ring er;
      bool ok = true;
      TileWalk wk;
      wk.init(cta_slot, cta_slots, P);
      for (int tile = cta_slot; tile < P.ntiles && ok; tile += cta_slots, wk.next(P)) {
        const int b = (wk.b << cl) + (int)crank;
        const int j0 = wk.ty * P.tile_dy + P.tile_o, i0 = wk.tx * P.tile_dx + P.tile_o;
        ok = mbar_wait(&bars->e_empty[er.idx], er.phase ^ 1u, abort_flag);
        if (!ok) break;
        if (P.dbg && blockIdx.x == 0 && tile / (int)gridDim.x < 64) P.dbg[(tile / gridDim.x) * 8 + 0] = clock64();
        uint8_t* eb = smem + P.e_off + (size_t)er.idx * P.e_stage_bytes;
        *reinterpret_cast<int4*>(eb + P.e_info_off) = make_int4(b, j0, i0, wk.tn);
        mbar_arrive_expect_tx(&bars->e_full[er.idx], (uint32_t)P.e_bytes);   // release: orders the info store
        // (pair mode: i0 counts pixel pairs = columns of the half-resolution skip image; the noise patch is 16 pixels wide)
        if (P.noise && !P.fb) tma_load_3d(eb, &M.noise, &bars->e_full[er.idx], P.pair ? 2 * i0 : i0, j0, P.noise_per_sample ? b : 0);
        if (RGB && P.rgb_skip) {
#pragma unroll
          for (int m = 0; m < MT; ++m)
            tma_load_3d(eb + P.e_noise_bytes + m * P.skip_box_bytes, &M.skip, &bars->e_full[er.idx],
                        (P.pair ? i0 : (i0 >> 1)) - 4, ((j0 + m * kSubTileH) >> 1) - 1, b * 3);
        }
        er.advance(kEStages);
      }
    }
  } else if (warp >= kMmaWarp) {
    // -------------------------------------------------------------------- MMA issuer(s)
    const int mw = warp - kMmaWarp;   // issuer index: tiles mw, mw + kMmaWarps, ... of this CTA's sequence
    // The whole warp runs the loop (warp-uniform control flow keeps descriptors in uniform
    // registers); only tcgen05.mma / tcgen05.commit are issued, by one elected lane.
    // instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major both, N>>3 @17, M>>4 @24
    // (kind::tf32: A = B = tf32, format code 2)
    const uint32_t idesc = (1u << 4) | ((TF32 ? 2u : 1u) << 7) | ((TF32 ? 2u : 1u) << 10) | ((uint32_t)(P.bn >> 3) << 17) |
                           ((128u >> 4) << 24);
    const uint32_t a_hi = desc_hi(kRowBytes, kPitch * kRowBytes);
    const uint32_t b_hi = desc_hi(kRowBytes, 8 * kRowBytes);
    // descriptor low words (start address >> 4) are LINEAR in the address: one add per operand per MMA
    const uint32_t a_lo0 = (smem_u32(a_base) & 0x3FFFFu) >> 4, b_lo0 = (smem_u32(b_base) & 0x3FFFFu) >> 4;
    const uint32_t a_stage16 = (uint32_t)P.a_stage_bytes >> 4, b_block16 = (uint32_t)P.b_block_bytes >> 4;
    const uint32_t bn = (uint32_t)P.bn;
    const int first_tap = __ffs(P.tap_mask) - 1;   // plain conv: the first USED tap overwrites the accumulator
    const int last_tap = TR ? 8 : 31 - __clz(P.tap_mask);
    Ring ar, br, cr;
    bool ok = true;
    // warp-uniform issue: every lane runs the loops (descriptors stay in uniform registers, per-tile state stays
    // consistent across lanes); only the elected lane executes tcgen05.mma / tcgen05.commit
    const bool leader = elect_one();
    if (WRES) {
      ok = mbar_wait_warp(&bars->w_full, 0, abort_flag);
      // (no tcgen05 fence: TMA -> mbarrier -> tcgen05.mma needs none)
    }
    // Two issuers only when each accumulator buffer and each A stage is always consumed by the SAME issuer
    // (even buffer count, stage count a multiple of 2 * kchunks): an issuer that saw only every other phase
    // of an mbarrier could mistake an old completion for the one it waits for.
    const int nmw = (WRES && kMmaWarps == 2 && !(P.flags & 2) && P.nbuf >= 2 && P.a_stages % (2 * kchunks) == 0) ? 2 : 1;
    // Weight-ring kernels: both issuers walk every tile and take alternate weight BLOCKS (one tap of one K
    // chunk), again ordered by the issue token.  The per-block bookkeeping of a single issuer (b_full wait,
    // commits, ring arithmetic: ~400 cycles next to 16 busy epilogue warps) exceeded the 256...384 cycles of MMAs
    // a block feeds (measured: 94 cycles per MMA instead of 48 at 128->64@256^2).
    const bool blk2 = !WRES && kMmaWarps == 2 && !(P.flags & 2);
    const int tile_step = nmw * cta_slots;
    const int tile0 = (mw < nmw || blk2) ? cta_slot + (blk2 ? 0 : mw * cta_slots) : P.ntiles;
    uint32_t nblk = 0, myblk = 0;   // weight blocks seen / issued by this warp (block-alternating mode)
    if (mw == 1 && !blk2) {   // the second issuer starts one tile into the rings
      for (int i = 0; i < kchunks; ++i) ar.advance(P.a_stages);
      cr.advance(P.nbuf);
    }
    TileWalk wk;
    wk.init(tile0 < P.ntiles ? tile0 : 0, tile_step, P);
    uint32_t turn = 0;   // tiles issued by this warp so far (parity of its token waits)
    for (int tile = tile0; tile < P.ntiles && ok; tile += tile_step, wk.next(P), ++turn) {
      // Transposed conv: the class grids are (h+1) x (w+1), so the last tile row / column holds only
      // the positions j == h / i == w, which see nothing but zero padding through every tap except
      // ky == 2 / kx == 2: the other taps (and the second sub-tile) of such EDGE tiles are skipped.
      // Their untouched accumulators belong to classes whose rows/columns are clipped by the stores.
      const bool edge_y = TR && !(P.flags & 1) && !P.fb && wk.ty * (kSubTileH * MT) >= P.grid_h - 1;
      const bool edge_x = TR && !(P.flags & 1) && !P.fb && wk.tx * kTileW >= P.grid_w - 1;
      // PER-CLASS HAND-OVER (transposed conv whose 4 classes x MT x bn accumulators fill all 512 TMEM columns, so there is
      // no second accumulator set): the tap rows are issued in the order 1, 0, 2 -- row 1 feeds exactly the classes
      // (1,0) and (1,1), rows 0 and 2 the classes (0,0) and (0,1) -- so in the last K chunk classes 2 and 3 are complete
      // after the FIRST row (they are committed there and the epilogue starts draining them while rows 0 and 2 still run)
      // and in the first K chunk of the next tile row 1 only needs classes 2 and 3 drained, rows 0 / 2 classes 0 and 1.
      // Measured before: MMA phase and a ~4.5k-cycle drain strictly alternated (tools/tc2_timeline.py).
      if (!(TR && P.percls)) {
        ok = mbar_wait_warp(&bars->acc_empty[cr.idx], cr.phase ^ 1u, abort_flag);
        if (!ok) break;
      }
      tc_fence_after();
      const bool dbg_on = TS && P.dbg && blockIdx.x == 0 && tile / (int)gridDim.x < 64 && lane == 0;
      if (dbg_on) P.dbg[(tile / gridDim.x) * 8 + 2] = clock64();
      uint32_t dcol[NG * MT];   // TMEM column of every accumulator of this tile
      dcol[0] = tmem_base + cr.idx * (uint32_t)(NG * MT) * bn;
#pragma unroll
      for (int i = 1; i < NG * MT; ++i) dcol[i] = dcol[i - 1] + bn;
      uint32_t started = 0;     // accumulators that already received their first (overwriting) MMA
      for (int kc = 0; kc < kchunks && ok; ++kc) {
        ok = mbar_wait_warp(&bars->a_full[ar.idx], ar.phase, abort_flag);
        if (!ok) break;
        // (no tcgen05 fence: TMA -> mbarrier -> tcgen05.mma needs none, and it was measured to drain the MMA queue)
        const uint32_t a_lo = a_lo0 + ar.idx * a_stage16;
        if (nmw == 2 && kc == 0 && !TR) {
          // (not for the transposed kernels: their four class accumulators make a tile's issue long enough that two
          // concurrent issuers win -- up 64->32 @512^2 0.617 -> 0.590 ms in the same binary, while the plain 32 / 64-channel
          // kernels lose 3...6 % without the token; profiles/r02_ab_issue_token.log)
          // issue token: the other issuer's MMAs of the previous tile are all queued.  Without it the two
          // warps interleave their MMAs, finish together and do their bookkeeping at the same time (measured).
          ok = mbar_wait_warp(&bars->mma_turn[mw], mw == 0 ? (turn & 1u) ^ 1u : (turn & 1u), abort_flag);
          if (!ok) break;
        }
        if (dbg_on && kc == 0) P.dbg[(tile / gridDim.x) * 8 + 3] = clock64();
        if (WRES) {
          uint32_t b_lo = b_lo0 + (uint32_t)(kc * 9) * b_block16;
          if constexpr (!TR && TS) {
            // plain conv: the three tap ROWS as a rolled loop (a third of the code: the issuing warp shares its scheduler's
            // instruction cache with four epilogue warps)
#pragma unroll 1
            for (int ky = 0; ky < 3; ++ky) {
              const uint32_t a_row = a_lo + (uint32_t)(ky * kPitch * kRowBytes / 16);
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const int t = ky * 3 + kx;
                if ((P.tap_mask >> t) & 1) {
#pragma unroll
                  for (int m = 0; m < MT; ++m) {
                    constexpr int kSub16 = kSubTileH * kPitch * kRowBytes / 16;
                    const uint32_t a_tap = a_row + (uint32_t)(kx * kRowBytes / 16 + m * kSub16);
                    const uint32_t first = (t == first_tap && kc == 0) ? 0u : 1u;
                    if (leader) {
                      const bool pr = KSTEPS == 4 && P.pair;
                      const int kfirst = (pr && kx == 0) ? 2 : 0;
#pragma unroll
                      for (int k = 0; k < KSTEPS; ++k) {
                        if (pr && ((kx == 0 && k < 2) || (kx == 2 && k >= 2))) continue;
                        umma_bf16_lohi(dcol[m], a_tap + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, k == kfirst ? first : 1u);
                      }
                    }
                  }
                }
                b_lo += b_block16;
              }
            }
          } else {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            if (!((edge_y && t / 3 != 2) || (edge_x && t % 3 != 2)) && (TR || ((P.tap_mask >> t) & 1))) {
#pragma unroll
              for (int m = 0; m < MT; ++m) {
                if (m == 1 && edge_y) continue;
                constexpr int kSub16 = kSubTileH * kPitch * kRowBytes / 16;
                const int ai = tap_group<TR>(t) * MT + m;
                // (the fused dgrad of the transposed convolution only exists in this instantiation: the issue loop of the
                // 32-channel kernels is issue-bound, and a run-time select per MMA cost them 0.50 -> 0.73 ms at 1024^2)
                constexpr bool kDg4 = !TR && !TS && MT == 1 && !RGB && !TF32;
                const uint32_t a_tap = (kDg4 && P.dg4)
                    ? a_lo + (uint32_t)(dg4_class(t) * P.cls16 + dg4_rows(t) * kRowBytes / 16 + m * kSub16)
                    : a_lo + (uint32_t)(tap_rows<TR>(t) * kRowBytes / 16 + m * kSub16);
                uint32_t first = 1u;
                if (TR) {
                  first = (kc == 0 && !((started >> ai) & 1u)) ? 0u : 1u;
                  started |= 1u << ai;
                } else if (t == first_tap) {
                  first = kc == 0 ? 0u : 1u;
                }
                if (leader) {
                  // pair mode (KSTEPS == 4: K = [pixel 2i | pixel 2i+1] x 32 channels): the left neighbour pair (t % 3 == 0)
                  // contributes only its RIGHT pixel (K steps 2, 3), the right neighbour pair only its LEFT pixel (0, 1)
                  const bool pr = KSTEPS == 4 && P.pair;
                  const int kfirst = (pr && t % 3 == 0) ? 2 : 0;
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k) {
                    if (pr && ((t % 3 == 0 && k < 2) || (t % 3 == 2 && k >= 2))) continue;
                    if constexpr (TF32) umma_tf32_lohi(dcol[ai], a_tap + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, k == kfirst ? first : 1u);
                    else umma_bf16_lohi(dcol[ai], a_tap + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, k == kfirst ? first : 1u);
                  }
                }
              }
            }
            b_lo += b_block16;
          }
          }
          if (leader) umma_commit(&bars->a_empty[ar.idx]);
        } else if (TS && P.bgroup == 3) {
          // grouped weight requests: one b_full / b_empty round trip (and one issue token) per TAP ROW of 3 blocks
          // (K-loop dgrad of the transposed convolution: the taps of THIS chunk's parity class)
          const int kmask = (!TR && P.dgk) ? (int)((P.dg_masks >> (9 * (kc / P.dgk))) & 0x1ff) : P.tap_mask;
          const int last_row = TR ? 2 : (31 - __clz(kmask)) / 3;
#pragma unroll
          for (int gi = 0; gi < 3; ++gi) {
            const int g = TR ? (gi == 0 ? 1 : (gi == 1 ? 0 : 2)) : gi;
            if (TR && P.percls && kc == 0 && gi < 2) {
              // first K chunk: the classes this row starts must have been drained (row 1: classes 2, 3; row 0: 0, 1)
              const int c0 = gi == 0 ? 2 : 0;
              ok = mbar_wait_warp(&bars->acc_empty[c0], cr.phase ^ 1u, abort_flag) &&
                   mbar_wait_warp(&bars->acc_empty[c0 + 1], cr.phase ^ 1u, abort_flag);
              if (!ok) break;
              tc_fence_after();
            }
            const bool skip_row = (edge_y && g != 2) || (!TR && !((kmask >> (3 * g)) & 7));
            if (TR && P.percls && kc == kchunks - 1 && skip_row && leader && gi != 1) {
              // (a skipped row still is a commit point of its classes: the hand-over protocol is the same for every tile)
              const int c0 = gi == 0 ? 2 : 0;
              umma_commit(&bars->acc_full[c0]);
              umma_commit(&bars->acc_full[c0 + 1]);
            }
            if (skip_row) continue;
            const bool mine = !blk2 || (int)(nblk & 1u) == mw;
            ++nblk;
            if (mine) {
              ok = mbar_wait_warp(&bars->b_full[br.idx], br.phase, abort_flag);
              if (ok && blk2) ok = mbar_wait_warp(&bars->mma_turn[mw], mw == 0 ? (myblk & 1u) ^ 1u : (myblk & 1u), abort_flag);
              if (!ok) break;
              ++myblk;
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int t = g * 3 + kx;
              if (edge_x && kx != 2) continue;
              if (!TR && !((kmask >> t) & 1)) continue;
              const uint32_t b_lo = b_lo0 + (br.idx + (uint32_t)kx) * b_block16;
#pragma unroll
              for (int m = 0; m < MT; ++m) {
                if (m == 1 && edge_y) continue;
                constexpr int kSub16 = kSubTileH * kPitch * kRowBytes / 16;
                const int ai = tap_group<TR>(t) * MT + m;
                const uint32_t a_tap = a_lo + (uint32_t)(tap_rows<TR>(t) * kRowBytes / 16 + m * kSub16);
                uint32_t first = 1u;
                if (TR) {
                  first = (kc == 0 && !((started >> ai) & 1u)) ? 0u : 1u;
                  started |= 1u << ai;
                } else if (t == first_tap) {
                  first = kc == 0 ? 0u : 1u;
                }
                if (leader && mine) {
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k)
                    umma_bf16_lohi(dcol[ai], a_tap + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, k == 0 ? first : 1u);
                }
              }
            }
            if (leader && mine) {
              umma_commit(&bars->b_empty[br.idx]);
              if (!blk2 && g == last_row) {
                umma_commit(&bars->a_empty[ar.idx]);
                if (kc == kchunks - 1 && !(TR && P.percls)) umma_commit(&bars->acc_full[cr.idx]);
              }
              if (blk2) mbar_arrive(&bars->mma_turn[mw ^ 1]);
            }
            if (TR && P.percls && kc == kchunks - 1 && leader && gi != 1 && (blk2 || mine)) {
              // last K chunk: row 1 completed classes 2 and 3, row 2 (the last one) classes 0 and 1.  Every issuer commits
              // (the barrier counts the issuers), each covering the MMAs it issued itself.
              const int c0 = gi == 0 ? 2 : 0;
              umma_commit(&bars->acc_full[c0]);
              umma_commit(&bars->acc_full[c0 + 1]);
            }
            br.advance_by(3, P.b_stages);
          }
          if (blk2 && ok && leader) {
            umma_commit(&bars->a_empty[ar.idx]);
            if (kc == kchunks - 1 && !(TR && P.percls)) umma_commit(&bars->acc_full[cr.idx]);
          }
        } else {
          const int kmask = (!TR && TS && P.dgk) ? (int)((P.dg_masks >> (9 * (kc / P.dgk))) & 0x1ff) : P.tap_mask;
          const int last_tap_k = (!TR && TS && P.dgk) ? 31 - __clz(kmask) : last_tap;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            if ((edge_y && t / 3 != 2) || (edge_x && t % 3 != 2)) continue;   // (tap 8 is never skipped)
            if (!TR && !((kmask >> t) & 1)) continue;
            const bool mine = !blk2 || (int)(nblk & 1u) == mw;
            ++nblk;
            if (mine) {
              ok = mbar_wait_warp(&bars->b_full[br.idx], br.phase, abort_flag);
              if (ok && blk2) ok = mbar_wait_warp(&bars->mma_turn[mw], mw == 0 ? (myblk & 1u) ^ 1u : (myblk & 1u), abort_flag);
              if (!ok) break;
              ++myblk;
            }
            // (no tcgen05 fence: TMA -> mbarrier -> tcgen05.mma needs none)
            const uint32_t b_lo = b_lo0 + br.idx * b_block16;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              if (m == 1 && edge_y) continue;
              constexpr int kSub16 = kSubTileH * kPitch * kRowBytes / 16;
              const int ai = tap_group<TR>(t) * MT + m;
              const uint32_t a_tap = a_lo + (uint32_t)(tap_rows<TR>(t) * kRowBytes / 16 + m * kSub16);
              uint32_t first = 1u;
              if (TR) {
                first = (kc == 0 && !((started >> ai) & 1u)) ? 0u : 1u;
                started |= 1u << ai;
              } else if (t == first_tap) {
                first = kc == 0 ? 0u : 1u;
              }
              if (leader && mine) {
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  if constexpr (TF32) umma_tf32_lohi(dcol[ai], a_tap + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, k == 0 ? first : 1u);
                  else umma_bf16_lohi(dcol[ai], a_tap + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, k == 0 ? first : 1u);
                }
              }
            }
            if (leader && mine) {
              if (cl) umma_commit_mc(&bars->b_empty[br.idx], (uint16_t)((1u << (1 << cl)) - 1u));   // free once ALL CTAs consumed it
              else umma_commit(&bars->b_empty[br.idx]);
              if (!blk2 && t == last_tap_k) {
                umma_commit(&bars->a_empty[ar.idx]);
                if (kc == kchunks - 1) umma_commit(&bars->acc_full[cr.idx]);   // last block of the tile
              }
              if (blk2) mbar_arrive(&bars->mma_turn[mw ^ 1]);
            }
            br.advance(P.b_stages);
          }
          if (blk2 && ok && leader) {   // this issuer's MMAs on the A stage (and, at the last chunk, on the tile) are queued
            umma_commit(&bars->a_empty[ar.idx]);
            if (kc == kchunks - 1) umma_commit(&bars->acc_full[cr.idx]);
          }
        }
        ar.advance(P.a_stages);
      }
      if (ok && leader && WRES) {
        umma_commit(&bars->acc_full[cr.idx]);
        if (nmw == 2) mbar_arrive(&bars->mma_turn[mw ^ 1]);
      }
      __syncwarp();
      if (dbg_on) P.dbg[(tile / gridDim.x) * 8 + 4] = clock64();
      cr.advance(P.nbuf);
      if (nmw == 2) {   // skip the other issuer's tile in both rings
        for (int i = 0; i < kchunks; ++i) ar.advance(P.a_stages);
        cr.advance(P.nbuf);
      }
    }

  } else if constexpr (TS) {
    // ------------------------------------------------------------------ epilogue, TS flavour
    // 16 epilogue warps = two groups of 8; group g drains the tiles that land in accumulator
    // buffer g (tile parity), so one group's TMEM/shared-memory latencies hide behind the other's
    // arithmetic (with a single accumulator buffer only group 0 works).  Inside a group, thread
    // (q, lane) owns accumulator row r = pixel (sy, sx) of a sub-tile; work units are
    // (accumulator, <=64-channel chunk) and the group's two halves take alternate units (fused
    // ToRGB: half = sub-tile, because a thread then needs its pixel's whole channel row).
    // Tile coordinates come from the producer (16 bytes in the epilogue-input stage) and the
    // per-channel constants are re-staged only when the (sample, channel block) changes, so the
    // per-tile bookkeeping is a handful of instructions.
    const int ew = warp;
    const int group = ew >> 3;
    const int half = (ew >> 2) & 1;
    const int q = warp & 3;
    const int gt = (int)threadIdx.x - group * kT2EpiThreads;   // 0..255 inside the group
    if constexpr (TR) {
      if (P.fb) {
        // -------------------------------------------------------------- fused up-convolution + Blur epilogue
        // (model.py:249-260 + 279-290 + fused_act: conv_transpose2d -> demod -> upfirdn2d(4x4, pad (1,1)) -> + noise
        // -> + bias -> leaky-ReLU*sqrt2 -> next layer's style.)  The (2h+1)^2 pre-blur tensor never reaches HBM:
        // the tile's 64 x 16 pre-blur pixels (32 x 8 input positions x 4 parity classes) are written, demodulated,
        // as bf16 into a swizzled shared-memory tile [64][16][32 channels]; then all 16 epilogue warps run the
        // separable FIR over it (register ring of horizontally filtered rows, packed fp32x2 math) and store the
        // 60 x 12 output pixels the tile fully covers.  Tiles therefore step by 30 x 6 input positions (the
        // MMA recomputes the 2 x 2 halo: x1.42 MMA work for 8 GB less HBM traffic per 1024^2 batch-32 step).
        const int t = (int)threadIdx.x;          // 0..511
        const int q = warp & 3, quad = warp >> 2;
        const int r = q * 32 + lane, sy = r >> 3, sx = r & 7;
        const bool lrelu = P.act == W2E_ACT_LRELU;
        const float gain = lrelu ? 1.41421356237309515f : 1.f;
        const float slope = lrelu ? 0.2f : 1.f;
        const uint64_t slope2 = pack2(slope, slope);
        const float nw = P.noise ? __ldg(P.noise_w) * gain : 0.f;
        const uint32_t z_base = smem_u32(smem + P.ts_off);
        const uint32_t nbuf_mask = (uint32_t)P.nbuf - 1u, nbuf_shift = P.nbuf == 4 ? 2u : (P.nbuf == 2 ? 1u : 0u);
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        // FIR role: 8 channels (cg) of output column fx, six consecutive output rows (segment seg)
        const int cg = t & 3, fx = (t >> 2) % 12, seg = t / 48;
        const bool fir_on = t < 480;
        const uint64_t fv0 = pack2(P.fbv[0] * gain, P.fbv[0] * gain), fv1 = pack2(P.fbv[1] * gain, P.fbv[1] * gain);
        const uint64_t fv2 = pack2(P.fbv[2] * gain, P.fbv[2] * gain), fv3 = pack2(P.fbv[3] * gain, P.fbv[3] * gain);
        const uint64_t fh0 = pack2(P.fbh[0], P.fbh[0]), fh1 = pack2(P.fbh[1], P.fbh[1]);
        const uint64_t fh2 = pack2(P.fbh[2], P.fbh[2]), fh3 = pack2(P.fbh[3], P.fbh[3]);
        auto zaddr = [&](int zr, int zc, int ck) -> uint32_t {   // 16-byte chunk ck (8 channels) of pre-blur pixel (zr, zc)
          return z_base + (uint32_t)(zr * 1024 + ((zc ^ ((zc >> 3) & 1)) * 64) + ((ck ^ ((zc >> 1) & 3)) * 16));
        };
        uint32_t gen = 0;
        int prev_b = -1, prev_tn = -1;
        bool ok = true;
        uint32_t k = 0;
        for (int tile = cta_slot; tile < P.ntiles; tile += cta_slots, ++k) {
          const uint32_t es = k % (uint32_t)kEStages;
          if (ok) ok = mbar_wait(&bars->e_full[es], (k / (uint32_t)kEStages) & 1u, abort_flag);
          int b, j0, i0, tn;
          lds_4i(e_base_fb(smem, P) + es * (uint32_t)P.e_stage_bytes + (uint32_t)P.e_info_off, b, j0, i0, tn);
          mbar_arrive(&bars->e_empty[es]);
          const bool dbg_on = P.dbg && blockIdx.x == 0 && k < 64 && t == 0;
          if (dbg_on) P.dbg[k * 8 + 5] = clock64();
          if (b != prev_b || tn != prev_tn) {
            prev_b = b; prev_tn = tn;
            gen ^= 1u;
            if (t < P.bn) {
              float* cst = bars->ts_consts(0, (int)gen);
              const int c = tn * P.bn + t, bc = b * P.Cout + c;
              cst[t] = P.out_scale ? __ldg(P.out_scale + bc) : 1.f;
              cst[128 + t] = (P.bias ? __ldg(P.bias + c) : 0.f) * gain;
              cst[256 + t] = P.next_scale ? __ldg(P.next_scale + bc) : 1.f;
            }
            named_bar_sync(7, 2 * kT2EpiThreads);
          }
          const uint32_t sc_a = smem_u32(bars->ts_consts(0, (int)gen));
          const uint32_t ci = k & nbuf_mask;
          const uint32_t t_tile = t_lane + ci * (uint32_t)(NG * MT * P.bn);
          if (ok) ok = mbar_wait(&bars->acc_full[ci], (k >> nbuf_shift) & 1u, abort_flag);
          tc_fence_after();
          if (dbg_on) P.dbg[k * 8 + 6] = clock64();
          const int Y0 = 2 * (j0 + 1), X0 = 2 * (i0 + 1);   // first output pixel of the tile
          const int nchunks = P.bn >> 5;
          for (int chunk = 0; chunk < nchunks; ++chunk) {
            // ---- phase 1: accumulators -> demodulated bf16 pre-blur tile
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int acc = quad + 4 * u;              // = class * MT + m (MT == 2)
              const int g = acc >> 1, m = acc & 1;
              const int zr = 2 * (m * kSubTileH + sy) + (g >> 1), zc = 2 * sx + (g & 1);
#pragma unroll
              for (int c16 = 0; c16 < 32; c16 += 16) {
                uint32_t v[16];
                tmem_ld16(t_tile + (uint32_t)(acc * P.bn + chunk * 32 + c16), v);
                tmem_ld_wait();
                if (chunk == nchunks - 1 && u == 1 && c16 == 16) {
                  tc_fence_before();
                  mbar_arrive(&bars->acc_empty[ci]);
                }
                const uint32_t ca = sc_a + (uint32_t)((chunk * 32 + c16) * 4);
                uint32_t pk[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  uint64_t a01, a23;
                  lds_2x2(ca + e * 16, a01, a23);
                  float x0, x1, x2, x3;
                  unpack2(mul2(pack2u(v[4 * e], v[4 * e + 1]), a01), x0, x1);
                  unpack2(mul2(pack2u(v[4 * e + 2], v[4 * e + 3]), a23), x2, x3);
                  pk[2 * e] = cvt_bf16x2(x0, x1);
                  pk[2 * e + 1] = cvt_bf16x2(x2, x3);
                }
                sts_128(zaddr(zr, zc, c16 >> 3), pk[0], pk[1], pk[2], pk[3]);
                sts_128(zaddr(zr, zc, (c16 >> 3) + 1), pk[4], pk[5], pk[6], pk[7]);
              }
            }
            named_bar_sync(7, 2 * kT2EpiThreads);
            if (dbg_on && chunk == 0) P.dbg[k * 8 + 0] = clock64();   // (overwrites the producer's stamp: FB timeline)
            // ---- phase 2: separable 4x4 FIR + noise + bias + activation + next style
            if (fir_on) {
              const int c0 = tn * P.bn + chunk * 32 + cg * 8;
              uint64_t bias2[4], nsc2[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(bias2[e]) : "r"(sc_a + (uint32_t)((128 + chunk * 32 + cg * 8 + 2 * e) * 4)));
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(nsc2[e]) : "r"(sc_a + (uint32_t)((256 + chunk * 32 + cg * 8 + 2 * e) * 4)));
              }
              const int Xg = X0 + fx;
              const bool x_ok = Xg < P.OW2;
              // the six noise values of this thread's outputs, all in flight before the filter loop
              float nzv[6];
#pragma unroll
              for (int rr = 0; rr < 6; ++rr) {
                const int Y = Y0 + 6 * seg + rr;
                nzv[rr] = (P.noise && x_ok && Y < P.OH2)
                    ? __ldg(P.noise + (P.noise_per_sample ? (int64_t)b * P.OH2 * P.OW2 : 0) + (int64_t)Y * P.OW2 + Xg) : 0.f;
              }
              uint32_t zcol[4];   // swizzled column/chunk part of the four window columns
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) zcol[kk] = zaddr(6 * seg + 1, fx + 1 + kk, cg);
              uint64_t win[4][4];
#pragma unroll
              for (int lr = 0; lr < 9; ++lr) {
                const int u4 = lr & 3;
                uint64_t f[4][4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  uint32_t w0, w1, w2, w3;
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                               : "r"(zcol[kk] + (uint32_t)(lr * 1024)));
                  f[kk][0] = pack2u(w0 << 16, w0 & 0xffff0000u);
                  f[kk][1] = pack2u(w1 << 16, w1 & 0xffff0000u);
                  f[kk][2] = pack2u(w2 << 16, w2 & 0xffff0000u);
                  f[kk][3] = pack2u(w3 << 16, w3 & 0xffff0000u);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  win[u4][e] = fma2(fh3, f[3][e], fma2(fh2, f[2][e], fma2(fh1, f[1][e], mul2(fh0, f[0][e]))));
                if (lr >= 3) {
                  const int Y = Y0 + 6 * seg + lr - 3;
                  if (x_ok && Y < P.OH2 && ok) {
                    const float nz = nw * nzv[lr - 3];
                    const uint64_t nz2 = pack2(nz, nz);
                    uint32_t po[4], pm[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const uint64_t a = fma2(fv3, win[u4][e],
                                              fma2(fv2, win[(u4 + 3) & 3][e],
                                                   fma2(fv1, win[(u4 + 2) & 3][e],
                                                        fma2(fv0, win[(u4 + 1) & 3][e], add2(bias2[e], nz2)))));
                      const uint64_t a_s = mul2(a, slope2);
                      float x0, x1, y0, y1;
                      unpack2(a, x0, x1);
                      unpack2(a_s, y0, y1);
                      x0 = fmaxf(x0, y0);
                      x1 = fmaxf(x1, y1);
                      if (P.out) po[e] = cvt_bf16x2(x0, x1);
                      if (P.out_mod) {
                        float m0, m1;
                        unpack2(mul2(pack2(x0, x1), nsc2[e]), m0, m1);
                        pm[e] = cvt_bf16x2(m0, m1);
                      }
                    }
                    const int64_t o = (((int64_t)b * P.OH2 + Y) * P.OW2 + Xg) * P.Cout + c0;
                    if (P.out) *reinterpret_cast<uint4*>(P.out + o) = make_uint4(po[0], po[1], po[2], po[3]);
                    if (P.out_mod) *reinterpret_cast<uint4*>(P.out_mod + o) = make_uint4(pm[0], pm[1], pm[2], pm[3]);
                  }
                }
              }
            }
            named_bar_sync(7, 2 * kT2EpiThreads);   // the tile may be overwritten
            if (dbg_on && chunk == 0) P.dbg[k * 8 + 1] = clock64();
          }
          if (dbg_on) P.dbg[k * 8 + 7] = clock64();
        }
      }
    }
    if (!(TR && P.fb))
    {
      const bool split = P.nbuf == 1;   // one accumulator buffer: the two groups share every tile and split its units
      const int r = q * 32 + lane;
      const int sy = r >> 3, sx = r & 7;
      const bool leader = (gt & 127) == 0;
      const int UC = P.ts_unit_ch;
      const int chunks = P.bn / UC;
      constexpr int NACC = NG * MT;
      const bool has_out = P.out != nullptr, has_mod = P.out_mod != nullptr;
      const int n_out = (has_out ? 1 : 0) + (has_mod ? 1 : 0);
      const bool one_in_flight = n_out == 1 && P.ts_slots == 2;
      const bool lrelu = P.act == W2E_ACT_LRELU;
      const float gain = lrelu ? 1.41421356237309515f : 1.f;
      const float slope = lrelu ? 0.2f : 1.f;
      const uint64_t slope2 = pack2(slope, slope);
      const bool has_noise = P.noise != nullptr, has_skip = RGB && P.rgb_skip != nullptr;
      const float nw = has_noise ? __ldg(P.noise_w) * gain : 0.f;
      const uint32_t stage0 =
          smem_u32(smem + P.ts_off) + (uint32_t)((group * 2 + half) * P.ts_slots * P.ts_unit_bytes);
      const uint32_t row_off = (uint32_t)(r * UC * 2);
      const uint32_t swz = (UC == 64) ? (uint32_t)(r & 7) : (uint32_t)((r >> 1) & 3);
      const uint32_t e_base = smem_u32(smem + P.e_off);
      const int bar_group = 1 + group * 3, bar_half = 2 + group * 3 + half;
      const int stride = split ? cta_slots : 2 * cta_slots;
      const int unit0 = RGB ? 0 : (split ? group * 2 + half : half), unit_step = RGB ? 1 : (split ? 4 : 2);
      const uint32_t nbuf_mask = (uint32_t)P.nbuf - 1u, nbuf_shift = P.nbuf == 4 ? 2u : (P.nbuf == 2 ? 1u : 0u);
      const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
      // skip-upsample polyphase taps of this thread's pixel (parity of (sy, sx): tile origins are even)
      const float cy0 = (sy & 1) ? P.kf[1] : P.kf[0], cy1 = (sy & 1) ? P.kf[3] : P.kf[2];
      const float cx0 = (sx & 1) ? P.kf[1] : P.kf[0], cx1 = (sx & 1) ? P.kf[3] : P.kf[2];
      // fused ToRGB: a thread finishes NP = MT/2 pixels per tile (row r of sub-tiles half*NP .. half*NP+NP-1) and
      // shares every per-channel constant it reads from shared memory between them
      constexpr int NP = RGB ? MT / 2 : 1;
      const uint32_t skip_off = (uint32_t)P.e_noise_bytes +
                                (uint32_t)(half * NP * P.skip_box_bytes + (((sy + 1) >> 1) * kSkipBoxW + ((sx + 1) >> 1) + 3) * 4);
      const uint32_t noise_off = (uint32_t)((sy * kTileW + sx) * 4);
      float rgbb[3] = {0.f, 0.f, 0.f};
      if (RGB && P.rgb_bias) {
#pragma unroll
        for (int o = 0; o < 3; ++o) rgbb[o] = __ldg(P.rgb_bias + o);
      }

      uint32_t un = 0, gen = 0;
      int prev_b = -1, prev_tn = -1;
      bool ok = true;
      uint32_t k = 0;
      for (int tile = cta_slot + (split ? 0 : group * cta_slots); tile < P.ntiles; tile += stride, ++k) {
        const uint32_t seq = split ? k : 2u * k + (uint32_t)group;   // position in the CTA's tile sequence
        const uint32_t es = seq % (uint32_t)kEStages;
        if (ok) ok = mbar_wait(&bars->e_full[es], (seq / (uint32_t)kEStages) & 1u, abort_flag);
        const bool dbg_on = P.dbg && blockIdx.x == 0 && seq < 64 && (gt & 255) == 0;
        if (dbg_on) P.dbg[seq * 8 + 5] = clock64();
        const uint32_t eb = e_base + es * (uint32_t)P.e_stage_bytes;
        int b, j0, i0, tn;
        lds_4i(eb + (uint32_t)P.e_info_off, b, j0, i0, tn);

        // per-channel constants of this (sample, channel block): staged when they change
        if (b != prev_b || tn != prev_tn) {
          prev_b = b; prev_tn = tn;
          gen ^= 1u;
          if (gt < P.bn) {
            float* cst = bars->ts_consts(group, (int)gen);
            // (pair mode: accumulator column gt = pixel (gt >> 5) of the pair, channel gt & 31 of the 32-channel arrays)
            const int nch = P.pair ? (P.Cout >> 1) : P.Cout;
            const int c = P.pair ? (gt & 31) : tn * P.bn + gt, bc = b * nch + c;
            cst[gt] = (P.out_scale ? __ldg(P.out_scale + bc) : 1.f) * gain;
            cst[128 + gt] = (P.bias ? __ldg(P.bias + c) : 0.f) * gain;
            cst[256 + gt] = P.next_scale ? __ldg(P.next_scale + bc) : 0.f;
            if (RGB) {
              const float rs = __ldg(P.rgb_style + bc);
#pragma unroll
              for (int o = 0; o < 3; ++o) cst[384 + o * 128 + gt] = __ldg(P.rgb_w + o * nch + c) * rs;
            }
          }
          named_bar_sync(bar_group, kT2EpiThreads);
        }
        const uint32_t sc_a = smem_u32(bars->ts_consts(group, (int)gen));

        if constexpr (!TR && MT == 2 && KSTEPS == 4 && WRES && RGB) {
          if (P.pair) {
            // ---------------------------------------------------------------- x-pair mode (32-channel RGB-only layer)
            // Accumulator row r = the pixel PAIR (j0 + 16*half + sy, 2*(i0 + sx) + a): columns [0,32) hold pixel a = 0,
            // [32,64) pixel a = 1 (the per-channel constants were staged from arrays expanded to 64 entries).  Both
            // pixels are finished here: noise, bias, leaky-ReLU, ToRGB dot product, skip upsample, image store.
            constexpr int kW = 16, kPlane = 10 * kW * 4, kRow = kW * 4;
            const int m = half;
            float nzp[2] = {0.f, 0.f};
            if (has_noise) {
#pragma unroll
              for (int a = 0; a < 2; ++a)
                nzp[a] = nw * lds_f32(eb + (uint32_t)((((m * kSubTileH + sy) * kW) + 2 * sx + a) * 4));
            }
            float init[2][3];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
              for (int o = 0; o < 3; ++o) init[a][o] = rgbb[o];
            if (has_skip) {
              // half-resolution column of pixel (.., 2*(i0+sx)+a) is i0 + sx + a - 1; the box starts at column i0 - 4
              const uint32_t sb0 = eb + (uint32_t)P.e_noise_bytes + (uint32_t)(m * P.skip_box_bytes) +
                                   (uint32_t)((((sy + 1) >> 1) * kW + sx + 3) * 4);
#pragma unroll
              for (int a = 0; a < 2; ++a) {
                const float ax0 = a ? P.kf[1] : P.kf[0], ax1 = a ? P.kf[3] : P.kf[2];
                const uint32_t sb = sb0 + (uint32_t)(a * 4);
#pragma unroll
                for (int o = 0; o < 3; ++o) {
                  const float t00 = lds_f32(sb + o * kPlane), t01 = lds_f32(sb + o * kPlane + 4);
                  const float t10 = lds_f32(sb + o * kPlane + kRow), t11 = lds_f32(sb + o * kPlane + kRow + 4);
                  init[a][o] += cy0 * fmaf(ax1, t01, ax0 * t00) + cy1 * fmaf(ax1, t11, ax0 * t10);
                }
              }
            }
            mbar_arrive(&bars->e_empty[es]);
            const uint32_t ci = seq & nbuf_mask;
            const uint32_t t_addr = t_lane + ci * (uint32_t)(NG * MT * P.bn) + (uint32_t)(m * P.bn);
            if (ok) ok = mbar_wait(&bars->acc_full[ci], (seq >> nbuf_shift) & 1u, abort_flag);
            tc_fence_after();
            if (dbg_on) P.dbg[seq * 8 + 6] = clock64();
            uint64_t racc[2][3];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
              for (int o = 0; o < 3; ++o) racc[a][o] = 0ull;
            // optional activation output (the training forward keeps it for the backward): the pair's 2 x 32 channels are
            // 128 contiguous bytes of the channels-last tensor, written straight from registers (16 bytes per store).  Two
            // copies of the loop (with / without the stores): the inference copy is the instruction sequence it was before.
            const int oy_p = j0 + m * kSubTileH + sy, ox_p = 2 * (i0 + sx);
            auto pair_body = [&](auto out_tag) {
              constexpr bool OUT = decltype(out_tag)::value;
              __nv_bfloat16* out_row = nullptr;
              if (OUT && oy_p < P.OH && ox_p < P.OW_real && ok)
                out_row = P.out + (((int64_t)b * P.OH + oy_p) * P.OW_real + ox_p) * 32;
#pragma unroll
              for (int a = 0; a < 2; ++a) {
                const uint64_t nz2 = pack2(nzp[a], nzp[a]);
#pragma unroll
                for (int c16 = 0; c16 < 32; c16 += 16) {
                  uint32_t v[16];
                  uint32_t po[OUT ? 8 : 1];
                  tmem_ld16(t_addr + (uint32_t)(a * 32 + c16), v);
                  tmem_ld_wait();
                  if (a == 1 && c16 == 16) {   // this thread's last TMEM read of the tile: hand the buffer back
                    tc_fence_before();
                    mbar_arrive(&bars->acc_empty[ci]);
                  }
                  const uint32_t ca = sc_a + (uint32_t)((a * 32 + c16) * 4);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    uint64_t a01, a23, b01, b23, w01[3], w23[3];
                    lds_2x2(ca + e * 16, a01, a23);
                    lds_2x2(ca + 512 + e * 16, b01, b23);
#pragma unroll
                    for (int o = 0; o < 3; ++o) lds_2x2(ca + (uint32_t)(1536 + o * 512) + e * 16, w01[o], w23[o]);
                    uint64_t f01 = add2(fma2(pack2u(v[4 * e], v[4 * e + 1]), a01, b01), nz2);
                    uint64_t f23 = add2(fma2(pack2u(v[4 * e + 2], v[4 * e + 3]), a23, b23), nz2);
                    const uint64_t g01 = mul2(f01, slope2), g23 = mul2(f23, slope2);
                    float x0, x1, x2, x3, y0, y1, y2, y3;
                    unpack2(f01, x0, x1); unpack2(f23, x2, x3);
                    unpack2(g01, y0, y1); unpack2(g23, y2, y3);
                    x0 = fmaxf(x0, y0); x1 = fmaxf(x1, y1); x2 = fmaxf(x2, y2); x3 = fmaxf(x3, y3);
                    f01 = pack2(x0, x1);
                    f23 = pack2(x2, x3);
                    if (OUT) {
                      po[2 * e] = cvt_bf16x2(x0, x1);
                      po[2 * e + 1] = cvt_bf16x2(x2, x3);
                    }
#pragma unroll
                    for (int o = 0; o < 3; ++o) racc[a][o] = fma2(f23, w23[o], fma2(f01, w01[o], racc[a][o]));
                  }
                  if (OUT && out_row) {
                    uint4* dst = reinterpret_cast<uint4*>(out_row + a * 32 + c16);
                    dst[0] = make_uint4(po[0], po[1], po[2], po[3]);
                    dst[1] = make_uint4(po[OUT ? 4 : 0], po[OUT ? 5 : 0], po[OUT ? 6 : 0], po[OUT ? 7 : 0]);
                  }
                }
              }
            };
            if (P.out) pair_body(std::true_type{}); else pair_body(std::false_type{});
            const int oy = oy_p, ox = ox_p;
            if (oy < P.OH && ox < P.OW_real && ok) {
              const int64_t plane = (int64_t)P.OH * P.OW_real;
              const int64_t di = ((int64_t)b * 3 * P.OH + oy) * P.OW_real + ox;
#pragma unroll
              for (int o = 0; o < 3; ++o) {
                float l0, h0, l1, h1;
                unpack2(racc[0][o], l0, h0);
                unpack2(racc[1][o], l1, h1);
                const float v0 = init[0][o] + (l0 + h0), v1 = init[1][o] + (l1 + h1);
                if (P.rgb_bf16 == 2) {
                  *reinterpret_cast<uchar2*>(reinterpret_cast<uint8_t*>(P.rgb) + di + o * plane) =
                      make_uchar2(quant_u8(v0), quant_u8(v1));
                } else if (P.rgb_bf16) {
                  *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(P.rgb) + di + o * plane) =
                      __floats2bfloat162_rn(v0, v1);
                } else {
                  *reinterpret_cast<float2*>(reinterpret_cast<float*>(P.rgb) + di + o * plane) = make_float2(v0, v1);
                }
              }
            }
            if (dbg_on) P.dbg[seq * 8 + 7] = clock64();
            continue;
          }
        }

        // epilogue inputs staged by the producer's TMA boxes
        // noise of this thread's pixel(s): fused ToRGB -> sub-tiles half*NP + p; otherwise one value per sub-tile
        constexpr int NZ = RGB ? NP : (MT < 2 ? 1 : 2);
        float nzm[NZ];
#pragma unroll
        for (int m = 0; m < NZ; ++m) nzm[m] = 0.f;
        uint64_t racc[NP][3];
        float rgb_init[NP][3];
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
          for (int o = 0; o < 3; ++o) { racc[p][o] = 0ull; rgb_init[p][o] = tn == 0 ? rgbb[o] : 0.f; }
        if (has_noise) {
#pragma unroll
          for (int m = 0; m < NZ; ++m)
            nzm[m] = nw * lds_f32(eb + noise_off + (uint32_t)(((RGB ? half * NP : 0) + m) * kSubTileH * kTileW * 4));
        }
        if (has_skip && tn == 0) {   // (several channel blocks: the first one adds bias + skip, see the final store)
          // upfirdn2d(skip, up=2, pad=(2,1)) = a 2x2-tap polyphase filter on rows ya, ya+1 / columns xa, xa+1.
          // The box starts one row and FOUR columns before the sub-tile's first source pixel: the innermost
          // start coordinate of a TMA box must be 16-byte aligned (x0 - 1 faults with an illegal instruction).
          constexpr int kPlane = 10 * kSkipBoxW * 4, kRow = kSkipBoxW * 4;
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const uint32_t sb = eb + skip_off + (uint32_t)(p * P.skip_box_bytes);
#pragma unroll
            for (int o = 0; o < 3; ++o) {
              const float t00 = lds_f32(sb + o * kPlane), t01 = lds_f32(sb + o * kPlane + 4);
              const float t10 = lds_f32(sb + o * kPlane + kRow), t11 = lds_f32(sb + o * kPlane + kRow + 4);
              rgb_init[p][o] += cy0 * fmaf(cx1, t01, cx0 * t00) + cy1 * fmaf(cx1, t11, cx0 * t10);
            }
          }
        }
        mbar_arrive(&bars->e_empty[es]);

        // accumulator buffer of this tile: with 4 buffers a group alternates between two of them, so the MMAs
        // of its next tile run while it still drains this one
        const uint32_t ci = seq & nbuf_mask;
        const uint32_t t_tile = t_lane + ci * (uint32_t)(NACC * P.bn);
        const bool percls = TR && P.percls;
        if (!percls) {
          if (ok) ok = mbar_wait(&bars->acc_full[ci], (seq >> nbuf_shift) & 1u, abort_flag);
          tc_fence_after();
        }
        if (dbg_on) P.dbg[seq * 8 + 6] = clock64();
        const int nunits = RGB ? chunks : NACC * chunks;
        bool released = percls;   // (per-class hand-over releases class by class below)
        auto run_units = [&](auto out_tag, auto mod_tag, const int u_begin, const int u_end) {
          constexpr bool OUT = decltype(out_tag)::value, MOD = decltype(mod_tag)::value;
          for (int ui = u_begin + unit0; ui < u_end; ui += unit_step) {
            const int acc = RGB ? half * NP : ui / chunks;     // (first) accumulator of this unit
            const int chunk = RGB ? ui : ui - acc * chunks;
            const int m = acc % MT;
            const uint32_t t_addr = t_tile + (uint32_t)(acc * P.bn + chunk * UC);
            uint64_t nz2[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              const float nzv = RGB ? nzm[p] : ((NZ == 2 && m == 1) ? nzm[NZ - 1] : nzm[0]);
              nz2[p] = pack2(nzv, nzv);
            }
            const uint32_t slot_o =
                stage0 + (uint32_t)(((OUT && MOD) ? 0 : (int)(un & (uint32_t)(P.ts_slots - 1))) * P.ts_unit_bytes);
            const uint32_t slot_m = OUT ? stage0 + (uint32_t)P.ts_unit_bytes : slot_o;
            const bool last_unit = !percls && ui + unit_step >= u_end;
            for (int c16 = 0; c16 < UC; c16 += 16) {
              uint32_t v[NP][16];
#pragma unroll
              for (int p = 0; p < NP; ++p) tmem_ld16(t_addr + (uint32_t)(p * P.bn + c16), v[p]);
              tmem_ld_wait();
              if (last_unit && c16 + 16 >= UC) {   // this thread's last TMEM read of the tile: hand the buffer back
                tc_fence_before();
                mbar_arrive(&bars->acc_empty[ci]);
                released = true;
              }
              const uint32_t ca = sc_a + (uint32_t)((chunk * UC + c16) * 4);
              uint32_t po[OUT ? 8 : 1], pm[MOD ? 8 : 1];   // staged outputs: pixel 0 only (NP == 1 whenever OUT || MOD)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                uint64_t a01, a23, b01 = 0, b23 = 0, w01[3], w23[3];
                lds_2x2(ca + e * 16, a01, a23);
                if (!TR) lds_2x2(ca + 512 + e * 16, b01, b23);
                if (RGB) {
#pragma unroll
                  for (int o = 0; o < 3; ++o) lds_2x2(ca + (uint32_t)(1536 + o * 512) + e * 16, w01[o], w23[o]);
                }
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                  uint64_t f01, f23;
                  if (TR) {
                    f01 = mul2(pack2u(v[p][4 * e], v[p][4 * e + 1]), a01);
                    f23 = mul2(pack2u(v[p][4 * e + 2], v[p][4 * e + 3]), a23);
                  } else {
                    f01 = add2(fma2(pack2u(v[p][4 * e], v[p][4 * e + 1]), a01, b01), nz2[p]);
                    f23 = add2(fma2(pack2u(v[p][4 * e + 2], v[p][4 * e + 3]), a23, b23), nz2[p]);
                    const uint64_t g01 = mul2(f01, slope2), g23 = mul2(f23, slope2);
                    float x0, x1, x2, x3, y0, y1, y2, y3;
                    unpack2(f01, x0, x1); unpack2(f23, x2, x3);
                    unpack2(g01, y0, y1); unpack2(g23, y2, y3);
                    f01 = pack2(fmaxf(x0, y0), fmaxf(x1, y1));
                    f23 = pack2(fmaxf(x2, y2), fmaxf(x3, y3));
                  }
                  if (RGB) {
#pragma unroll
                    for (int o = 0; o < 3; ++o) racc[p][o] = fma2(f23, w23[o], fma2(f01, w01[o], racc[p][o]));
                  }
                  if (OUT && p == 0) {
                    float x0, x1, x2, x3;
                    unpack2(f01, x0, x1); unpack2(f23, x2, x3);
                    po[2 * e] = cvt_bf16x2(x0, x1);
                    po[2 * e + 1] = cvt_bf16x2(x2, x3);
                  }
                  if (MOD && p == 0) {
                    uint64_t n01, n23;
                    lds_2x2(ca + 1024 + e * 16, n01, n23);
                    float x0, x1, x2, x3;
                    unpack2(mul2(f01, n01), x0, x1); unpack2(mul2(f23, n23), x2, x3);
                    pm[2 * e] = cvt_bf16x2(x0, x1);
                    pm[2 * e + 1] = cvt_bf16x2(x2, x3);
                  }
                }
              }
              if (OUT || MOD) {
                if (c16 == 0) {
                  // slot acquire: the TMA store that last read these slots must have drained
                  if (leader) {
                    if (one_in_flight) bulk_wait_read<1>(); else bulk_wait_read<0>();
                  }
                  named_bar_sync(bar_half, 128);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const uint32_t col = (((uint32_t)(c16 >> 3) + j) ^ swz) << 4;
                  if (OUT) sts_128(slot_o + row_off + col, po[4 * j], po[4 * j + 1], po[4 * j + 2], po[4 * j + 3]);
                  if (MOD) sts_128(slot_m + row_off + col, pm[4 * j], pm[4 * j + 1], pm[4 * j + 2], pm[4 * j + 3]);
                }
              }
            }
            if (OUT || MOD) {
              fence_proxy_async_smem();
              named_bar_sync(bar_half, 128);
              if (leader) {
                const int cc = tn * P.bn + chunk * UC, yy = j0 + m * kSubTileH;
                if (TR) {
                  tma_store_4d(&M.st[acc / MT], slot_o, cc, i0, yy, b);
                } else {
                  if (OUT) {
                    if (P.reduce_add) tma_reduce_add_4d(&M.st[0], slot_o, cc, i0, yy, b);
                    else tma_store_4d(&M.st[0], slot_o, cc, i0, yy, b);
                  }
                  if (MOD) tma_store_4d(&M.st[1], slot_m, cc, i0, yy, b);
                }
                bulk_commit();
              }
              ++un;
            }
          }
        };
        using T1 = std::true_type;
        using T0 = std::false_type;
        if (percls) {
          // classes in the order the issuer completes them: 2, 3 (tap row 1), then 0, 1
          const int per_class = MT * chunks;
#pragma unroll 1
          for (int q = 0; q < 4; ++q) {
            const int cls = (q + 2) & 3;
            if (ok) ok = mbar_wait(&bars->acc_full[cls], seq & 1u, abort_flag);
            tc_fence_after();
            run_units(T1{}, T0{}, cls * per_class, (cls + 1) * per_class);
            tc_fence_before();
            mbar_arrive(&bars->acc_empty[cls]);
          }
        } else if (TR) {
          run_units(T1{}, T0{}, 0, nunits);
        } else if (has_out) {
          if (has_mod) run_units(T1{}, T1{}, 0, nunits); else run_units(T1{}, T0{}, 0, nunits);
        } else {
          if (has_mod) run_units(T0{}, T1{}, 0, nunits);
          else if (RGB) run_units(T0{}, T0{}, 0, nunits);
        }
        if (!released) {
          tc_fence_before();
          mbar_arrive(&bars->acc_empty[ci]);
        }
        if (RGB) {
          const int64_t plane = (int64_t)P.OH * P.OW;
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const int oy = j0 + (half * NP + p) * kSubTileH + sy, ox = i0 + sx;
            if (oy < P.OH && ox < P.OW && ok) {
              const int64_t di = ((int64_t)b * 3 * P.OH + oy) * P.OW + ox;
#pragma unroll
              for (int o = 0; o < 3; ++o) {
                float lo, hi;
                unpack2(racc[p][o], lo, hi);
                const float v = rgb_init[p][o] + (lo + hi);
                // two channel blocks (Cout = 256 as two 128-column tiles): both ADD their partial sums to the zero-
                // initialised image; two addends commute, so the result does not depend on the arrival order
                if (P.tiles_n > 1) atomicAdd(reinterpret_cast<float*>(P.rgb) + di + o * plane, v);
                else if (P.rgb_bf16 == 2) reinterpret_cast<uint8_t*>(P.rgb)[di + o * plane] = quant_u8(v);
                else if (P.rgb_bf16) reinterpret_cast<__nv_bfloat16*>(P.rgb)[di + o * plane] = __float2bfloat16_rn(v);
                else reinterpret_cast<float*>(P.rgb)[di + o * plane] = v;
              }
            }
          }
        }
        if (dbg_on) P.dbg[seq * 8 + 7] = clock64();
      }
      if (leader) bulk_wait_all();
    }
  } else {
    // ------------------------------------------------------------------ epilogue: warps 0..7
    // warp w reads TMEM lane quarter w%4 (hardware rule); warps 0..3 take the first half of the
    // tile's BN columns, warps 4..7 the second half.  With the fused ToRGB (RGB) a thread needs its
    // pixel's whole channel row, so warps 0..3 take sub-tile 0 and warps 4..7 sub-tile 1 instead.
    const int q = warp & 3;
    const int half = warp >> 2;
    const int et = threadIdx.x;  // 0..255; also the channel whose constants this thread stages
    const int r = q * 32 + lane;      // accumulator row == pixel inside the sub-tile
    const int sy = r >> 3, sx = r & 7;
    const bool lrelu = P.act == W2E_ACT_LRELU;
    const float gain = lrelu ? 1.41421356237309515f : 1.f;
    const float nw = P.noise ? __ldg(P.noise_w) * gain : 0.f;
    const int c_begin = RGB ? 0 : ((P.bn >= 32) ? half * (P.bn >> 1) : 0);
    const int c_end = RGB ? P.bn : ((P.bn >= 32) ? c_begin + (P.bn >> 1) : (half == 0 ? P.bn : 0));
    constexpr int out_stride = TR ? 2 : 1;
    constexpr int NPIX = RGB ? 1 : NG * MT;   // pixels this thread finishes per tile

    // Everything a tile's epilogue needs from global memory, fetched ONE TILE AHEAD into registers
    // so that its latency hides behind the current tile's arithmetic.
    struct Pre {                          // RAW loaded values only: nothing here may be consumed
      int b, j0, i0, co0;                 // arithmetically until the tile becomes current
      float scale, shift, next, rgbs, rgbw[3];
      float nz[NPIX];
      float rgbb[3], tap[3][4];           // ToRGB bias and the 2x2 skip taps of this thread's pixel
    };
    auto pixel_of = [&](const Pre& t, int gm, int& oy, int& ox) -> bool {
      const int g = gm / MT, m = RGB ? half : gm % MT;
      const int j = t.j0 + m * kSubTileH + sy, i = t.i0 + sx;
      oy = j * out_stride + (TR ? (g >> 1) : 0);
      ox = i * out_stride + (TR ? (g & 1) : 0);
      return j < P.grid_h && i < P.grid_w && oy < P.OH && ox < P.OW;
    };
    auto prefetch = [&](int tile) -> Pre {
      Pre t;
      const int tn = tile / tiles_per_n;
      int rem = tile - tn * tiles_per_n;
      t.b = rem / tiles_xy;   // (pair-local sample index in cluster mode, mapped below)
      rem -= t.b * tiles_xy;
      t.b = (t.b << cl) + (int)crank;
      const int ty = rem / P.tiles_x, tx = rem - ty * P.tiles_x;
      t.j0 = ty * (kSubTileH * MT); t.i0 = tx * kTileW; t.co0 = tn * P.bn;
      t.scale = 1.f; t.shift = 0.f; t.next = 0.f; t.rgbs = 0.f; t.rgbw[0] = t.rgbw[1] = t.rgbw[2] = 0.f;
      if (et < P.bn) {
        const int64_t bc = (int64_t)t.b * P.Cout + t.co0 + et;
        if (P.out_scale) t.scale = __ldg(P.out_scale + bc);
        if (P.bias) t.shift = __ldg(P.bias + t.co0 + et);
        if (P.next_scale) t.next = __ldg(P.next_scale + bc);
        if (RGB) {
          t.rgbs = __ldg(P.rgb_style + bc);
#pragma unroll
          for (int o = 0; o < 3; ++o) t.rgbw[o] = __ldg(P.rgb_w + o * P.Cout + t.co0 + et);
        }
      }
#pragma unroll
      for (int gm = 0; gm < NPIX; ++gm) {
        int oy, ox;
        const bool valid = pixel_of(t, gm, oy, ox);
        t.nz[gm] = (valid && P.noise)
            ? __ldg(P.noise + (P.noise_per_sample ? (int64_t)t.b * P.OH * P.OW : 0) + (int64_t)oy * P.OW + ox)
            : 0.f;
      }
      if (RGB) {
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          t.rgbb[o] = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) t.tap[o][k] = 0.f;
        }
        int oy, ox;
        if (pixel_of(t, 0, oy, ox)) {
#pragma unroll
          for (int o = 0; o < 3; ++o) t.rgbb[o] = P.rgb_bias ? __ldg(P.rgb_bias + o) : 0.f;
          if (P.rgb_skip) {
            // upfirdn2d(skip, up=2, pad=(2,1)) is a 2x2-tap polyphase filter: rows ya, ya+1, columns xa, xa+1
            const int h = P.OH >> 1, wd = P.OW >> 1;
            const int ya = ((oy + 1) >> 1) - 1, xa = ((ox + 1) >> 1) - 1;
#pragma unroll
            for (int o = 0; o < 3; ++o) {
              const float* sp = P.rgb_skip + ((int64_t)t.b * 3 + o) * h * wd;
#pragma unroll
              for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                  const int iy = ya + dy, ix = xa + dx;
                  if (iy >= 0 && iy < h && ix >= 0 && ix < wd) t.tap[o][dy * 2 + dx] = __ldg(sp + (int64_t)iy * wd + ix);
                }
            }
          }
        }
      }
      return t;
    };

    Ring cr;
    uint32_t parity = 0;
    bool ok = true;
    int tile = cta_slot;
    Pre cur;
    if (tile < P.ntiles) cur = prefetch(tile);
    for (; tile < P.ntiles; tile += cta_slots) {
      // publish this tile's per-channel constants (double-buffered by tile parity: the single named
      // barrier per tile also orders the reuse of the other buffer)
      const int cb = parity;
      parity ^= 1u;
      if (et < P.bn) {
        bars->ep_scale[cb][et] = cur.scale * gain;
        bars->ep_shift[cb][et] = cur.shift * gain;
        bars->ep_next[cb][et] = cur.next;
        if (RGB) {
#pragma unroll
          for (int o = 0; o < 3; ++o) bars->ep_rgb[cb][o][et] = cur.rgbw[o] * cur.rgbs;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kT2EpiThreads) : "memory");
      const Pre me = cur;
      if (tile + cta_slots < P.ntiles) cur = prefetch(tile + cta_slots);  // in flight during this tile
      if (ok) ok = mbar_wait(&bars->acc_full[cr.idx], cr.phase, abort_flag);
      tc_fence_after();
      const float* sc = bars->ep_scale[cb];
      const float* sh = bars->ep_shift[cb];
      const float* nx = bars->ep_next[cb];
      const uint32_t t_tile = tmem_base + ((uint32_t)(q * 32) << 16) + cr.idx * (uint32_t)(NG * MT) * (uint32_t)P.bn;
      float rgb_acc[3] = {0.f, 0.f, 0.f};
      if (RGB) {
        int oy, ox;
        pixel_of(me, 0, oy, ox);
        const float cy0 = (oy & 1) ? P.kf[1] : P.kf[0], cy1 = (oy & 1) ? P.kf[3] : P.kf[2];
        const float cx0 = (ox & 1) ? P.kf[1] : P.kf[0], cx1 = (ox & 1) ? P.kf[3] : P.kf[2];
#pragma unroll
        for (int o = 0; o < 3; ++o)
          rgb_acc[o] = me.rgbb[o] + cy0 * fmaf(cx1, me.tap[o][1], cx0 * me.tap[o][0]) +
                       cy1 * fmaf(cx1, me.tap[o][3], cx0 * me.tap[o][2]);
        // several channel blocks (Cout > 256): every block adds its partial sum to the zero-initialised image, the
        // first one also bias + skip.  Two addends commute, so the result does not depend on the arrival order.
        if (me.co0 != 0) rgb_acc[0] = rgb_acc[1] = rgb_acc[2] = 0.f;
      }
#pragma unroll
      for (int gmi = 0; gmi < NPIX; ++gmi) {
        const int gm = RGB ? half : gmi;   // accumulator index inside the tile
        int oy, ox;
        const bool valid = pixel_of(me, gmi, oy, ox) && ok;
        const int64_t pix = valid ? ((int64_t)me.b * P.OH + oy) * P.OW + ox : 0;
        const float nzv = nw * me.nz[gmi];
        const uint32_t t_addr = t_tile + (uint32_t)gm * (uint32_t)P.bn;
        __nv_bfloat16* o_row = (P.out && !TF32) ? P.out + pix * P.Cout + me.co0 : nullptr;
        __nv_bfloat16* m_row = (P.out_mod && !TF32) ? P.out_mod + pix * P.Cout + me.co0 : nullptr;
        // tf32 mode: the same tensors hold fp32
        float* o_row32 = (P.out && TF32) ? reinterpret_cast<float*>(P.out) + pix * P.Cout + me.co0 : nullptr;
        float* m_row32 = (P.out_mod && TF32) ? reinterpret_cast<float*>(P.out_mod) + pix * P.Cout + me.co0 : nullptr;
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 16) {
          uint32_t v[16];
          tmem_ld16(t_addr + (uint32_t)c, v);
          tmem_ld_wait();
          if (valid) {
            float f[16];
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 a4 = *reinterpret_cast<const float4*>(sc + c + 4 * e4);
              const float4 b4 = *reinterpret_cast<const float4*>(sh + c + 4 * e4);
              // (acc * scale + bias) + noise, the operation order of the staged epilogue: the two are bit-identical
              f[4 * e4 + 0] = fmaf(__uint_as_float(v[4 * e4 + 0]), a4.x, b4.x) + nzv;
              f[4 * e4 + 1] = fmaf(__uint_as_float(v[4 * e4 + 1]), a4.y, b4.y) + nzv;
              f[4 * e4 + 2] = fmaf(__uint_as_float(v[4 * e4 + 2]), a4.z, b4.z) + nzv;
              f[4 * e4 + 3] = fmaf(__uint_as_float(v[4 * e4 + 3]), a4.w, b4.w) + nzv;
            }
            if (lrelu) {
#pragma unroll
              for (int e = 0; e < 16; ++e) f[e] = fmaxf(f[e], 0.2f * f[e]);
            }
            if (RGB) {
              const float* w0 = bars->ep_rgb[cb][0] + c;
              const float* w1 = bars->ep_rgb[cb][1] + c;
              const float* w2 = bars->ep_rgb[cb][2] + c;
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const float4 x0 = *reinterpret_cast<const float4*>(w0 + 4 * e4);
                const float4 x1 = *reinterpret_cast<const float4*>(w1 + 4 * e4);
                const float4 x2 = *reinterpret_cast<const float4*>(w2 + 4 * e4);
                rgb_acc[0] = fmaf(f[4 * e4 + 3], x0.w, fmaf(f[4 * e4 + 2], x0.z, fmaf(f[4 * e4 + 1], x0.y, fmaf(f[4 * e4], x0.x, rgb_acc[0]))));
                rgb_acc[1] = fmaf(f[4 * e4 + 3], x1.w, fmaf(f[4 * e4 + 2], x1.z, fmaf(f[4 * e4 + 1], x1.y, fmaf(f[4 * e4], x1.x, rgb_acc[1]))));
                rgb_acc[2] = fmaf(f[4 * e4 + 3], x2.w, fmaf(f[4 * e4 + 2], x2.z, fmaf(f[4 * e4 + 1], x2.y, fmaf(f[4 * e4], x2.x, rgb_acc[2]))));
              }
            }
            if (o_row) {
              uint4 pk[2];
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
              for (int e = 0; e < 8; ++e) h[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
              uint4* dst = reinterpret_cast<uint4*>(o_row + c);
              dst[0] = pk[0];
              dst[1] = pk[1];
            }
            if (m_row) {
              uint4 pk[2];
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const float4 n4 = *reinterpret_cast<const float4*>(nx + c + 4 * e4);
                h[2 * e4] = __floats2bfloat162_rn(f[4 * e4] * n4.x, f[4 * e4 + 1] * n4.y);
                h[2 * e4 + 1] = __floats2bfloat162_rn(f[4 * e4 + 2] * n4.z, f[4 * e4 + 3] * n4.w);
              }
              uint4* dst = reinterpret_cast<uint4*>(m_row + c);
              dst[0] = pk[0];
              dst[1] = pk[1];
            }
            if constexpr (TF32) {
              if (o_row32) {
                float4* dst = reinterpret_cast<float4*>(o_row32 + c);
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) dst[e4] = make_float4(f[4 * e4], f[4 * e4 + 1], f[4 * e4 + 2], f[4 * e4 + 3]);
              }
              if (m_row32) {
                float4* dst = reinterpret_cast<float4*>(m_row32 + c);
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                  const float4 n4 = *reinterpret_cast<const float4*>(nx + c + 4 * e4);
                  dst[e4] = make_float4(f[4 * e4] * n4.x, f[4 * e4 + 1] * n4.y, f[4 * e4 + 2] * n4.z, f[4 * e4 + 3] * n4.w);
                }
              }
            }
          }
        }
        if (RGB && valid) {
#pragma unroll
          for (int o = 0; o < 3; ++o) {
            const int64_t di = (((int64_t)me.b * 3 + o) * P.OH + oy) * P.OW + ox;
            if (P.rgb_bf16 == 2) reinterpret_cast<uint8_t*>(P.rgb)[di] = quant_u8(rgb_acc[o]);   // (tiles_n == 1)
            else if (P.rgb_bf16) reinterpret_cast<__nv_bfloat16*>(P.rgb)[di] = __float2bfloat16_rn(rgb_acc[o]);
            else if (P.tiles_n > 1) atomicAdd(reinterpret_cast<float*>(P.rgb) + di, rgb_acc[o]);
            else reinterpret_cast<float*>(P.rgb)[di] = rgb_acc[o];
          }
        }
      }
      // accumulator buffer drained: hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(&bars->acc_empty[cr.idx]);
      cr.advance(P.nbuf);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (cl) cluster_sync();   // no CTA exits while its peer may still multicast into it
  if (threadIdx.x == 0 && bars->abort_flag && P.error_flag) *P.error_flag = 1;
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  }
}

template <bool TR, int MT, int KSTEPS, bool WRES, bool RGB, bool TS, bool TF32 = false>
static int launch_tc2(const CUtensorMap& ma, const CUtensorMap& mb, const Tc2Params& P, const Tc2Maps& M, int smem_bytes,
                      int max_ctas, cudaStream_t s) {
  auto kern = modconv_tc2_kernel<TR, MT, KSTEPS, WRES, RGB, TS, TF32>;
  static PerDeviceOnce configured;
  if (!configured.done()) {
    W2E_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured.mark();
  }
  int per_sm = 1;
  constexpr int kThreads = TS ? (WRES ? kT2ThreadsTS2 : kT2ThreadsTS) : kT2Threads;
  W2E_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem_bytes));
  if (per_sm * P.tmem_cols > 512) per_sm = 512 / P.tmem_cols;
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int ctas = sm_count() * per_sm;
  if (max_ctas > 0 && ctas > max_ctas) ctas = max_ctas;
  if (P.cluster) {
    // clusters of 2^P.cluster CTAs (one per SM); P.ntiles counts tile GROUPS (same tile of consecutive samples)
    const int csize = 1 << P.cluster;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3((unsigned)kThreads);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)csize; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)(sm_count() / csize * csize));
    static int max_clusters[4] = {0, 0, 0, 0};   // per cluster size (this kernel instantiation, this smem plan class)
    int nclusters = 0;
    W2E_CUDA_OK(cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg));
    (void)max_clusters;
    if (nclusters < 1) return set_error(W2E_ERR_UNSUPPORTED, "modconv_tc2: no %d-CTA cluster fits", csize);
    if (max_ctas > 0 && nclusters * csize > max_ctas) nclusters = max_ctas / csize > 0 ? max_ctas / csize : 1;
    if (nclusters > P.ntiles) nclusters = P.ntiles;
    cfg.gridDim = dim3((unsigned)(nclusters * csize));
    W2E_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ma, mb, P, M));
    return W2E_OK;
  }
  if (ctas > P.ntiles) ctas = P.ntiles;
  kern<<<ctas, kThreads, smem_bytes, s>>>(ma, mb, P, M);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

}  // namespace w2e

using namespace w2e;

struct RgbArgs {
  const float* w; const float* style; const float* bias; const float* skip; const float* host_taps1d; void* rgb; int rgb_dtype;
  int pair;   // x-pair mode: the tensors describe pixel PAIRS (Cin = Cout = 64, in_w = image width / 2), see w2e_modconv_tc2_rgb_pair
};

struct FbArgs {   // fused up-convolution + Blur: the separable 4x4 FIR (flipped taps)
  float fv[4], fh[4];
};

struct ViewArgs {   // plain conv on a strided view of a channels-last tensor, with a subset of the 9 filter taps
  int64_t stride_x, stride_y, stride_b;   // in elements: between pixels of a row, rows, samples
  int tap_mask;
  int out_h, out_w;   // extent of the `out` buffer ([B,out_h,out_w,Cout], <= the convolution grid: the rest is clipped)
  int reduce;         // 1: add the result to `out` (TMA reduce) instead of storing it
  int dg4;            // fused dgrad of the transposed convolution: xs = the (2 h + 1) x (2 w + 1) upstream gradient, read through
                      // its four parity-class views; the strides above are not used.  1 = four class tiles per A stage, weights
                      // resident (in_h x in_w = h x w); 2 = K-loop form, weights [9][Cout][4 K] in the ring, Cin = 4 K,
                      // in_h x in_w = (h + 1) x (w + 1) = the grid of class (0,0), out_h x out_w = h x w
  long long dg_masks; // dg4 == 2: the four 9-bit tap masks, class c = py * 2 + px at bit 9 c
  int pair;           // plain 32 -> 32 channel convolution on PIXEL PAIRS (w2e_modconv_tc2_pair): tensors described as 64-channel
                      // pairs, only the K-step skipping of the issue loop differs from an ordinary 64 -> 64 layer
};

static int run_tc2(const void* xs, const void* w, const float* out_scale, const float* bias, const float* noise,
                   const float* noise_w, int noise_batch, const float* next_scale, void* out, void* out_mod,
                   int* error_flag, int B, int Cin, int Cout, int in_h, int in_w, int transposed, int act,
                   const RgbArgs* rgb, const w2e_tc2_config* cfg, void* stream, bool allow_mt4 = true,
                   const FbArgs* fb = nullptr, bool tf32 = false, const ViewArgs* view = nullptr) {
  // per-call tuning / A-B switches (include/w2e.h: w2e_tc2_config); NULL = defaults
  const int g_max_ctas = cfg ? cfg->max_ctas : 0;
  const int g_ts_mode = cfg ? cfg->ts_mode : 1;
  const int g_flags = cfg ? (cfg->flags & 1019) : 0;
  const int g_cluster_mode = cfg ? (cfg->cluster_log2 < 0 ? 0 : (cfg->cluster_log2 > 3 ? 3 : cfg->cluster_log2)) : 0;
  long long* const g_dbg = cfg ? (long long*)cfg->timeline : nullptr;
  W2E_CHECK_ARG(xs && w && (out || out_mod || rgb), "modconv_tc2: null pointer");
  W2E_CHECK_ARG(out_mod == nullptr || next_scale != nullptr, "modconv_tc2: out_mod needs next_scale");
  W2E_CHECK_ARG(B > 0 && in_h > 0 && in_w > 0, "modconv_tc2: bad shape");
  W2E_CHECK_ARG(Cin % (tf32 ? 16 : 32) == 0 && Cout % 16 == 0, "modconv_tc2: needs Cin %% %d == 0 and Cout %% 16 == 0 (got %d, %d)",
                tf32 ? 16 : 32, Cin, Cout);
  W2E_CHECK_ARG(!tf32 || (!rgb && !fb), "modconv_tc2_tf32: no fused ToRGB / blur in tf32 mode");
  W2E_CHECK_ARG(noise == nullptr || (noise_w != nullptr && (noise_batch == 1 || noise_batch == B)), "modconv_tc2: noise");
  W2E_CHECK_ARG(((uintptr_t)xs & 15) == 0 && ((uintptr_t)w & 15) == 0, "modconv_tc2: operands must be 16-byte aligned");
  if (!w2e_modconv_tc_supported()) return set_error(W2E_ERR_UNSUPPORTED, "modconv_tc2: device is not sm_100");

  Tc2Params P;
  memset(&P, 0, sizeof(P));
  P.out_scale = out_scale; P.bias = bias; P.noise = noise; P.noise_w = noise_w; P.next_scale = next_scale;
  P.out = (__nv_bfloat16*)out; P.out_mod = (__nv_bfloat16*)out_mod; P.error_flag = error_flag;
  P.dbg = g_dbg;
  P.flags = g_flags;
  P.noise_per_sample = (noise && noise_batch != 1) ? 1 : 0;
  P.B = B; P.Cin = Cin; P.Cout = Cout; P.act = act;
  if (rgb) {
    W2E_CHECK_ARG(rgb->w && rgb->style && rgb->rgb, "modconv_tc2_rgb: null pointer");
    W2E_CHECK_ARG(!transposed && Cout <= 512 && in_h > kSubTileH,
                  "modconv_tc2_rgb: the fused ToRGB needs a plain conv with Cout <= 512 and more than %d rows", kSubTileH);
    W2E_CHECK_ARG(rgb->skip == nullptr || (rgb->host_taps1d && in_h % 2 == 0 && in_w % 2 == 0),
                  "modconv_tc2_rgb: skip needs taps and even H, W");
    W2E_CHECK_ARG(rgb->rgb_dtype == W2E_F32 || ((rgb->rgb_dtype == W2E_BF16 || rgb->rgb_dtype == W2E_U8) && Cout < 256),
                  "modconv_tc2_rgb: rgb_dtype must be W2E_F32, or W2E_BF16 / W2E_U8 with Cout < 256");
    P.rgb_w = rgb->w; P.rgb_style = rgb->style; P.rgb_bias = rgb->bias; P.rgb_skip = rgb->skip; P.rgb = rgb->rgb;
    P.rgb_bf16 = rgb->rgb_dtype == W2E_BF16 ? 1 : (rgb->rgb_dtype == W2E_U8 ? 2 : 0);
    if (rgb->skip)
      for (int i = 0; i < 4; ++i) P.kf[i] = rgb->host_taps1d[3 - i];
  }
  P.pitch = kPitch;
  P.ntaps = 9;
  P.tap_mask = 0x1ff;
  const bool dg4 = view && view->dg4 == 1;
  const bool dgk = view && view->dg4 == 2;
  if (dg4) {
    W2E_CHECK_ARG(!transposed && !rgb && !fb && !tf32, "modconv_tc2_dgrad_up: plain bf16 convolution only");
    P.dg4 = 1;
  } else if (dgk) {
    W2E_CHECK_ARG(!transposed && !rgb && !fb && !tf32 && Cin % 256 == 0, "modconv_tc2_dgrad_up: K-loop form needs 4 x 64 k channels");
    P.dgk = Cin / 4 / 64;
    P.dg_masks = view->dg_masks;
    P.tap_mask = (int)(view->dg_masks & 0x1ff);   // class (0,0): its first used tap overwrites the accumulator
    W2E_CHECK_ARG(P.tap_mask != 0, "modconv_tc2_dgrad_up: empty tap mask of class (0,0)");
  } else if (view && !view->pair) {
    W2E_CHECK_ARG(!transposed && !rgb && !fb && !tf32, "modconv_tc2_view: plain bf16 convolution only");
    W2E_CHECK_ARG(view->tap_mask > 0 && view->tap_mask <= 0x1ff, "modconv_tc2_view: tap mask %d", view->tap_mask);
    W2E_CHECK_ARG(view->stride_x > 0 && (view->stride_x * 2) % 16 == 0 && (view->stride_y * 2) % 16 == 0 &&
                      (view->stride_b * 2) % 16 == 0, "modconv_tc2_view: strides must be multiples of 8 elements");
    W2E_CHECK_ARG(view->out_h > 0 && view->out_h <= in_h && view->out_w > 0 && view->out_w <= in_w, "modconv_tc2_view: out extent");
    P.tap_mask = view->tap_mask;
    P.reduce_add = view->reduce ? 1 : 0;
  }
  if (!transposed) {
    P.OH = in_h; P.OW = in_w; P.grid_h = in_h; P.grid_w = in_w; P.out_stride = 1; P.ng = 1;
  } else {
    P.OH = 2 * in_h + 1; P.OW = 2 * in_w + 1; P.grid_h = in_h + 1; P.grid_w = in_w + 1; P.out_stride = 2; P.ng = 4;
  }
  // K chunk = one 128-byte (or 64-byte) swizzled row of channels: 64 / 32 bf16 or 32 / 16 fp32 (tf32 mode)
  const int esize = tf32 ? 4 : 2;
  P.bk = (Cin % (128 / esize) == 0) ? 128 / esize : 64 / esize;
  const int row_bytes = P.bk * esize;
  P.mt = (P.grid_h > kSubTileH) ? 2 : 1;
  // the RGB-only last layer (32 -> 32 channels): 512-pixel tiles, two pixels per epilogue thread
  const bool pair = rgb && rgb->pair;
  const bool pair_plain = view && view->pair;   // no fused ToRGB: the ordinary staged epilogue stores the 64-channel "pairs"
  if (pair) {
    W2E_CHECK_ARG(!transposed && !out_mod && Cin == 64 && Cout == 64 && in_h > kSubTileH && in_h % 2 == 0 && !tf32 && !view,
                  "modconv_tc2_rgb_pair: needs the 32-channel layer without a modulated output (as 64-channel pixel pairs), even height");
    P.pair = 1;
    P.OW_real = 2 * in_w;
  }
  if (pair_plain) {
    W2E_CHECK_ARG(!transposed && !rgb && !out_mod && out && !noise && !bias && !out_scale && Cin == 64 && Cout == 64 &&
                      in_h > kSubTileH && !tf32,
                  "modconv_tc2_pair: a plain 32 -> 32 channel convolution without epilogue terms, more than 16 rows");
    P.pair = 1;
    P.OW_real = 2 * in_w;
  }
  // fused dgrad of the transposed convolution: an A stage is four class tiles, so 128-pixel tiles (direct-store epilogue)
  if (dg4) P.mt = 1;
  const bool mt4 = allow_mt4 && g_ts_mode != 0 && rgb && !out && !out_mod && Cin == 32 && Cout == 32 && in_h >= 64;
  if (mt4) P.mt = 4;
  // Transposed conv: 4 parity classes x MT sub-tiles x bn columns must fit 512 TMEM columns.  Measured (B=32):
  // (MT 1, bn 128) beats (MT 2, bn 64) on every >=128-channel up-layer (0.48 vs 0.59 ms at 512->256@64^2) although
  // its 16 KB weight blocks feed only 4 MMAs each and the ring is then bound by the SM's TMA ingest (~42 B/clk:
  // 385 cycles per block against 256 cycles of MMAs) -- bn 64 serialises MMA and epilogue on one accumulator set.
  // Plain conv + fused ToRGB at Cout == 256 (256 -> 256 @128^2): two 128-column tiles instead of one 256-column tile.
  // A 256-pixel x 256-column tile fills the 512 TMEM columns, so MMA and epilogue alternate (measured 0.51 ms against
  // 0.27 ms of MMAs); 128 columns leave room for two accumulator sets and the staged 16-warp epilogue, and the two
  // column tiles add their partial ToRGB sums into the zero-initialised image.  Flag bit 7 = off (A/B).
  const bool clipped_out = view && !dg4 && !view->pair && (view->reduce || view->out_h != in_h || view->out_w != in_w);   // needs the TMA-store epilogue
  const bool rgb_split = !transposed && rgb && Cout == 256 && g_ts_mode != 0 && !(g_flags & 128) && in_h > kSubTileH &&
                         rgb->rgb_dtype == W2E_F32;
  const int bn_max = transposed ? ((g_flags & 8) ? 64 : 128) : ((rgb_split || clipped_out) ? 128 : 256);
  P.bn = 16;
  for (int cand : {256, 128, 64, 32, 16})
    if (cand <= bn_max && Cout % cand == 0) { P.bn = cand; break; }
  if (P.ng * P.mt * P.bn > 512) P.mt = 1;
  // 32-channel transposed layer: 128-pixel tiles leave room for FOUR accumulator sets, so each epilogue group's
  // MMAs of the next tile overlap its drain of the current one (0.64 -> 0.57 ms at 64->32@512^2, the HBM floor is
  // 0.49).  Not for 64 channels: its weights are not resident and smaller tiles make the ring ingest-bound.
  if (transposed && P.bn <= 32 && !fb && !(g_flags & 16)) P.mt = 1;
  // flag bit 5 (A/B): the same trade for the 64-channel transposed layer (two accumulator sets of 128 pixels instead
  // of one of 256: MMA and epilogue overlap, at the price of weight blocks that feed half as many MMAs)
  if (transposed && P.bn == 64 && Cout == 64 && !fb && (g_flags & 32)) P.mt = 1;
  const int acc_cols = P.ng * P.mt * P.bn;
  const int nbuf_plain = (acc_cols * 2 <= 512) ? 2 : 1;
  const int nbuf_ts = (acc_cols * 4 <= 512) ? 4 : nbuf_plain;   // TS flavour: two buffers per epilogue group
  auto set_nbuf = [&](int n) {
    P.nbuf = n;
    P.tmem_cols = 32;
    while (P.tmem_cols < acc_cols * n) P.tmem_cols <<= 1;
  };
  set_nbuf(nbuf_plain);
  P.box_rows = kSubTileH * P.mt + 2;
  P.a_box_bytes = P.box_rows * P.pitch * row_bytes;
  P.a_stage_bytes = (P.a_box_bytes + 1023) / 1024 * 1024;
  if (dg4) {
    P.cls16 = P.a_stage_bytes >> 4;
    P.a_box_bytes *= 4;
    P.a_stage_bytes *= 4;
  }
  P.b_block_bytes = P.bn * row_bytes;
  P.tiles_x = ceil_div(P.grid_w, kTileW);
  P.tiles_y = ceil_div(P.grid_h, kSubTileH * P.mt);
  P.tile_dy = kSubTileH * P.mt; P.tile_dx = kTileW; P.tile_o = 0;
  if (fb) {
    // 64 x 16 pre-blur pixels per tile, of which the 60 x 12 outputs with a complete 4x4 window are kept
    if (!(transposed && P.mt == 2 && acc_cols <= 512 && P.bn % 32 == 0 && g_ts_mode != 0))
      return set_error(W2E_ERR_UNSUPPORTED, "modconv_tc2_upblur: shape not supported by the fused kernel");
    P.fb = 1; P.OH2 = 2 * in_h; P.OW2 = 2 * in_w;
    for (int i = 0; i < 4; ++i) { P.fbv[i] = fb->fv[i]; P.fbh[i] = fb->fh[i]; }
    P.tile_dy = 30; P.tile_dx = 6; P.tile_o = -1;
    P.tiles_x = ceil_div(P.OW2, 12);
    P.tiles_y = ceil_div(P.OH2, 60);
  }
  P.tiles_n = Cout / P.bn;
  const int64_t ntiles = (int64_t)P.tiles_x * P.tiles_y * P.tiles_n * B;
  W2E_CHECK_ARG(ntiles < (1ll << 31), "modconv_tc2: too many tiles");
  P.ntiles = (int)ntiles;
  const int kchunks = Cin / P.bk;

  // TS epilogue eligibility (see the header): staging units of <= 64 channels, TMA-able strides
  const int n_out = pair ? 0 : (out ? 1 : 0) + (out_mod ? 1 : 0);   // (pair mode stores its activation from registers: no staging)
  bool ts = !tf32 && g_ts_mode != 0 && (transposed || P.mt >= 2) && P.bn >= 32 && P.bn <= 128;
  if (rgb && nbuf_plain != 2) ts = false;   // fused ToRGB needs a thread's whole channel row: no unit split
  if (transposed && !fb) ts = ts && !noise && !bias && !next_scale && !out_mod && act == W2E_ACT_NONE;
  if (noise && !fb) ts = ts && (P.OW * 4) % 16 == 0 && (((uintptr_t)noise & 15) == 0);
  if (rgb && rgb->skip) ts = ts && ((pair ? P.OW : P.OW / 2) * 4) % 16 == 0 && (((uintptr_t)rgb->skip & 15) == 0);
  if (out) ts = ts && (((uintptr_t)out & 15) == 0);
  if (out_mod) ts = ts && (((uintptr_t)out_mod & 15) == 0);
  if (rgb && P.bn != Cout && !rgb_split) ts = false;

  // shared-memory plan: resident weights when small, else a ring of weight blocks; with the TS
  // epilogue also the output staging slots (per epilogue group and half) and the two stages of
  // epilogue inputs.  TS candidates in order of preference: 64-channel units with two slots, one
  // slot, then 32-channel units; if none fits the direct-store epilogue is used.
  const int kBarsBytes = ((int)sizeof(Tc2Bars) + 15) / 16 * 16;
  const int w_bytes = 9 * kchunks * P.b_block_bytes;
  int smem_bytes = 0;
  struct Cand { int unit_ch, slots; };
  // transposed conv with a single accumulator set (per-class hand-over): 32-channel units, so that all four
  // half-groups of epilogue warps have a unit of every class (and the weight ring gets the shared memory)
  const bool small_units = transposed && !fb && acc_cols == 512 && !(g_flags & 256) && P.bn >= 64;
  const Cand cands_std[5] = {{64, 2}, {64, 1}, {32, 2}, {32, 1}, {0, 0}};
  const Cand cands_small[5] = {{32, 2}, {32, 1}, {64, 2}, {64, 1}, {0, 0}};
  const Cand* cands = small_units ? cands_small : cands_std;
  // Candidates 0..3 (staged epilogue) are tried twice when the weights could be resident: first with resident weights, then
  // -- if no staged plan fits next to them -- with the weight RING, which needs less shared memory; only then the direct-
  // store epilogue.  (64 -> 64 @512^2 with both outputs and the fused ToRGB, the training forward: 74 KB of weights + two
  // A stages leave no room for the staging slots, and the 8-warp direct-store epilogue took 0.74 ms at batch 16 against
  // 0.31 ms for the one-output inference launch.)  Flag bit 9 = off (A/B).
  const bool can_be_resident = Cout / P.bn == 1 && w_bytes <= 80 * 1024;
  const bool ring_retry = ts && can_be_resident && !fb && !pair && !dg4 && !(g_flags & 512);
  const int n_plans = ring_retry ? 9 : 5;
  for (int pi = ts ? 0 : (n_plans - 1); pi < n_plans && smem_bytes == 0; ++pi) {
    // plan index -> (candidate, weights may be resident): 0..3 staged + resident, [4..7 staged + ring,] last = direct store
    const int ci = pi == n_plans - 1 ? 4 : pi % 4;
    const bool allow_wres = !(ring_retry && pi >= 4 && pi < 8);
    const bool use_ts = cands[ci].unit_ch != 0;
    int extra = 0, ts_bytes = 0;
    if (use_ts) {
      P.ts_unit_ch = P.bn < cands[ci].unit_ch ? P.bn : cands[ci].unit_ch;
      if (cands[ci].unit_ch == 32 && !small_units && P.bn <= 32) continue;   // same unit as the 64-channel candidates
      P.ts_slots = cands[ci].slots;
      if (!fb && n_out > P.ts_slots) continue;
      P.ts_unit_bytes = 128 * P.ts_unit_ch * 2;
      P.use_e = 1;
      P.e_noise_bytes = (P.mt == 4 || pair) ? 2048 : 1024;
      P.skip_box_bytes = pair ? 2048 : kSkipBoxBytes;
      const int skip_w = pair ? 16 : kSkipBoxW, noise_w_px = pair ? 2 * kTileW : kTileW;
      P.e_info_off = P.e_noise_bytes + ((rgb && rgb->skip) ? P.mt * P.skip_box_bytes : 0);
      P.e_stage_bytes = P.e_info_off + 128;
      P.e_bytes = (noise ? noise_w_px * kSubTileH * P.mt * 4 : 0) + ((rgb && rgb->skip) ? P.mt * 3 * 10 * skip_w * 4 : 0);
      ts_bytes = fb ? 64 * 1024 : (n_out ? 2 * 2 * P.ts_slots * P.ts_unit_bytes : 0);
      if (fb) { P.e_bytes = 0; P.e_info_off = 0; P.e_stage_bytes = 128; }
      extra = 1024 /*alignment of the staging area*/ + kEStages * P.e_stage_bytes;
    } else {
      P.ts_unit_ch = P.ts_unit_bytes = P.ts_slots = P.use_e = P.e_stage_bytes = P.e_bytes = 0;
    }
    const int smem_limit = 227 * 1024 - 1024 /*alignment slack*/ - kBarsBytes - extra;
    P.a_stages = kchunks == 1 ? 3 : 2;
    P.wres = (allow_wres && P.tiles_n == 1 && w_bytes <= 80 * 1024) ? 1 : 0;
    int b_bytes = 0;
    if (P.wres) {
      if (use_ts) P.a_stages = kT2MaxA;   // epilogue-bound layers: a deep A ring hides the DRAM latency of the tile loads
      while (P.a_stages * P.a_stage_bytes + w_bytes + ts_bytes > smem_limit && P.a_stages > 2) --P.a_stages;
      // two MMA issuers (kernel: kMmaWarps) need a stage count that is a multiple of 2 * kchunks
      if (use_ts && P.a_stages > 2 * kchunks) P.a_stages -= P.a_stages % (2 * kchunks);
      if (P.a_stages * P.a_stage_bytes + w_bytes + ts_bytes > smem_limit) continue;
      P.b_stages = 1;
      b_bytes = w_bytes;
    } else {
      int avail = smem_limit - P.a_stages * P.a_stage_bytes - ts_bytes;
      if (use_ts && P.a_stages > 2 && avail / P.b_block_bytes < 4) {   // a third A stage is a luxury: the ring comes first
        P.a_stages = 2;
        avail = smem_limit - P.a_stages * P.a_stage_bytes - ts_bytes;
      }
      P.b_stages = avail / P.b_block_bytes;
      if (P.b_stages > kT2MaxB) P.b_stages = kT2MaxB;
      if (P.b_stages < (use_ts ? 4 : 2)) continue;
      b_bytes = P.b_stages * P.b_block_bytes;
    }
    const int ab = P.a_stages * P.a_stage_bytes + b_bytes;
    P.ts_off = (ab + 1023) / 1024 * 1024;
    P.e_off = P.ts_off + ts_bytes;
    P.bars_off = use_ts ? P.e_off + kEStages * P.e_stage_bytes : (ab + 15) / 16 * 16;
    smem_bytes = P.bars_off + kBarsBytes + 1024;
    ts = use_ts;
    set_nbuf(use_ts ? nbuf_ts : nbuf_plain);
  }
  W2E_CHECK_ARG(smem_bytes > 0, "modconv_tc2: shared memory plan does not fit (Cin %d Cout %d)", Cin, Cout);
  if (P.mt == 4 && !ts)   // the direct-store epilogue has no 4-sub-tile variant: plan again with 256-pixel tiles
    return run_tc2(xs, w, out_scale, bias, noise, noise_w, noise_batch, next_scale, out, out_mod, error_flag, B, Cin, Cout,
                   in_h, in_w, transposed, act, rgb, cfg, stream, false, fb, tf32, view);
  W2E_CHECK_ARG(smem_bytes > 0 && smem_bytes <= 227 * 1024, "modconv_tc2: %d bytes of shared memory needed", smem_bytes);

  CUtensorMap ma, mb;
  Tc2Maps M;
  memset(&M, 0, sizeof(M));
  if (dg4 && !P.wres)
    return set_error(W2E_ERR_UNSUPPORTED, "modconv_tc2_dgrad_up: needs weights resident in shared memory (Cin %d, Cout %d)", Cin, Cout);
  {
    const uint64_t es = (uint64_t)esize;
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)in_w, (uint64_t)in_h, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)Cin * es, (uint64_t)in_w * Cin * es, (uint64_t)in_h * in_w * Cin * es};
    if (view && !dg4 && !dgk && !view->pair) {
      strides[0] = (uint64_t)view->stride_x * es; strides[1] = (uint64_t)view->stride_y * es;
      strides[2] = (uint64_t)view->stride_b * es;
    }
    const uint32_t box[4] = {(uint32_t)P.bk, (uint32_t)P.pitch, (uint32_t)P.box_rows, 1u};
    if (dg4 || dgk) {
      // class (py, px) of the [B, 2h+1, 2w+1, K] gradient: pixels (2r+py, 2s+px), extent (h+1-py) x (w+1-px)
      const int K = dgk ? Cin / 4 : Cin, h = dgk ? in_h - 1 : in_h, w_ = dgk ? in_w - 1 : in_w;
      const uint64_t zh = 2 * (uint64_t)h + 1, zw = 2 * (uint64_t)w_ + 1;
      const uint64_t cstrides[3] = {2 * (uint64_t)K * es, 2 * zw * K * es, zh * zw * K * es};
      for (int c = 0; c < 4; ++c) {
        const int py = c >> 1, px = c & 1;
        const uint64_t cdims[4] = {(uint64_t)K, (uint64_t)(w_ + 1 - px), (uint64_t)(h + 1 - py), (uint64_t)B};
        const __nv_bfloat16* base = (const __nv_bfloat16*)xs + ((int64_t)py * (int64_t)zw + px) * K;
        int rc = make_bf16_map(c == 0 ? &ma : &M.a4[c - 1], base, 4, cdims, cstrides, box, row_bytes);
        if (rc) return rc;
      }
    } else {
      int rc = tf32 ? make_f32_swizzled_map(&ma, xs, 4, dims, strides, box, row_bytes)
                    : make_bf16_map(&ma, xs, 4, dims, strides, box, row_bytes);
      if (rc) return rc;
    }
  }
  {
    const uint64_t es = (uint64_t)esize;
    const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 9u};
    const uint64_t strides[2] = {(uint64_t)Cin * es, (uint64_t)Cout * Cin * es};
    const uint32_t box[3] = {(uint32_t)P.bk, (uint32_t)P.bn, 1u};
    int rc = tf32 ? make_f32_swizzled_map(&mb, w, 3, dims, strides, box, row_bytes)
                  : make_bf16_map(&mb, w, 3, dims, strides, box, row_bytes);
    if (rc) return rc;
  }
  // 2-CTA clusters: the pair works on the same tile of two consecutive samples and shares every weight block
  int clog = g_cluster_mode;
  while (clog > 0 && (B % (1 << clog) != 0 || (P.bn >> clog) < 8)) --clog;
  if (clog > 0 && !P.wres && !fb && sm_count() >= (1 << clog)) {
    P.cluster = clog;
    P.B = B >> clog;
    P.ntiles = P.ntiles >> clog;
    const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 9u};
    const uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
    const uint32_t box[3] = {(uint32_t)P.bk, (uint32_t)(P.bn >> clog), 1u};
    int rc = make_bf16_map(&M.bh, w, 3, dims, strides, box, row_bytes);
    if (rc) return rc;
  }
  // grouped weight requests (one TMA box per tap row) for the weight-ring kernels of the TS flavour; flag bit 6 = off (A/B)
  if (ts && !P.wres && !P.cluster && !tf32 && !(g_flags & 64) && P.b_stages >= 6) {
    P.bgroup = 3;
    P.b_stages -= P.b_stages % 3;
    const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 9u};
    const uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
    const uint32_t box[3] = {(uint32_t)P.bk, (uint32_t)P.bn, 3u};
    int rc = make_bf16_map(&M.b3, w, 3, dims, strides, box, row_bytes);
    if (rc) return rc;
  } else {
    P.bgroup = 1;
  }
  // per-class accumulator hand-over (transposed conv with a single accumulator set); flag bit 8 = off (A/B)
  P.percls = (transposed && ts && !fb && P.nbuf == 1 && P.bgroup == 3 && !(g_flags & 256)) ? 1 : 0;
  if (fb && !ts) return set_error(W2E_ERR_UNSUPPORTED, "modconv_tc2_upblur: shared memory plan does not fit");
  if (dgk && !(ts && !P.wres && P.bk == 64 && !P.cluster))
    return set_error(W2E_ERR_UNSUPPORTED, "modconv_tc2_dgrad_up: the K-loop form needs the staged epilogue with the weight ring");
  if ((pair || pair_plain) && !(ts && P.wres && P.mt == 2 && P.bk == 64))
    return set_error(W2E_ERR_UNSUPPORTED, "modconv_tc2_rgb_pair: needs the staged epilogue with resident weights");
  if (clipped_out && !ts)
    return set_error(W2E_ERR_UNSUPPORTED, "modconv_tc2_view: a clipped / accumulating output needs the TMA-store epilogue "
                                          "(more than 16 rows, 32..128-column tiles)");
  if (ts && !fb) {
    if (noise) {
      const uint64_t ow = pair ? (uint64_t)P.OW_real : (uint64_t)P.OW;
      const uint64_t dims[3] = {ow, (uint64_t)P.OH, (uint64_t)noise_batch};
      const uint64_t strides[2] = {ow * 4, (uint64_t)P.OH * ow * 4};
      const uint32_t box[3] = {(uint32_t)(pair ? 2 * kTileW : kTileW), (uint32_t)(kSubTileH * P.mt), 1u};
      int rc = make_f32_map(&M.noise, noise, 3, dims, strides, box);
      if (rc) return rc;
    }
    if (rgb && rgb->skip) {
      const uint64_t h2 = (uint64_t)P.OH / 2, w2 = pair ? (uint64_t)P.OW : (uint64_t)P.OW / 2;   // (pairs = half-res columns)
      const uint64_t dims[3] = {w2, h2, (uint64_t)B * 3};
      const uint64_t strides[2] = {w2 * 4, h2 * w2 * 4};
      const uint32_t box[3] = {(uint32_t)(pair ? 16 : kSkipBoxW), 10u, 3u};
      int rc = make_f32_map(&M.skip, rgb->skip, 3, dims, strides, box);
      if (rc) return rc;
    }
    const uint32_t sbox[4] = {(uint32_t)P.ts_unit_ch, (uint32_t)kTileW, (uint32_t)kSubTileH, 1u};
    if (!transposed) {
      const bool vw = view && !view->pair;
      const uint64_t ow = vw ? (uint64_t)view->out_w : (uint64_t)P.OW, oh = vw ? (uint64_t)view->out_h : (uint64_t)P.OH;
      const uint64_t dims[4] = {(uint64_t)Cout, ow, oh, (uint64_t)B};
      const uint64_t strides[3] = {(uint64_t)Cout * 2, ow * Cout * 2, oh * ow * Cout * 2};
      if (out && !pair) {
        int rc = make_bf16_map(&M.st[0], out, 4, dims, strides, sbox, P.ts_unit_ch * 2);
        if (rc) return rc;
      }
      if (out_mod) {
        int rc = make_bf16_map(&M.st[1], out_mod, 4, dims, strides, sbox, P.ts_unit_ch * 2);
        if (rc) return rc;
      }
    } else {
      // class (py, px): pixels (2j+py, 2i+px) of the (2h+1) x (2w+1) output as a strided [B, rows, cols, C] view
      for (int g = 0; g < 4; ++g) {
        const int py = g >> 1, px = g & 1;
        const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)((P.OW - px + 1) / 2), (uint64_t)((P.OH - py + 1) / 2), (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)Cout * 4, (uint64_t)P.OW * Cout * 4, (uint64_t)P.OH * P.OW * Cout * 2};
        const __nv_bfloat16* base = (const __nv_bfloat16*)out + ((int64_t)py * P.OW + px) * Cout;
        int rc = make_bf16_map(&M.st[g], base, 4, dims, strides, sbox, P.ts_unit_ch * 2);
        if (rc) return rc;
      }
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int ks = row_bytes / 32;   // MMAs (K = 32 bytes each) per K chunk
  if (tf32) {
    if (P.cluster) return set_error(W2E_ERR_UNSUPPORTED, "modconv_tc2_tf32: no cluster mode");
#define W2E_TC2_TF32(TR_, MT_, KS_, WR_) \
  if ((transposed != 0) == TR_ && P.mt == MT_ && ks == KS_ && (P.wres != 0) == WR_) \
    return launch_tc2<TR_, MT_, KS_, WR_, false, false, true>(ma, mb, P, M, smem_bytes, g_max_ctas, st);
    W2E_TC2_TF32(false, 1, 4, false) W2E_TC2_TF32(false, 2, 4, false) W2E_TC2_TF32(false, 1, 2, false)
    W2E_TC2_TF32(false, 2, 2, false) W2E_TC2_TF32(false, 1, 4, true) W2E_TC2_TF32(false, 2, 4, true)
    W2E_TC2_TF32(false, 1, 2, true) W2E_TC2_TF32(false, 2, 2, true)
    W2E_TC2_TF32(true, 1, 4, false) W2E_TC2_TF32(true, 2, 4, false) W2E_TC2_TF32(true, 1, 2, false)
    W2E_TC2_TF32(true, 2, 2, false) W2E_TC2_TF32(true, 1, 4, true) W2E_TC2_TF32(true, 2, 4, true)
    W2E_TC2_TF32(true, 1, 2, true) W2E_TC2_TF32(true, 2, 2, true)
#undef W2E_TC2_TF32
    return set_error(W2E_ERR_UNSUPPORTED, "modconv_tc2_tf32: no kernel variant");
  }
  if (rgb) {
    W2E_CHECK_ARG((P.mt == 2 || P.mt == 4) && (P.tiles_n == 1 || (P.tiles_n == 2 && (!ts || rgb_split))),
                  "modconv_tc2_rgb: unsupported tiling (mt %d, n tiles %d)", P.mt, P.tiles_n);
    if (P.mt == 4) {
      if (P.wres) return launch_tc2<false, 4, 2, true, true, true>(ma, mb, P, M, smem_bytes, g_max_ctas, st);
      return launch_tc2<false, 4, 2, false, true, true>(ma, mb, P, M, smem_bytes, g_max_ctas, st);
    }
#define W2E_TC2_RGB(KS_, WR_) \
  if (ks == KS_ && (P.wres != 0) == WR_) { \
    if (ts) return launch_tc2<false, 2, KS_, WR_, true, true>(ma, mb, P, M, smem_bytes, g_max_ctas, st); \
    return launch_tc2<false, 2, KS_, WR_, true, false>(ma, mb, P, M, smem_bytes, g_max_ctas, st); \
  }
    W2E_TC2_RGB(4, false) W2E_TC2_RGB(4, true) W2E_TC2_RGB(2, false) W2E_TC2_RGB(2, true)
#undef W2E_TC2_RGB
  }
  if (ts) {
#define W2E_TC2_TS(TR_, MT_, KS_, WR_) \
  if ((transposed != 0) == TR_ && P.mt == MT_ && ks == KS_ && (P.wres != 0) == WR_) \
    return launch_tc2<TR_, MT_, KS_, WR_, false, true>(ma, mb, P, M, smem_bytes, g_max_ctas, st);
    W2E_TC2_TS(false, 2, 4, false) W2E_TC2_TS(false, 2, 4, true) W2E_TC2_TS(false, 2, 2, false) W2E_TC2_TS(false, 2, 2, true)
    W2E_TC2_TS(true, 2, 4, false) W2E_TC2_TS(true, 2, 4, true) W2E_TC2_TS(true, 2, 2, false) W2E_TC2_TS(true, 2, 2, true)
    W2E_TC2_TS(true, 1, 4, false) W2E_TC2_TS(true, 1, 4, true) W2E_TC2_TS(true, 1, 2, false) W2E_TC2_TS(true, 1, 2, true)
#undef W2E_TC2_TS
  }
#define W2E_TC2_CASE(TR_, MT_, KS_, WR_) \
  if ((transposed != 0) == TR_ && P.mt == MT_ && ks == KS_ && (P.wres != 0) == WR_) \
    return launch_tc2<TR_, MT_, KS_, WR_, false, false>(ma, mb, P, M, smem_bytes, g_max_ctas, st);
  W2E_TC2_CASE(false, 1, 4, false) W2E_TC2_CASE(false, 2, 4, false) W2E_TC2_CASE(false, 1, 2, false)
  W2E_TC2_CASE(false, 2, 2, false) W2E_TC2_CASE(false, 1, 4, true) W2E_TC2_CASE(false, 2, 4, true)
  W2E_TC2_CASE(false, 1, 2, true) W2E_TC2_CASE(false, 2, 2, true)
  W2E_TC2_CASE(true, 1, 4, false) W2E_TC2_CASE(true, 2, 4, false) W2E_TC2_CASE(true, 1, 2, false)
  W2E_TC2_CASE(true, 2, 2, false) W2E_TC2_CASE(true, 1, 4, true) W2E_TC2_CASE(true, 2, 4, true)
  W2E_TC2_CASE(true, 1, 2, true) W2E_TC2_CASE(true, 2, 2, true)
#undef W2E_TC2_CASE
  return set_error(W2E_ERR_UNSUPPORTED, "modconv_tc2: no kernel variant");
}

extern "C" int w2e_modconv_tc2(const void* xs, const void* w, const float* out_scale, const float* bias,
                               const float* noise, const float* noise_w, int noise_batch, const float* next_scale,
                               void* out, void* out_mod, int* error_flag, int B, int Cin, int Cout, int in_h, int in_w,
                               int transposed, int act, const w2e_tc2_config* cfg, void* stream) {
  return run_tc2(xs, w, out_scale, bias, noise, noise_w, noise_batch, next_scale, out, out_mod, error_flag, B, Cin, Cout,
                 in_h, in_w, transposed, act, nullptr, cfg, stream);
}

// Plain 3x3 convolution over a STRIDED VIEW of a channels-last bf16 tensor with a subset of the nine taps: the dgrad of
// the transposed x2 convolution is a stride-2 convolution of the (2h+1)^2 upstream gradient, run as four such launches --
// one per output-parity class of the gradient (a view with doubled pixel / row strides) with the 4 / 2 / 2 / 1 taps
// that class feeds -- whose results w2e_sum4_nhwc adds.
extern "C" int w2e_modconv_tc2_view(const void* xs, const void* w, const float* out_scale, const float* next_scale,
                                    void* out, void* out_mod, int* error_flag, int B, int Cin, int Cout, int in_h,
                                    int in_w, int64_t stride_x, int64_t stride_y, int64_t stride_b, int tap_mask,
                                    int out_h, int out_w, int accumulate, const w2e_tc2_config* cfg, void* stream) {
  const ViewArgs v{stride_x, stride_y, stride_b, tap_mask, out_h, out_w, accumulate, 0, 0, 0};
  return run_tc2(xs, w, out_scale, nullptr, nullptr, nullptr, 0, next_scale, out, out_mod, error_flag, B, Cin, Cout, in_h,
                 in_w, 0, W2E_ACT_NONE, nullptr, cfg, stream, true, nullptr, false, &v);
}

// Fused dgrad of the transposed x2 convolution (conv_transpose2d(stride 2), models/stylegan2/model.py:249-258): the input
// gradient gx [B,h,w,Cout] from the upstream gradient gz [B,2h+1,2w+1,Cin] in ONE launch,
//   gx[j,i] = sum_{ky,kx} gz[2j+ky, 2i+kx] . w[ky*3+kx],   w: bf16 [9][Cout][Cin] (the forward weight, equalised-lr scale folded in),
// the four parity classes of gz read through strided TMA views and all nine taps accumulated in one TMEM accumulator:
// gz is read once and gx written once (w2e_modconv_tc2_view x 4 re-reads gx three times through its reduce-adds).
// Returns W2E_ERR_UNSUPPORTED when the nine weight blocks do not fit in shared memory next to the four class tiles
// (then the per-class launches are the path).  out_scale [B,Cout] (or NULL) multiplies the result.
extern "C" int w2e_modconv_tc2_dgrad_up(const void* gz, const void* w, const float* out_scale, void* gx, int* error_flag,
                                        int B, int Cin, int Cout, int h, int w_, const w2e_tc2_config* cfg, void* stream) {
  ViewArgs v;
  memset(&v, 0, sizeof(v));
  v.tap_mask = 0x1ff; v.out_h = h; v.out_w = w_; v.dg4 = 1;
  return run_tc2(gz, w, out_scale, nullptr, nullptr, nullptr, 0, nullptr, gx, nullptr, error_flag, B, Cin, Cout, h, w_, 0,
                 W2E_ACT_NONE, nullptr, cfg, stream, true, nullptr, false, &v);
}

// Plain 32 -> 32 channel 3x3 convolution on PIXEL PAIRS (the dgrad of the last layer in the training step: no epilogue
// terms): xs bf16 [B,H,W,32] read as [B,H,W/2,64], w_pair bf16 [9][64][64] (see w2e_modconv_tc2_rgb_pair), out bf16
// [B,H,W,32] written as [B,H,W/2,64] by the ordinary staged epilogue.  N = 64 MMAs with the zero K slices skipped.
extern "C" int w2e_modconv_tc2_pair(const void* xs, const void* w_pair, void* out, int* error_flag, int B, int H, int W,
                                    const w2e_tc2_config* cfg, void* stream) {
  W2E_CHECK_ARG(W > 0 && W % 16 == 0, "modconv_tc2_pair: the width must be a multiple of 16");
  ViewArgs v;
  memset(&v, 0, sizeof(v));
  v.tap_mask = 0x1ff; v.pair = 1;
  return run_tc2(xs, w_pair, nullptr, nullptr, nullptr, nullptr, 0, nullptr, out, nullptr, error_flag, B, 64, 64, H, W / 2, 0,
                 W2E_ACT_NONE, nullptr, cfg, stream, true, nullptr, false, &v);
}

// K-LOOP form of the fused dgrad for the layers whose weights do not fit next to four class tiles (>= 64 output channels
// of the forward): the K loop of ONE accumulator walks the four parity classes of the gradient one after the other --
// chunk kc reads class kc / (K / 64) through that class's strided view and issues only the taps the class feeds (4 / 2 /
// 2 / 1) -- so gx is written once instead of stored + three times read-modify-written (w2e_modconv_tc2_view x 4).
// w4: bf16 [9][Cout][4 K], the per-class weights of w2e_modconv_tc2_view concatenated along K in the order (0,0), (0,1),
// (1,0), (1,1) (unused taps are never read); tap_masks: class c = py * 2 + px at bit 9 c.  h > 15, K a multiple of 64.
extern "C" int w2e_modconv_tc2_dgrad_up_k(const void* gz, const void* w4, long long tap_masks, const float* out_scale, void* gx,
                                          int* error_flag, int B, int K, int Cout, int h, int w_, const w2e_tc2_config* cfg,
                                          void* stream) {
  W2E_CHECK_ARG(K > 0 && K % 64 == 0 && h > 0 && w_ > 0, "modconv_tc2_dgrad_up_k: K must be a multiple of 64");
  ViewArgs v;
  memset(&v, 0, sizeof(v));
  v.tap_mask = 0x1ff; v.out_h = h; v.out_w = w_; v.dg4 = 2; v.dg_masks = tap_masks;
  return run_tc2(gz, w4, out_scale, nullptr, nullptr, nullptr, 0, nullptr, gx, nullptr, error_flag, B, 4 * K, Cout, h + 1, w_ + 1,
                 0, W2E_ACT_NONE, nullptr, cfg, stream, true, nullptr, false, &v);
}

// tf32 mode (north star item 1: "bf16 and tf32 modes"): xs / w / out / out_mod are fp32 (channels-last activations,
// [9][Cout][Cin] weights), the MMA reads them as tf32 and accumulates in fp32.
extern "C" int w2e_modconv_tc2_tf32(const float* xs, const float* w, const float* out_scale, const float* bias,
                                    const float* noise, const float* noise_w, int noise_batch,
                                    const float* next_scale, float* out, float* out_mod, int* error_flag, int B, int Cin,
                                    int Cout, int in_h, int in_w, int transposed, int act, const w2e_tc2_config* cfg,
                                    void* stream) {
  return run_tc2(xs, w, out_scale, bias, noise, noise_w, noise_batch, next_scale, out, out_mod, error_flag, B, Cin, Cout,
                 in_h, in_w, transposed, act, nullptr, cfg, stream, false, nullptr, true);
}

extern "C" int w2e_modconv_tc2_rgb(const void* xs, const void* w, const float* out_scale, const float* bias,
                                   const float* noise, const float* noise_w, int noise_batch, const float* next_scale,
                                   void* out, void* out_mod, int* error_flag, int B, int Cin, int Cout, int in_h,
                                   int in_w, int act, const float* rgb_w, const float* rgb_style, const float* rgb_bias,
                                   const float* rgb_skip, const float* host_taps1d, void* rgb, int rgb_dtype,
                                   const w2e_tc2_config* cfg, void* stream) {
  const RgbArgs a{rgb_w, rgb_style, rgb_bias, rgb_skip, host_taps1d, rgb, rgb_dtype, 0};
  return run_tc2(xs, w, out_scale, bias, noise, noise_w, noise_batch, next_scale, out, out_mod, error_flag, B, Cin, Cout,
                 in_h, in_w, 0, act, &a, cfg, stream);
}

// x-pair mode of the RGB-only 32 -> 32 channel layer (the last layer of the 1024^2 generator: 13 % of a step).  A
// tcgen05.mma with N = 32 costs 44 cycles, one with N = 64 only 48 (operand fetch, tools/umma_bench.cu), so the layer is
// run on PIXEL PAIRS: the channels-last input [B,H,W,32] IS [B,H,W/2,64] (a pair's 2 x 32 channels are contiguous), the
// output pair has 2 x 32 = 64 channels, and the 3x3 kernel becomes a 3x3 kernel over pairs whose left / right taps use
// only one pixel of the neighbouring pair: 3 x (2 + 4 + 2) K steps of N = 64 per 128 pairs instead of 2 x 18 of N = 32
// per 256 pixels (1152 vs 1584 cycles).  w_pair: bf16 [9][64][64] pair weights,
// w_pair[ky*3+dj+1][a*32+o][b*32+c] = W[ky][2dj+b-a+1][o][c] (zero where that tap index falls outside 0..2);
// out_scale [B,32], bias [32], rgb_w [3,32], rgb_style [B,32]: the ordinary 32-channel arrays (the epilogue indexes them
// with column & 31); out: optional bf16 [B,H,W,32] activation (the training forward keeps it), written from registers; H, W:
// the image size in pixels (W a multiple of 16); everything else as w2e_modconv_tc2_rgb.
extern "C" int w2e_modconv_tc2_rgb_pair(const void* xs, const void* w_pair, const float* out_scale, const float* bias,
                                        const float* noise, const float* noise_w, int noise_batch, void* out, int* error_flag, int B,
                                        int H, int W, int act, const float* rgb_w, const float* rgb_style,
                                        const float* rgb_bias, const float* rgb_skip, const float* host_taps1d, void* rgb,
                                        int rgb_dtype, const w2e_tc2_config* cfg, void* stream) {
  W2E_CHECK_ARG(W > 0 && W % 16 == 0, "modconv_tc2_rgb_pair: the width must be a multiple of 16");
  const RgbArgs a{rgb_w, rgb_style, rgb_bias, rgb_skip, host_taps1d, rgb, rgb_dtype, 1};
  W2E_CHECK_ARG(out == nullptr || (((uintptr_t)out & 15) == 0), "modconv_tc2_rgb_pair: out must be 16-byte aligned");
  return run_tc2(xs, w_pair, out_scale, bias, noise, noise_w, noise_batch, nullptr, out, nullptr, error_flag, B, 64, 64,
                 H, W / 2, 0, act, &a, cfg, stream);
}


extern "C" int w2e_modconv_tc2_upblur(const void* xs, const void* w, const float* out_scale, const float* host_taps,
                                      const float* bias, const float* noise, const float* noise_w, int noise_batch,
                                      const float* next_scale, void* out, void* out_mod, int* error_flag, int B,
                                      int Cin, int Cout, int in_h, int in_w, int act, const w2e_tc2_config* cfg,
                                      void* stream) {
  W2E_CHECK_ARG(host_taps, "modconv_tc2_upblur: null taps");
  // separable factorisation of the 4x4 kernel (rank-1 check as in w2e_blur_act_nhwc)
  int bi = 0, bj = 0;
  float best = 0.f;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (fabsf(host_taps[i * 4 + j]) > best) best = fabsf(host_taps[i * 4 + j]), bi = i, bj = j;
  W2E_CHECK_ARG(best > 0.f, "modconv_tc2_upblur: zero kernel");
  float kv[4], kh[4];
  for (int i = 0; i < 4; ++i) kv[i] = host_taps[i * 4 + bj];
  for (int j = 0; j < 4; ++j) kh[j] = host_taps[bi * 4 + j] / host_taps[bi * 4 + bj];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (fabsf(host_taps[i * 4 + j] - kv[i] * kh[j]) > 1e-6f * best)
        return set_error(W2E_ERR_UNSUPPORTED, "modconv_tc2_upblur: the 4x4 kernel is not separable");
  FbArgs fb;
  for (int i = 0; i < 4; ++i) { fb.fv[i] = kv[3 - i]; fb.fh[i] = kh[3 - i]; }
  return run_tc2(xs, w, out_scale, bias, noise, noise_w, noise_batch, next_scale, out, out_mod, error_flag, B, Cin, Cout,
                 in_h, in_w, 1, act, nullptr, cfg, stream, false, &fb);
}
