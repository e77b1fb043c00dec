// Region-mask construction of the cluster-style mapper: the step that turns the
// mapper's per-pixel attention into the [B,1,S,S] attention_map consumed by Generator.forward's blend.
//
//  * w2e_cluster_assign      attention/run_attention.py:775-794 -- argmin_k ||(feature ++ x-pos ++ y-pos) - centre_k||^2
//                            per pixel, then nearest resize of the ids to the attention resolution.  The reference
//                            broadcasts a [B*h*h, K, D] difference tensor (189 MB per sample at 64x64, K=20, D=576,
//                            utils.py:244-263); here a thread owns a pixel, centres sit in shared memory and the
//                            distances never leave registers.
//  * w2e_region_mask_fwd     run_attention.py:852-884 -- per-cluster mean attention (the reference loops over
//                            B*clusters boolean masks, one sync each), loss_reg = sum relu(mean - margin) / B,
//                            loss_tv = mse(each, same), threshold (straight-through), 5x5 gaussian blur with reflect
//                            padding (torchvision gaussian_blur).  One CTA per sample, fixed reduction order.
//  * w2e_region_mask_bwd     autograd of the above w.r.t. the per-pixel attention.
#include "common.cuh"

namespace w2e {

// ------------------------------------------------------------------------------------------ cluster assignment
// Squared differences are fp32 (the reference's terms); they are summed in fp32 over groups of kAssignGroup
// channels and the group sums are accumulated in fp64, so the total carries less rounding error than one fp32
// ulp of the result -- any fp32 reduction order the reference's `.sum(-1)` may use lands within that.  The fp32
// rounding of the total is compared with first-index ties.
constexpr int kAssignKC = 8;        // clusters per pass (accumulators in registers)
constexpr int kAssignThreads = 128;
constexpr int kAssignGroup = 16;    // channels per fp32 partial sum

__global__ void __launch_bounds__(kAssignThreads)
cluster_assign_kernel(const float* __restrict__ feature, const float* __restrict__ centres, int* __restrict__ low,
                      int C, int h, int K, int pc) {
  extern __shared__ __align__(16) float ctr_s[];   // [kAssignKC][Dp], Dp = D rounded up to 4 (16-byte rows)
  const int D = C + 2 * pc;
  const int Dp = (D + 3) & ~3;
  const int b = blockIdx.y;
  const int hw = h * h;
  const int p = blockIdx.x * kAssignThreads + threadIdx.x;
  const bool live = p < hw;
  const int y = live ? p / h : 0, x = live ? p - (p / h) * h : 0;
  // arange(h).float() * 2 / float(h - 1) - 1   (run_attention.py:779-780)
  const float xpos = __fdiv_rn((float)x * 2.f, (float)(h - 1)) - 1.f;
  const float ypos = __fdiv_rn((float)y * 2.f, (float)(h - 1)) - 1.f;
  const float* f = feature + (int64_t)b * C * hw + (live ? p : 0);
  const int C4 = C & ~3;
  float best = 0.f;
  int best_k = -1;
  for (int k0 = 0; k0 < K; k0 += kAssignKC) {
    const int kc = min(kAssignKC, K - k0);
    __syncthreads();
    for (int e = threadIdx.x; e < kAssignKC * Dp; e += kAssignThreads) {
      const int k = e / Dp, m = e - k * Dp;
      ctr_s[e] = (k < kc && m < D) ? __ldg(centres + (int64_t)(k0 + k) * D + m) : 0.f;   // unused clusters: zeros
    }
    __syncthreads();
    if (!live) continue;
    double acc[kAssignKC];
    float part[kAssignKC];
#pragma unroll
    for (int k = 0; k < kAssignKC; ++k) { acc[k] = 0.0; part[k] = 0.f; }
    for (int c = 0; c < C4; c += 4) {
      const float v0 = __ldg(f + (int64_t)c * hw), v1 = __ldg(f + (int64_t)(c + 1) * hw);
      const float v2 = __ldg(f + (int64_t)(c + 2) * hw), v3 = __ldg(f + (int64_t)(c + 3) * hw);
#pragma unroll
      for (int k = 0; k < kAssignKC; ++k) {
        const float4 q = *reinterpret_cast<const float4*>(ctr_s + k * Dp + c);
        const float d0 = v0 - q.x, d1 = v1 - q.y, d2 = v2 - q.z, d3 = v3 - q.w;
        part[k] = fmaf(d0, d0, part[k]);
        part[k] = fmaf(d1, d1, part[k]);
        part[k] = fmaf(d2, d2, part[k]);
        part[k] = fmaf(d3, d3, part[k]);
      }
      if (((c + 4) & (kAssignGroup - 1)) == 0) {
#pragma unroll
        for (int k = 0; k < kAssignKC; ++k) { acc[k] += (double)part[k]; part[k] = 0.f; }
      }
    }
    for (int c = C4; c < C; ++c) {   // channel tail (C % 4)
      const float v = __ldg(f + (int64_t)c * hw);
#pragma unroll
      for (int k = 0; k < kAssignKC; ++k) {
        const float d = v - ctr_s[k * Dp + c];
        part[k] = fmaf(d, d, part[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < kAssignKC; ++k) { acc[k] += (double)part[k]; part[k] = 0.f; }
    for (int j = 0; j < 2 * pc; ++j) {
      const float v = j < pc ? xpos : ypos;
#pragma unroll
      for (int k = 0; k < kAssignKC; ++k) {
        const float d = v - ctr_s[k * Dp + C + j];
        part[k] = fmaf(d, d, part[k]);
      }
      if (((j + 1) & (kAssignGroup - 1)) == 0) {
#pragma unroll
        for (int k = 0; k < kAssignKC; ++k) { acc[k] += (double)part[k]; part[k] = 0.f; }
      }
    }
#pragma unroll
    for (int k = 0; k < kAssignKC; ++k) {
      if (k < kc) {
        const float dist = (float)(acc[k] + (double)part[k]);
        if (best_k < 0 || dist < best) { best = dist; best_k = k0 + k; }
      }
    }
  }
  if (live) low[(int64_t)b * hw + p] = best_k;
}

// ids[b, y, x] = b * K + low[b, floor(y * h / S), floor(x * h / S)]  (F.interpolate nearest, run_attention.py:794)
__global__ void __launch_bounds__(256)
cluster_resize_kernel(const int* __restrict__ low, int64_t* __restrict__ ids, int B, int h, int S, int K) {
  const int64_t n = (int64_t)B * S * S;
  const float scale = (float)h / (float)S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % S), y = (int)((i / S) % S), b = (int)(i / ((int64_t)S * S));
    const int sy = min((int)floorf((float)y * scale), h - 1), sx = min((int)floorf((float)x * scale), h - 1);
    ids[i] = (int64_t)b * K + __ldg(low + ((int64_t)b * h + sy) * h + sx);
  }
}

// --------------------------------------------------------------------------------------------- region mask
struct RegionTaps { float k[25]; };
constexpr int kRegionThreads = 256;
constexpr int kRegionWarps = kRegionThreads / 32;

__device__ __forceinline__ int reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

// One CTA per sample.  Phase 1: warp w owns clusters w, w+8, ...: lane-strided sums over the sample's pixels in a
// fixed order, shuffle tree.  Phase 2: same[p] = mean[id[p]], loss partials.  Phase 3: threshold + blur.
__global__ void __launch_bounds__(kRegionThreads)
region_mask_fwd_kernel(const float* __restrict__ each, const int64_t* __restrict__ ids, float* __restrict__ final_map,
                       float* same, float* __restrict__ stats, float* __restrict__ parts, int S, int K,
                       float threshold, float margin, const RegionTaps taps) {
  extern __shared__ float rs[];      // mean[K], cnt[K], red[32]
  float* mean_s = rs;
  float* cnt_s = rs + K;
  float* red = rs + 2 * K;
  const int b = blockIdx.x;
  const int N = S * S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* e_b = each + (int64_t)b * N;
  const int64_t* id_b = ids + (int64_t)b * N;
  const int64_t id0 = (int64_t)b * K;
  for (int k = warp; k < K; k += kRegionWarps) {
    float s = 0.f, c = 0.f;
    for (int p = lane; p < N; p += 32)
      if (__ldg(id_b + p) - id0 == k) { s += __ldg(e_b + p); c += 1.f; }
    s = warp_sum(s);
    c = warp_sum(c);
    if (lane == 0) {
      const float m = c > 0.f ? __fdiv_rn(s, c) : 0.f;
      mean_s[k] = m;
      cnt_s[k] = c;
      stats[((int64_t)b * K + k) * 2] = m;
      stats[((int64_t)b * K + k) * 2 + 1] = c;
    }
  }
  __syncthreads();
  float tv = 0.f;
  float* same_b = same + (int64_t)b * N;
  for (int p = threadIdx.x; p < N; p += kRegionThreads) {
    const int64_t lid = __ldg(id_b + p) - id0;
    const float v = (lid >= 0 && lid < K) ? mean_s[lid] : 1.f;   // torch.ones(...) where no cluster of this sample claims the pixel
    same_b[p] = v;
    const float d = __ldg(e_b + p) - v;
    tv = fmaf(d, d, tv);
  }
  tv = block_sum(tv, red);           // contains __syncthreads: same_b is visible to the whole CTA afterwards
  if (threadIdx.x == 0) {
    float reg = 0.f;                 // run_attention.py:865-871: sequential over the sample's clusters, empty ones skipped
    for (int k = 0; k < K; ++k)
      if (cnt_s[k] > 0.f) reg += fmaxf(mean_s[k] - margin, 0.f);
    parts[2 * b] = reg;
    parts[2 * b + 1] = tv;
  }
  float* out_b = final_map + (int64_t)b * N;
  for (int p = threadIdx.x; p < N; p += kRegionThreads) {
    const int y = p / S, x = p - y * S;
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
      const int yy = reflect(y + dy - 2, S);
#pragma unroll
      for (int dx = 0; dx < 5; ++dx) {
        const int xx = reflect(x + dx - 2, S);
        const float v = same_b[yy * S + xx];
        acc = fmaf(taps.k[dy * 5 + dx], v < threshold ? 0.f : v, acc);   // a - a.detach() == 0 below the threshold (:883)
      }
    }
    out_b[p] = acc;
  }
}

// losses[0] = loss_reg = (sum_b reg_b) / B ; losses[1] = loss_tv = (sum_b tv_b) / (B * S * S); sequential over b
__global__ void region_mask_finish_kernel(const float* __restrict__ parts, float* __restrict__ losses, int B, int N) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float reg = 0.f, tv = 0.f;
    for (int b = 0; b < B; ++b) { reg += parts[2 * b]; tv += parts[2 * b + 1]; }
    losses[0] = reg / (float)B;
    losses[1] = tv / ((float)B * (float)N);
  }
}

// Padded coordinates u in [-2, n+1] that reflect onto q: q itself, -q (q = 1, 2), 2(n-1)-q (q = n-3, n-2).
__device__ __forceinline__ int mirror_sources(int q, int n, int* u) {
  int c = 0;
  u[c++] = q;
  if (q >= 1 && q <= 2) u[c++] = -q;
  if (q >= n - 3 && q <= n - 2) u[c++] = 2 * (n - 1) - q;
  return c;
}

// g_each = d/d each of  <final, g_final> + g_reg * loss_reg + g_tv * loss_tv.  g_each doubles as the scratch that
// holds d/d same between the phases.
__global__ void __launch_bounds__(kRegionThreads)
region_mask_bwd_kernel(const float* __restrict__ g_final, const float* __restrict__ g_losses, const float* __restrict__ each,
                       const int64_t* __restrict__ ids, const float* __restrict__ same, const float* __restrict__ stats,
                       float* g_each, int B, int S, int K, float margin, const RegionTaps taps) {
  extern __shared__ float rs[];      // gm[K]
  float* gm = rs;
  const int b = blockIdx.x;
  const int N = S * S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float g_reg = g_losses ? __ldg(g_losses) : 0.f, g_tv = g_losses ? __ldg(g_losses + 1) : 0.f;
  const int64_t* id_b = ids + (int64_t)b * N;
  const int64_t id0 = (int64_t)b * K;
  float* gs = g_each + (int64_t)b * N;
  // phase A: adjoint of reflect-pad + 5x5 correlation (the threshold passes the gradient unchanged)
  for (int p = threadIdx.x; p < N; p += kRegionThreads) {
    float acc = 0.f;
    if (g_final) {
      const float* g_b = g_final + (int64_t)b * N;
      const int qy = p / S, qx = p - qy * S;
      int us[3], vs[3];
      const int nu = mirror_sources(qy, S, us), nv = mirror_sources(qx, S, vs);
      for (int iu = 0; iu < nu; ++iu)
        for (int iv = 0; iv < nv; ++iv) {
#pragma unroll
          for (int dy = 0; dy < 5; ++dy) {
            const int y = us[iu] + 2 - dy;
            if (y < 0 || y >= S) continue;
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) {
              const int x = vs[iv] + 2 - dx;
              if (x < 0 || x >= S) continue;
              acc = fmaf(taps.k[dy * 5 + dx], __ldg(g_b + y * S + x), acc);
            }
          }
        }
    }
    gs[p] = acc;
  }
  __syncthreads();
  // phase B: per-cluster gradient of the mean
  for (int k = warp; k < K; k += kRegionWarps) {
    float s = 0.f;
    for (int p = lane; p < N; p += 32)
      if (__ldg(id_b + p) - id0 == k) s += gs[p];
    s = warp_sum(s);
    if (lane == 0) {
      const float m = __ldg(stats + ((int64_t)b * K + k) * 2), c = __ldg(stats + ((int64_t)b * K + k) * 2 + 1);
      gm[k] = c > 0.f ? (s + (m - margin > 0.f ? g_reg / (float)B : 0.f)) / c : 0.f;
    }
  }
  __syncthreads();
  const float tv_scale = g_tv * 2.f / ((float)B * (float)N);
  for (int p = threadIdx.x; p < N; p += kRegionThreads) {
    const int64_t lid = __ldg(id_b + p) - id0;
    const float d = __ldg(each + (int64_t)b * N + p) - __ldg(same + (int64_t)b * N + p);
    gs[p] = ((lid >= 0 && lid < K) ? gm[lid] : 0.f) + tv_scale * d;
  }
}

static RegionTaps make_taps(const float* k25) {
  RegionTaps t;
  for (int i = 0; i < 25; ++i) t.k[i] = k25[i];
  return t;
}

}  // namespace w2e

using namespace w2e;

extern "C" int w2e_cluster_assign(const float* feature, const float* centres, int* low_ws, int64_t* ids, int B, int C, int h,
                                  int K, int pos_channels, int S, void* stream) {
  W2E_CHECK_ARG(feature && centres && low_ws && ids, "cluster_assign: null pointer");
  W2E_CHECK_ARG(B >= 0 && B <= 65535 && C > 0 && h >= 2 && K > 0 && pos_channels >= 0 && S > 0,
                "cluster_assign: bad shape (h >= 2, clusters > 0)");
  if (B == 0) return W2E_OK;
  const int D = C + 2 * pos_channels;
  const size_t smem = (size_t)kAssignKC * ((D + 3) & ~3) * sizeof(float);
  W2E_CHECK_ARG(smem <= 200 * 1024, "cluster_assign: feature dimension too large for the shared-memory centre tile");
  static PerDeviceOnce configured;
  if (!configured.done()) {
    W2E_CUDA_OK(cudaFuncSetAttribute(cluster_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured.mark();
  }
  cluster_assign_kernel<<<dim3((unsigned)ceil_div(h * h, kAssignThreads), (unsigned)B), kAssignThreads, smem,
                          (cudaStream_t)stream>>>(feature, centres, low_ws, C, h, K, pos_channels);
  W2E_LAUNCH_OK();
  const int64_t n = (int64_t)B * S * S;
  cluster_resize_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n, 256), 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      low_ws, ids, B, h, S, K);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_region_mask_fwd(const float* each, const int64_t* ids, const float* taps25, float* final_map, float* same,
                                   float* stats, float* parts_ws, float* losses, int B, int S, int K, float threshold,
                                   float margin, void* stream) {
  W2E_CHECK_ARG(each && ids && taps25 && final_map && same && stats && parts_ws && losses, "region_mask_fwd: null pointer");
  W2E_CHECK_ARG(B > 0 && S >= 3 && S <= 4096 && K > 0 && K <= 4096, "region_mask_fwd: bad shape (S >= 3, 0 < clusters <= 4096)");
  region_mask_fwd_kernel<<<(unsigned)B, kRegionThreads, (size_t)(2 * K + 32) * sizeof(float), (cudaStream_t)stream>>>(
      each, ids, final_map, same, stats, parts_ws, S, K, threshold, margin, make_taps(taps25));
  W2E_LAUNCH_OK();
  region_mask_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(parts_ws, losses, B, S * S);
  W2E_LAUNCH_OK();
  return W2E_OK;
}

extern "C" int w2e_region_mask_bwd(const float* g_final, const float* g_losses, const float* each, const int64_t* ids,
                                   const float* same, const float* stats, const float* taps25, float* g_each, int B, int S,
                                   int K, float margin, void* stream) {
  W2E_CHECK_ARG(each && ids && same && stats && taps25 && g_each, "region_mask_bwd: null pointer");
  W2E_CHECK_ARG(B > 0 && S >= 3 && S <= 4096 && K > 0 && K <= 4096, "region_mask_bwd: bad shape");
  region_mask_bwd_kernel<<<(unsigned)B, kRegionThreads, (size_t)K * sizeof(float), (cudaStream_t)stream>>>(
      g_final, g_losses, each, ids, same, stats, g_each, B, S, K, margin, make_taps(taps25));
  W2E_LAUNCH_OK();
  return W2E_OK;
}
