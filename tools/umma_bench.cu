// Microbenchmark: tcgen05.mma (kind::f16, cta_group::1, SS mode) issue/execute rate as a function
// of N, swizzle mode, SBO (canonical 8-row groups vs the 10-row haloed-tile pitch) and the number
// of independent accumulators.  One CTA per SM, operands are whatever is in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/umma_bench tools/umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc64(uint32_t hi, uint32_t addr) { return ((uint64_t)hi << 32) | ((addr & 0x3FFFF) >> 4); }

__global__ void __launch_bounds__(128, 1) k(int N, int row_bytes, int sbo_rows, int nacc, int iters, int kadv, int amode, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
    const uint32_t a_hi = (((uint32_t)(sbo_rows * row_bytes) >> 4) & 0x3FFF) | (1u << 14) | (layout << 29);
    const uint32_t b_hi = (((uint32_t)(8 * row_bytes) >> 4) & 0x3FFF) | (1u << 14) | (layout << 29);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    // 8 precomputed (d, a, b) triples cycled by an unrolled loop: the issue loop itself is ~3 instructions per MMA
    uint32_t dd[8]; uint64_t aa[8], bb[8];
    for (int i = 0; i < 8; ++i) {
      dd[i] = tmem + (uint32_t)((i % nacc) * N);
      aa[i] = desc64(a_hi, a0 + (amode ? (uint32_t)((i % 9) * row_bytes) : 0u) + (uint32_t)((i % kadv) * 32));
      bb[i] = desc64(b_hi, b0 + (uint32_t)((i % 4) * 8192) + (uint32_t)((i % kadv) * 32));
    }
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(dd[u]), "l"(aa[u]),
                     "l"(bb[u]), "r"(idesc), "r"(1));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)));
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4096;
  printf("N row_bytes sbo_rows nacc kadv amode ctas | issue_cyc/mma  total_cyc/mma  (math floor = N/2 cyc)\n");
  for (int ctas : {1, 148})
    for (int row_bytes : {128, 64})
      for (int sbo : {8, 10})
        for (int N : {32, 64, 128, 256})
          for (int nacc : {1, 2})
            for (int amode : {0, 1}) {
              if (N * nacc > 512) continue;
              const int kadv = row_bytes / 32;
              k<<<ctas, 128, 200 * 1024>>>(N, row_bytes, sbo, nacc, iters, kadv, amode, d);
              long long h[2];
              cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
              if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
              printf("%3d %3d %2d %d %d %d %3d | %7.1f %7.1f  (%d)\n", N, row_bytes, sbo, nacc, kadv, amode, ctas,
                     (double)h[0] / iters, (double)h[1] / iters, N / 2);
            }
  return 0;
}
