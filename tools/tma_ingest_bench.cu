// Microbenchmark: how fast can an SM ingest L2-resident data through TMA into shared memory?
// The weight-ring kernels of csrc/modconv_tc2.cu re-fetch their weight blocks per pixel tile and were measured at one
// 16 KB block per ~385 cycles.  Round-2 finding (first version of this file): a single thread that waits for block i and
// then issues block i+stages sees a CONSTANT ~326 cycles per block for 8 / 16 / 32 KB blocks, 2..8 blocks in flight,
// 1..148 CTAs -- i.e. the limit is per REQUEST, not per byte.  This version separates the candidates:
//   ring1  : one issuing thread, wait(i) then issue(i + stages)                (the production pattern)
//   ring2  : two issuing threads in different warps, each with its own ring    (does the request rate double?)
//   batch  : issue `stages` requests back to back, wait for all, repeat       (are requests pipelined at all?)
//   tensor : ring1 with cp.async.bulk.tensor.3d boxes {64, rows, 1} of a [9][256][512] bf16 weight tensor
// and prints the cycles spent inside the issue instruction itself.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/tma_ingest_bench tools/tma_ingest_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

struct Args {
  const uint8_t* src;
  size_t region_bytes, cta_stride;
  int block_bytes, stages, iters, mode;   // mode 0 ring1, 1 ring2, 2 batch, 3 tensor
  int rows;                               // tensor mode: box rows (block_bytes = rows * 128)
  long long* out;                         // [ctas][2]: total cycles, cycles inside issue instructions
};

__global__ void __launch_bounds__(128, 1) ingest(const __grid_constant__ CUtensorMap map, const __grid_constant__ Args A) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[2][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int r = 0; r < 2; ++r)
      for (int s = 0; s < A.stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[r][s])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const int nrings = A.mode == 1 ? 2 : 1;
  if (lane == 0 && warp < nrings) {
    const int ring = warp;
    uint8_t* buf = smem + (size_t)ring * A.stages * A.block_bytes;
    const uint8_t* base = A.src + (size_t)blockIdx.x * A.cta_stride + (size_t)ring * (A.region_bytes / 2);
    const int nblocks = (int)(A.region_bytes / nrings / A.block_bytes);
    long long issue_cycles = 0;
    auto issue = [&](int i) {
      const int s = i % A.stages;
      const long long c0 = clock64();
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[ring][s])), "r"(A.block_bytes) : "memory");
      if (A.mode == 3) {
        const int blk = i % (9 * 8 * (256 / A.rows));      // (tap, k chunk, row block) of the weight tensor
        const int t = blk % 9, kc = (blk / 9) % 8, rb = blk / 72;
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(buf + (size_t)s * A.block_bytes)), "l"(reinterpret_cast<uint64_t>(&map)),
                     "r"(smem_u32(&full[ring][s])), "r"(kc * 64), "r"(rb * A.rows), "r"(t) : "memory");
      } else {
        const uint8_t* g = base + (size_t)(i % nblocks) * A.block_bytes;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(buf + (size_t)s * A.block_bytes)),
                     "l"(g), "r"(A.block_bytes), "r"(smem_u32(&full[ring][s]))
                     : "memory");
      }
      issue_cycles += clock64() - c0;
    };
    long long t0, t1;
    if (A.mode == 2) {
      t0 = clock64();
      for (int i = 0; i < A.iters; i += A.stages) {
        for (int s = 0; s < A.stages; ++s) issue(i + s);
        for (int s = 0; s < A.stages; ++s) wait_bar(&full[ring][s], (uint32_t)((i / A.stages) & 1));
      }
      t1 = clock64();
    } else {
      for (int i = 0; i < A.stages; ++i) issue(i);
      t0 = clock64();
      for (int i = 0; i < A.iters; ++i) {
        wait_bar(&full[ring][i % A.stages], (uint32_t)((i / A.stages) & 1));
        issue(i + A.stages);
      }
      t1 = clock64();
      for (int i = A.iters; i < A.iters + A.stages; ++i) wait_bar(&full[ring][i % A.stages], (uint32_t)((i / A.stages) & 1));
    }
    if (ring == 0) { A.out[blockIdx.x * 2] = t1 - t0; A.out[blockIdx.x * 2 + 1] = issue_cycles; }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const size_t pool = (size_t)96 << 20;   // stays inside the 126 MB L2 after the first pass
  uint8_t* src;
  long long* d;
  cudaMalloc(&src, pool);
  cudaMemset(src, 1, pool);
  cudaMalloc(&d, 148 * 2 * sizeof(long long));
  cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const int iters = 4096;
  const char* names[4] = {"ring1 ", "ring2 ", "batch ", "tensor"};
  printf("mode   data    ctas block_KB stages | cycles/block  B/clk/SM  B/clk chip  cycles in issue/block\n");
  for (int mode = 0; mode < 4; ++mode)
    for (int shared = 1; shared >= 0; --shared)
      for (int ctas : {1, 148})
        for (int block_kb : {4, 8, 16, 32})
          for (int stages : {2, 4}) {
            if (mode == 3 && !shared) continue;
            const int nr = mode == 1 ? 2 : 1;
            if (block_kb * stages * nr > 192) continue;
            Args A;
            memset(&A, 0, sizeof(A));
            A.src = src;
            A.region_bytes = shared ? ((size_t)2304 << 10) : ((size_t)512 << 10);
            A.cta_stride = shared ? 0 : ((size_t)640 << 10);
            A.block_bytes = block_kb << 10; A.stages = stages; A.iters = iters; A.mode = mode; A.out = d;
            A.rows = (block_kb << 10) / 128;
            CUtensorMap map;
            memset(&map, 0, sizeof(map));
            if (mode == 3) {
              if (A.rows > 256 || !enc) continue;
              const cuuint64_t dims[3] = {512, 256, 9};
              const cuuint64_t strides[2] = {512 * 2, 256 * 512 * 2};
              const cuuint32_t box[3] = {64, (cuuint32_t)A.rows, 1};
              const cuuint32_t es[3] = {1, 1, 1};
              if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                continue;
            }
            for (int rep = 0; rep < 2; ++rep)   // first pass warms the L2
              ingest<<<ctas, 128, 200 * 1024>>>(map, A);
            long long h[296];
            cudaError_t e = cudaMemcpy(h, d, ctas * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            double mx = 0, iss = 0;
            for (int i = 0; i < ctas; ++i) { if ((double)h[2 * i] > mx) mx = (double)h[2 * i]; iss += (double)h[2 * i + 1]; }
            const double blocks = (double)iters * nr;          // blocks landed in this SM during the timed region
            const double cyc = mx / blocks, per_sm = (double)(block_kb << 10) / cyc;
            printf("%s %s %3d %2d %d | %8.1f %7.1f %8.0f %8.1f\n", names[mode], shared ? "shared " : "private", ctas, block_kb, stages,
                   cyc, per_sm, per_sm * ctas, iss / ctas / (iters + stages));
          }
  return 0;
}
