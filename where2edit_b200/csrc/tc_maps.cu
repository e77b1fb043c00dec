// Tensor-map (TMA descriptor) helpers shared by the tcgen05 kernels, and the device capability query.
#include "tc_ptx.cuh"

namespace w2e {

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

static int make_map(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle sw,
                    CUtensorMapL2promotion promo, const char* what) {
  EncodeTiledFn fn = tensor_map_encoder();
  if (!fn) return set_error(W2E_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, dtype, (cuuint32_t)rank, const_cast<void*>(base), reinterpret_cast<const cuuint64_t*>(dims),
                  reinterpret_cast<const cuuint64_t*>(strides_bytes), reinterpret_cast<const cuuint32_t*>(box), estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(W2E_ERR_CUDA, "cuTensorMapEncodeTiled (%s) failed with CUresult %d", what, (int)r);
  return W2E_OK;
}

static CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
}

int make_bf16_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int row_bytes) {
  return make_map(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swizzle_for(row_bytes),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "bf16");
}

int make_f32_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box) {
  return make_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "fp32");
}

// fp32 operand map of the tf32 tensor-core mode: same byte geometry as the bf16 operand maps (128B / 64B rows)
int make_f32_swizzled_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, int row_bytes) {
  return make_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, swizzle_for(row_bytes),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "fp32 operand");
}

}  // namespace w2e

using namespace w2e;

extern "C" int w2e_modconv_tc_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return (major == 10 && tensor_map_encoder() != nullptr) ? 1 : 0;
}
