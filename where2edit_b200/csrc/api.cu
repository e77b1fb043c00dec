// Library-wide plumbing of libw2e: error reporting, version, device query.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace w2e {

char* last_error_buffer() {
  static thread_local char buf[512] = "";
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;  // B200; do not cache a failed query
  }
  return cached;
}

}  // namespace w2e

extern "C" int w2e_version(void) { return 100; }

extern "C" const char* w2e_last_error_string(void) { return w2e::last_error_buffer(); }

extern "C" int w2e_device_info(int* sm_count_out, int* cc_major, int* cc_minor) {
  int dev = 0;
  W2E_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  W2E_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (sm_count_out) *sm_count_out = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return W2E_OK;
}
