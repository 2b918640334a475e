// Error plumbing and version for the C ABI (include/cnhead.h).
#include <string>

#include "common.cuh"

namespace cnh {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  cudaGetLastError();   // clear the sticky-free error state
  return (int)e;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

static long long* g_debug = nullptr;
long long* debug_buffer() { return g_debug; }

}  // namespace cnh

// not part of the public ABI: tools/ install a device buffer [grid][16] of int64 for stage timestamps
extern "C" void cnh_debug_set_buffer(void* p) { cnh::g_debug = static_cast<long long*>(p); }

extern "C" int cnh_version(void) { return CNH_VERSION; }
extern "C" const char* cnh_last_error(void) { return cnh::g_err; }

extern "C" int cnh_copy_async(void* dst, const void* src, size_t bytes, cnh_stream_t stream) {
  if (bytes == 0) return CNH_OK;
  if (dst == nullptr || src == nullptr) {
    cnh::set_error("copy_async: NULL pointer");
    return CNH_E_NULL;
  }
  const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? CNH_OK : cnh::cuda_fail(e, "copy_async");
}
