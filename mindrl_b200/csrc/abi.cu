// abi.cu -- library-level entry points of mindrl_b200.h: errors, device selection, raw memory.
#include <stdarg.h>

#include "common.cuh"

namespace mrl {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static std::mutex mu;
  static int cached[64];  // 0 = not queried yet
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> g(mu);
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool first_use_on_device(const void *kernel) {
  static std::mutex mu;
  static std::unordered_map<const void *, uint64_t> seen;  // bit d: configured on device d
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  std::lock_guard<std::mutex> g(mu);
  uint64_t &m = seen[kernel];
  if (m & (1ull << dev)) return false;
  m |= 1ull << dev;
  return true;
}

int stream_wait_spin(cudaStream_t st) {
  static thread_local cudaEvent_t ev = nullptr;
  if (!ev) MRL_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  MRL_CUDA(cudaEventRecord(ev, st));
  cudaError_t e;
  while ((e = cudaEventQuery(ev)) == cudaErrorNotReady) {
  }
  MRL_CUDA(e);
  return MRL_OK;
}

}  // namespace mrl

using namespace mrl;

extern "C" int mrl_version(void) { return 100; }  // 0.1.0
extern "C" const char *mrl_last_error(void) { return g_err; }
extern "C" int64_t mrl_launch_count(void) { return g_launches.load(); }

extern "C" int mrl_device_count(int *count) {
  MRL_REQUIRE(count, "mrl_device_count: NULL output");
  *count = 0;
  MRL_CUDA(cudaGetDeviceCount(count));
  return MRL_OK;
}
extern "C" int mrl_set_device(int device) {
  MRL_CUDA(cudaSetDevice(device));
  return MRL_OK;
}
extern "C" int mrl_get_device(int *device) {
  MRL_REQUIRE(device, "mrl_get_device: NULL output");
  MRL_CUDA(cudaGetDevice(device));
  return MRL_OK;
}
extern "C" int mrl_malloc(void **dptr, size_t bytes) {
  MRL_REQUIRE(dptr, "mrl_malloc: NULL output");
  *dptr = nullptr;
  if (bytes == 0) return MRL_OK;
  MRL_CUDA(cudaMalloc(dptr, bytes));
  return MRL_OK;
}
extern "C" int mrl_free(void *dptr) {
  if (dptr) MRL_CUDA(cudaFree(dptr));
  return MRL_OK;
}
extern "C" int mrl_host_alloc(void **hptr, size_t bytes) {
  MRL_REQUIRE(hptr, "mrl_host_alloc: NULL output");
  *hptr = nullptr;
  if (bytes == 0) return MRL_OK;
  MRL_CUDA(cudaMallocHost(hptr, bytes));
  return MRL_OK;
}
extern "C" int mrl_host_free(void *hptr) {
  if (hptr) MRL_CUDA(cudaFreeHost(hptr));
  return MRL_OK;
}
extern "C" int mrl_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
  if (bytes) MRL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
  return MRL_OK;
}
extern "C" int mrl_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
  if (bytes) MRL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
  return MRL_OK;
}
extern "C" int mrl_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream) {
  if (bytes) MRL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
  return MRL_OK;
}
extern "C" int mrl_memset(void *dst, int value, size_t bytes, void *stream) {
  if (bytes) MRL_CUDA(cudaMemsetAsync(dst, value, bytes, as_stream(stream)));
  return MRL_OK;
}
extern "C" int mrl_stream_sync(void *stream) { return stream_wait_spin(as_stream(stream)); }
