// Library-level C ABI: version, per-thread error string, launch counter.
#include <stdarg.h>
#include <atomic>

#include "common.cuh"

namespace mstr {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// SM count of the CURRENT device; cached per device ordinal (relaxed atomics: a race only repeats the query).
int sm_count() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  cached[dev].store(n, std::memory_order_relaxed);
  return n;
}

int ensure_dynamic_smem(const void* kernel, int bytes, unsigned long long* done_mask) {
  int dev = 0;
  MSTR_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (__atomic_load_n(done_mask, __ATOMIC_ACQUIRE) & bit) return MSTR_OK;
  MSTR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));   // idempotent
  __atomic_fetch_or(done_mask, bit, __ATOMIC_RELEASE);
  return MSTR_OK;
}

}  // namespace mstr

extern "C" __attribute__((visibility("default"))) int mstr_version(void) { return 1000; }
extern "C" __attribute__((visibility("default"))) const char* mstr_last_error_string(void) { return mstr::g_error; }
extern "C" __attribute__((visibility("default"))) int64_t mstr_launch_count(void) { return (int64_t)mstr::g_launches.load(); }

// Plain asynchronous host<->device copies on a caller-supplied stream (pinned host memory expected), used by
// pipeline.HostBatchPrefetcher so that the input copies are exactly one cudaMemcpyAsync each.
extern "C" __attribute__((visibility("default"))) int mstr_copy_h2d_async(void* dst_device, const void* src_host,
                                                                          int64_t bytes, mstr_stream_t stream) {
  MSTR_REQUIRE(dst_device && src_host && bytes > 0, MSTR_ERR_BAD_ARG, "mstr_copy_h2d_async: bad argument");
  MSTR_CUDA(cudaMemcpyAsync(dst_device, src_host, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return MSTR_OK;
}
extern "C" __attribute__((visibility("default"))) int mstr_copy_d2h_async(void* dst_host, const void* src_device,
                                                                          int64_t bytes, mstr_stream_t stream) {
  MSTR_REQUIRE(dst_host && src_device && bytes > 0, MSTR_ERR_BAD_ARG, "mstr_copy_d2h_async: bad argument");
  MSTR_CUDA(cudaMemcpyAsync(dst_host, src_device, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return MSTR_OK;
}
