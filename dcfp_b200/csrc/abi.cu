// Error reporting, launch accounting and small ABI queries of libdcfp_b200.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "common.cuh"

namespace dcfp {
namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
std::mutex g_attr_mu;
struct AttrKey {
  const void* func;
  int device;
  int bytes;
};
AttrKey g_attr_seen[64];
int g_attr_n = 0;
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return static_cast<int>(e);
}

int finish_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int num_sms() {
  static std::atomic<int> cached[16] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return kNumSMs;
  int v = cached[dev].load(std::memory_order_relaxed);
  if (v > 0) return v;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = kNumSMs;
  cached[dev].store(v, std::memory_order_relaxed);
  return v;
}

int ensure_smem(const void* func, int bytes) {
  if (bytes <= 48 * 1024) return 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  std::lock_guard<std::mutex> lk(g_attr_mu);
  for (int i = 0; i < g_attr_n; ++i)
    if (g_attr_seen[i].func == func && g_attr_seen[i].device == dev && g_attr_seen[i].bytes >= bytes) return 0;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  if (g_attr_n < 64) g_attr_seen[g_attr_n++] = AttrKey{func, dev, bytes};
  return 0;
}
}  // namespace dcfp

extern "C" const char* dcfp_last_error(void) { return dcfp::g_err; }
extern "C" int dcfp_abi_version(void) { return DCFP_ABI_VERSION; }
extern "C" int64_t dcfp_launch_count(int reset) {
  return reset ? dcfp::g_launches.exchange(0) : dcfp::g_launches.load();
}
