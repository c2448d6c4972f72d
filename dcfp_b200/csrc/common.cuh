// Shared helpers for the dcfp_b200 C-ABI library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "dcfp_b200.h"

namespace dcfp {

// thread-local error text behind dcfp_last_error()
void set_error(const char* fmt, ...);
// cudaGetLastError() -> return code (+ error text); also bumps the launch counter on success
int finish_launch(const char* what);
int cuda_fail(cudaError_t e, const char* what);

// opt a kernel into > 48 KB dynamic shared memory once per (kernel, device)
int ensure_smem(const void* func, int bytes);

constexpr int kNumSMs = 148;  // B200 (compile-time default for table sizing)
// SM count of the current device (queried once per device; persistent grids are sized with it)
int num_sms();

#define DCFP_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) {                    \
      ::dcfp::set_error(__VA_ARGS__); \
      return (code);                  \
    }                                 \
  } while (0)

// ---- PTX wrappers: mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP) ------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16; 16-B aligned both sides), completion on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// streaming 128-bit global load/store (read-once / write-once data)
__device__ __forceinline__ uint4 ldg_stream128(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

}  // namespace dcfp
