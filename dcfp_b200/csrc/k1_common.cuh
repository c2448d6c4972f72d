// K1 shared device / host helpers: TMA tensor-tile copy, packed fp32x2 arithmetic, bf16 widening, tensor-map encoder.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace dcfp {
namespace {

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], "
      "[%4], %5;" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar), "l"(policy)
      : "memory");
}

// ---- packed fp32x2 helpers (Blackwell FADD2 / FFMA2 / FMUL2) ------------------------------------
using f2 = unsigned long long;
__device__ __forceinline__ f2 pack2(float lo, float hi) {
  return static_cast<f2>(__float_as_uint(lo)) | (static_cast<f2>(__float_as_uint(hi)) << 32);
}
__device__ __forceinline__ float lo2(f2 v) { return __uint_as_float(static_cast<unsigned>(v)); }
__device__ __forceinline__ float hi2(f2 v) { return __uint_as_float(static_cast<unsigned>(v >> 32)); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// 128-bit group -> pairs of fp32 values (fp32: 2 pairs = 4 px; bf16: 4 pairs = 8 px)
template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static constexpr int kPairs = 2;
  __device__ static __forceinline__ void unpack(const uint4& r, f2* v) {
    v[0] = static_cast<f2>(r.x) | (static_cast<f2>(r.y) << 32);
    v[1] = static_cast<f2>(r.z) | (static_cast<f2>(r.w) << 32);
  }
};
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int kPairs = 4;
  __device__ static __forceinline__ f2 widen(unsigned w) {  // two bf16 -> two fp32 (exact)
    return static_cast<f2>(w << 16) | (static_cast<f2>(w & 0xffff0000u) << 32);
  }
  __device__ static __forceinline__ void unpack(const uint4& r, f2* v) {
    v[0] = widen(r.x);
    v[1] = widen(r.y);
    v[2] = widen(r.z);
    v[3] = widen(r.w);
  }
};


using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

}  // namespace
}  // namespace dcfp
