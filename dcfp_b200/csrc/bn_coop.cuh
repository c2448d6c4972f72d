// Fused BN(+ReLU) FORWARD in ONE cooperative launch per layer (sm_100a).
//
//   phase 1  every warp streams its [G px x 128 ch] boxes (cp.async.bulk.tensor.2d into a private ring, as K1 does) and sums
//            x and x^2 for the lane's 4 channels; the LAST ring-full of boxes stays in shared memory
//   reduce   CTA partials -> global [chunk][2][C] with plain stores; grid barrier; the channels are dealt over the CTAs,
//            each sums its channels' partials in a FIXED order in fp64 and writes mean / invstd / running statistics and
//            the fp32 (scale, shift) pair; grid barrier.  No atomics anywhere: the statistics are bit-reproducible.
//   phase 2  y = [relu](fma(x, scale, shift)), walking the boxes BACKWARDS: first the ones still resident in shared memory
//            (148 SMs x 192 KB = 28 MB: a whole 16.8 MB layer of c2 never leaves the chip between the two passes), then the
//            rest re-fetched most-recent-first, which is what the 126 MB L2 still holds.
//
// One launch instead of two (the launch / drain gap of a short kernel costs as much as moving a 17 MB layer) and at most
// one HBM read of x.  The grid is ONE tile per persistent CTA (<= #SMs), launched with cudaLaunchCooperativeKernel so
// that the grid barrier cannot deadlock.
#pragma once
#include "bn_common.cuh"
#include "k1_common.cuh"

namespace dcfp {
namespace {

constexpr int kCoopWarps = 16;
constexpr int kCoopSlab = 128;      // channels per warp row: 32 lanes x 4
constexpr int kCoopBoxBytes = 4096;  // [G px][128 ch]: G = 8 (fp32), 16 (bf16)
constexpr int kCoopStages = 3;       // ring per warp: 12 KB -> 192 KB per CTA stay resident after phase 1

struct CoopArgs {
  alignas(64) CUtensorMap map;  // x viewed as [rows][cols], box [G][128]
  void* y;
  float* partial;     // [n_chunks][2][cols] workspace (fully overwritten)
  unsigned* barrier;  // grid barrier flags (one per CTA), zero on entry
  BnFinal fin;
  int32_t rows, cols;  // the view: cols = C, or 128 with rows = N*h*w / 2 for a 64-channel layer (pixel-pair rows)
  int32_t fold2;
  int32_t spc;  // slabs per CTA (power of two <= 16); warps with equal (warp % spc) share a slab
  int32_t n_slab_groups, px_per_chunk, n_chunks;
};

// Grid barrier without same-address atomics (148 atomicAdds on one counter serialise for ~3 us in the L2 slice): every
// CTA publishes its epoch in a flag of its own, warp 0 of every CTA polls all flags.  Needs all CTAs co-resident
// (cooperative launch) and flags zero on entry; epochs 1, 2, ... within a launch.
__device__ __forceinline__ void grid_barrier(unsigned* flags, unsigned epoch) {
  __syncthreads();
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(epoch) : "memory");
    }
    bool ok;
    do {
      ok = true;
      for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
        ok = ok && v >= epoch;
      }
      ok = __all_sync(0xffffffffu, ok);
    } while (!ok);
  }
  __syncthreads();
}

template <typename T>
struct CoopRow;  // the lane's 4 channels of one box row: load as fp32, store to global
template <>
struct CoopRow<float> {
  static constexpr int kRowBytes = kCoopSlab * 4, kLaneBytes = 16;
  __device__ static __forceinline__ float4 load(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
  }
  __device__ static __forceinline__ void store(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
};
template <>
struct CoopRow<__nv_bfloat16> {
  static constexpr int kRowBytes = kCoopSlab * 2, kLaneBytes = 8;
  __device__ static __forceinline__ float4 load(uint32_t addr) {
    unsigned lo, hi;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr));
    return make_float4(__uint_as_float(lo << 16), __uint_as_float(lo & 0xffff0000u), __uint_as_float(hi << 16),
                       __uint_as_float(hi & 0xffff0000u));
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float4& v) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const unsigned*>(&a), *reinterpret_cast<const unsigned*>(&b));
  }
};

template <typename T, bool RELU>
__global__ void __launch_bounds__(kCoopWarps * 32, 1) bn_forward_coop_kernel(const __grid_constant__ CoopArgs A) {
  constexpr int G = kCoopBoxBytes / CoopRow<T>::kRowBytes;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // [warp][stage] boxes | [warp][lane][8] fp32 partial sums | [warp][stage] mbarriers
  float* red = reinterpret_cast<float*>(smem + kCoopWarps * kCoopStages * kCoopBoxBytes);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(red + kCoopWarps * 32 * 8);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t my_bufs = smem_u32(smem + static_cast<size_t>(warp) * kCoopStages * kCoopBoxBytes);
  const uint32_t my_bars = smem_u32(&bars[warp * kCoopStages]);
  if (lane < kCoopStages) mbar_init(my_bars + lane * 8, 1);
  mbar_fence_init();
  __syncwarp();

  // this CTA's tile = (slab group, chunk of rows); this warp = (slab, phase)
  const int tile = blockIdx.x;
  const int sg = tile / A.n_chunks, chunk = tile - sg * A.n_chunks;
  const int spc = A.spc, phases = kCoopWarps / spc, phase = warp / spc;
  const int col0 = (sg * spc + warp % spc) * kCoopSlab;
  const int c0v = col0 + lane * 4;          // column of the view
  const bool lane_on = c0v < A.cols;        // cols % 4 == 0
  const int c0 = A.fold2 ? (c0v & 63) : c0v;  // channel
  const int p_begin = chunk * A.px_per_chunk;
  const int p_end = min(p_begin + A.px_per_chunk, A.rows);
  const int n_groups = (p_end - p_begin + G - 1) / G;
  const int n_my = col0 < A.cols ? (n_groups - phase + phases - 1) / phases : 0;
  const uint64_t pol_keep = policy_evict_normal(), pol_last = policy_evict_first();
  unsigned parity_bits = 0;
  auto box_px = [&](int it) { return p_begin + (phase + it * phases) * G; };
  auto issue = [&](int it, uint64_t policy) {  // box `it` -> stage it % kCoopStages
    if (lane == 0) {
      const int st = it % kCoopStages;
      mbar_expect_tx(my_bars + st * 8, kCoopBoxBytes);
      tma_load_2d(my_bufs + st * kCoopBoxBytes, &A.map, col0, box_px(it), my_bars + st * 8, policy);
    }
  };
  auto wait = [&](int it) {
    const int st = it % kCoopStages;
    mbar_wait(my_bars + st * 8, (parity_bits >> st) & 1u);
    parity_bits ^= 1u << st;
  };
  const uint32_t lane_off = static_cast<uint32_t>(lane * CoopRow<T>::kLaneBytes);

  // ---- phase 1: sum x, sum x^2 (rows past the tensor arrive as zeros) ------------------------------------------------
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  for (int it = 0; it < min(kCoopStages, n_my); ++it) issue(it, pol_keep);
  for (int it = 0; it < n_my; ++it) {
    wait(it);
    const uint32_t box = my_bufs + (it % kCoopStages) * kCoopBoxBytes + lane_off;
#pragma unroll
    for (int e = 0; e < G; ++e) {
      const float4 v = CoopRow<T>::load(box + e * CoopRow<T>::kRowBytes);
      s1[0] += v.x, s1[1] += v.y, s1[2] += v.z, s1[3] += v.w;
      s2[0] = fmaf(v.x, v.x, s2[0]), s2[1] = fmaf(v.y, v.y, s2[1]), s2[2] = fmaf(v.z, v.z, s2[2]), s2[3] = fmaf(v.w, v.w, s2[3]);
    }
    __syncwarp();
    if (it + kCoopStages < n_my) issue(it + kCoopStages, pol_keep);  // the ring ends up holding the last boxes
  }
  // boxes [max(0, n_my - kCoopStages), n_my) are resident in their stages now

  // ---- CTA partial: warps sharing a slab combine through shared memory; one row of the global partial table per chunk ---
  {
    float4* r = reinterpret_cast<float4*>(red) + (warp * 32 + lane) * 2;
    r[0] = make_float4(s1[0], s1[1], s1[2], s1[3]);
    r[1] = make_float4(s2[0], s2[1], s2[2], s2[3]);
  }
  __syncthreads();
  if (phase == 0 && lane_on) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    for (int ph = 0; ph < phases; ++ph) {  // fixed order
      const float4* r = reinterpret_cast<const float4*>(red) + ((warp + ph * spc) * 32 + lane) * 2;
      const float4 u = r[0], v = r[1];
      a.x += u.x, a.y += u.y, a.z += u.z, a.w += u.w;
      b.x += v.x, b.y += v.y, b.z += v.z, b.w += v.w;
    }
    float* dst = A.partial + static_cast<size_t>(chunk) * 2 * A.cols + c0v;
    *reinterpret_cast<float4*>(dst) = a;
    *reinterpret_cast<float4*>(dst + A.cols) = b;
  }
  grid_barrier(A.barrier, 1u);

  // ---- statistics: channel c is finalised by CTA (c % gridDim.x), one warp per channel, fixed summation order ------------
  {
    const BnFinal& F = A.fin;
    float* coef = bn_coef(F.scratch, F.C);
    for (int c = blockIdx.x + warp * gridDim.x; c < F.C; c += gridDim.x * kCoopWarps) {
      double s = 0.0, q = 0.0;
      for (int ch = lane; ch < A.n_chunks; ch += 32) {
        const float* p = A.partial + static_cast<size_t>(ch) * 2 * A.cols;
        s += static_cast<double>(__ldcg(p + c)) + (A.fold2 ? static_cast<double>(__ldcg(p + c + 64)) : 0.0);
        q += static_cast<double>(__ldcg(p + A.cols + c)) + (A.fold2 ? static_cast<double>(__ldcg(p + A.cols + c + 64)) : 0.0);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (lane == 0) {
        const double mean = s * F.inv_m;
        const double var = fmax(q * F.inv_m - mean * mean, 0.0);
        const float mean_f = static_cast<float>(mean);
        const float invstd_f = static_cast<float>(rsqrt(var + static_cast<double>(F.eps)));
        const float scale = __fmul_rn(F.gamma[c], invstd_f);
        coef[c] = scale;
        coef[F.C + c] = __fmaf_rn(-mean_f, scale, F.beta[c]);
        F.mean[c] = mean_f;
        F.invstd[c] = invstd_f;
        if (F.running_mean != nullptr) {
          F.running_mean[c] = static_cast<float>((1.0 - F.momentum) * F.running_mean[c] + F.momentum * mean);
          F.running_var[c] = static_cast<float>((1.0 - F.momentum) * F.running_var[c] + F.momentum * var * F.unbias);
        }
      }
    }
  }
  grid_barrier(A.barrier, 2u);

  // ---- phase 2: normalise (+ReLU), last box first ------------------------------------------------------------------------
  float sc[4] = {0.f, 0.f, 0.f, 0.f}, sf[4] = {0.f, 0.f, 0.f, 0.f};
  if (lane_on) {
    const float* coef = bn_coef(A.fin.scratch, A.fin.C);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sc[j] = __ldcg(coef + c0 + j);
      sf[j] = __ldcg(coef + A.fin.C + c0 + j);
    }
  }
  T* y = reinterpret_cast<T*>(A.y);
  const int first_resident = max(0, n_my - kCoopStages);
  for (int it = n_my - 1; it >= 0; --it) {
    if (it < first_resident) wait(it);  // re-fetched below, kCoopStages boxes ahead
    const uint32_t box = my_bufs + (it % kCoopStages) * kCoopBoxBytes + lane_off;
    const int p = box_px(it);
    float4 v[G];
#pragma unroll
    for (int e = 0; e < G; ++e) v[e] = CoopRow<T>::load(box + e * CoopRow<T>::kRowBytes);
    __syncwarp();
    if (it - kCoopStages >= 0) issue(it - kCoopStages, pol_last);  // the stage just read is free
    if (lane_on) {
#pragma unroll
      for (int e = 0; e < G; ++e) {
        if (p + e < p_end) {
          float4 z;
          z.x = __fmaf_rn(v[e].x, sc[0], sf[0]);
          z.y = __fmaf_rn(v[e].y, sc[1], sf[1]);
          z.z = __fmaf_rn(v[e].z, sc[2], sf[2]);
          z.w = __fmaf_rn(v[e].w, sc[3], sf[3]);
          if (RELU) {  // NaN passes through, as torch.relu
            z.x = z.x < 0.f ? 0.f : z.x;
            z.y = z.y < 0.f ? 0.f : z.y;
            z.z = z.z < 0.f ? 0.f : z.z;
            z.w = z.w < 0.f ? 0.f : z.w;
          }
          CoopRow<T>::store(y + static_cast<size_t>(p + e) * A.cols + c0v, z);
        }
      }
    }
  }
}

// x viewed as [rows][cols] row-major, box [G rows][128 cols], no swizzle
inline int coop_make_map(CUtensorMap* map, const void* base, int dtype, long long rows, long long cols) {
  EncodeTiledFn enc = encode_tiled_fn();
  DCFP_REQUIRE(enc != nullptr, DCFP_EUNSUPPORTED, "bn_forward: cuTensorMapEncodeTiled is not available in this driver");
  const size_t es = dtype == DCFP_F32 ? 4 : 2;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * es};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kCoopSlab), static_cast<cuuint32_t>(kCoopBoxBytes / (kCoopSlab * es))};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(map, dtype == DCFP_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                         const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCFP_REQUIRE(r == CUDA_SUCCESS, DCFP_EINVAL, "bn_forward: cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld cols=%lld",
               static_cast<int>(r), rows, cols);
  return 0;
}

constexpr size_t kCoopSmem = static_cast<size_t>(kCoopWarps) * kCoopStages * kCoopBoxBytes + kCoopWarps * 32 * 8 * sizeof(float) +
                             kCoopWarps * kCoopStages * 8 + 1024;

// bytes of the (unzeroed, reusable) workspace one cooperative forward needs: the [chunks][2][cols] partial table
inline size_t coop_workspace_bytes(int C) { return static_cast<size_t>(kNumSMs * 2) * 2 * std::max(C, 128) * sizeof(float); }

template <typename T>
int coop_forward(const dcfp_bn_desc* d, const BnFinal& fin, cudaStream_t stream) {
  constexpr int G = kCoopBoxBytes / CoopRow<T>::kRowBytes;
  CoopArgs A{};
  const long long M = static_cast<long long>(d->N) * d->h * d->w;
  A.fold2 = (d->C == 64 && M % 2 == 0) ? 1 : 0;
  A.rows = static_cast<int32_t>(A.fold2 ? M / 2 : M);
  A.cols = A.fold2 ? 128 : d->C;
  int rc = coop_make_map(&A.map, d->x, d->dtype, A.rows, A.cols);
  if (rc) return rc;
  const int sms = num_sms();
  const int n_slabs = (A.cols + kCoopSlab - 1) / kCoopSlab;
  int spc = 1;
  while (spc < kCoopWarps && spc < n_slabs) spc <<= 1;
  A.spc = spc;
  A.n_slab_groups = (n_slabs + spc - 1) / spc;
  const int chunks = std::max(1, sms / A.n_slab_groups);
  long long px = (static_cast<long long>(A.rows) + chunks - 1) / chunks;
  px = std::max<long long>((px + G - 1) / G * G, G);
  A.px_per_chunk = static_cast<int32_t>(px);
  A.n_chunks = static_cast<int32_t>((A.rows + px - 1) / px);
  const size_t need = static_cast<size_t>(A.n_chunks) * 2 * A.cols * sizeof(float);
  DCFP_REQUIRE(d->workspace != nullptr && static_cast<size_t>(d->workspace_bytes) >= need, DCFP_EINVAL,
               "bn_forward: workspace of %zu bytes needed (dcfp_bn_workspace_bytes), got %lld", need, static_cast<long long>(d->workspace_bytes));
  A.y = d->y;
  A.partial = static_cast<float*>(d->workspace);
  A.barrier = bn_flags(d->scratch, d->C);
  DCFP_REQUIRE(A.n_slab_groups * A.n_chunks <= kBnMaxCtas, DCFP_ETOOBIG, "bn_forward: grid larger than %d CTAs", kBnMaxCtas);
  A.fin = fin;
  void (*kern)(CoopArgs) = d->relu ? bn_forward_coop_kernel<T, true> : bn_forward_coop_kernel<T, false>;
  rc = ensure_smem(reinterpret_cast<const void*>(kern), static_cast<int>(kCoopSmem));
  if (rc) return rc;
  void* params[] = {&A};
  const cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(A.n_slab_groups * A.n_chunks),
                                                    dim3(kCoopWarps * 32), params, kCoopSmem, stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchCooperativeKernel(bn_forward_coop)");
  return finish_launch("bn_forward_coop");
}

}  // namespace
}  // namespace dcfp
