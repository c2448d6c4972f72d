// K1 -- label-keyed segmented reduction over conv/BN feature maps (sm_100a).
//
//   S1[k][c] += sum_{pixels p of class k} v(p, c),      S2[k][c] += sum v(p, c)^2
//
// HBM-bound: every feature-map byte is read exactly once; there is no dense contraction, so no
// tensor cores.  Design:
//
//  * `dcfp_label_keys` nearest-down-samples the label map in registers ONCE per label resolution
//    into a compact uint8 class-key plane (and counts pixels per class).  Doing it inside the
//    reduction was measured at 27 % of all issued instructions, repeated by every 32-channel
//    group of every layer (profiles/r01_k1_notes.md).
//  * NCHW planes are pixel-contiguous, but the class key varies along pixels and is identical
//    across channels.  Each warp therefore pulls [32 channels x 128 B] boxes into shared memory
//    with ONE TMA tensor-tile copy (cp.async.bulk.tensor.2d, SWIZZLE_128B, mbarrier completion)
//    and reads them back TRANSPOSED -- lane == channel -- with conflict-free 128-bit loads.  The
//    class key of every pixel is then warp-uniform.
//  * With a warp-uniform key the reduction is a run-length accumulate in registers (packed
//    FADD2/FFMA2): while the key does not change, a1 += v, a2 += v*v; on a change the run is
//    flushed to a [K x 32] accumulator in shared memory.  A box whose 32/64 keys all equal the
//    current run's takes a branch-free path.
//  * Every warp owns a private multi-stage pipeline (its own mbarriers): the main loop has no
//    CTA-wide synchronisation.  A CTA covers 32 channels x one pixel chunk; its 4 warps take the
//    chunk's boxes round-robin.  Shared accumulators are per-warp copies (plain RMW) when K is
//    small, one CTA-wide copy updated with shared atomics otherwise.
//  * At the end the CTA adds its [K x 32] partials into the fp64 arena with coalesced RED.F64
//    (only classes it met): the cross-CTA / cross-image combine is done in fp64.
//  * `dcfp_class_stats_grouped` runs many resident layers in ONE launch: layer table and tensor
//    maps travel in kernel parameter space; each CTA binary-searches its layer.
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace dcfp {
namespace {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kBoxRowBytes = 128;                 // SWIZZLE_128B span
constexpr int kBoxBytes = 32 * kBoxRowBytes;      // one [32 channels x 128 B] box = 4 KB
constexpr int kGroups = kBoxRowBytes / 16;        // 128-bit groups per row (8)
constexpr int kPrivateAccMaxK = 24;               // per-warp accumulator copies up to this K
constexpr int kTargetBoxesPerChunk = 128;         // 128 x 4 KB = 512 KB of input per CTA

struct LayerDev {
  const uint8_t* keys;  // [N][HW] class keys at this layer's resolution (K == dropped)
  const float* scale;
  const float* shift;
  double* S1;
  double* S2;
  int32_t C, HW, n_cg;
  int32_t boxes_per_plane, n_boxes, boxes_per_chunk;
};

// Layer table + TMA descriptors in kernel parameter space (no H2D copy, no workspace).
template <int MAXL, int TENS>
struct GroupParams {
  alignas(64) CUtensorMap maps[MAXL * TENS];  // [layer][x, dy]
  LayerDev L[MAXL];
  int32_t tile_prefix[MAXL + 1];
  int32_t n_layers;
  int32_t K;
};
constexpr int kSmallGroup = 4;
constexpr int kBigGroupFwd = 160;  // 160 * (128 + 64) B  = 30.0 KB  (< 32 KB parameter space)
constexpr int kBigGroupBwd = 96;   //  96 * (256 + 64) B  = 30.0 KB

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

// Flush one finished run of one lane (= one channel) into the shared accumulators.  Out of line:
// it sits on the rare path (key change) and would otherwise be replicated in the unrolled loop.
template <bool SHARED_ACC>
__device__ __noinline__ void flush_run(float* acc1, float* acc2, unsigned* seen, unsigned cur, int lane, float a1, float a2) {
  if (SHARED_ACC) {
    atomicAdd(&acc1[cur * 32 + lane], a1);
    atomicAdd(&acc2[cur * 32 + lane], a2);
  } else {
    acc1[cur * 32 + lane] += a1;
    acc2[cur * 32 + lane] += a2;
  }
  if (lane == 0) seen[cur] = 1u;
}

// Run-length accumulator of one lane; everything except the partial sums is warp-uniform.
template <bool SHARED_ACC>
struct RunAcc {
  f2 s1a = 0, s1b = 0, s2a = 0, s2b = 0;  // packed partial sums of the current run
  unsigned cur, curw;
  float* acc1;
  float* acc2;
  unsigned* seen;
  int K, lane;

  __device__ __forceinline__ void flush() {
    if (cur < static_cast<unsigned>(K)) {
      const f2 t1 = add2(s1a, s1b), t2 = add2(s2a, s2b);
      flush_run<SHARED_ACC>(acc1, acc2, seen, cur, lane, lo2(t1) + hi2(t1), lo2(t2) + hi2(t2));
    }
    s1a = s1b = s2a = s2b = 0;
  }
  __device__ __forceinline__ void add_pair_fast(f2 p, f2 q) {  // 4 px of the current run
    s1a = add2(s1a, p);
    s1b = add2(s1b, q);
    s2a = fma2(p, p, s2a);
    s2b = fma2(q, q, s2b);
  }
  __device__ __forceinline__ void add_px(float v, unsigned key) {
    if (key != cur) {
      flush();
      cur = key;
      curw = key * 0x01010101u;
    }
    const f2 pv = pack2(v, 0.f);
    s1a = add2(s1a, pv);
    s2a = fma2(pv, pv, s2a);
  }
  // four consecutive pixels (two pairs) with packed keys `wv` (one byte each, warp-uniform)
  __device__ __forceinline__ void add4(f2 p, f2 q, unsigned wv) {
    if (wv == curw) {
      add_pair_fast(p, q);
    } else {
      add_px(lo2(p), wv & 0xffu);
      add_px(hi2(p), (wv >> 8) & 0xffu);
      add_px(lo2(q), (wv >> 16) & 0xffu);
      add_px(hi2(q), wv >> 24);
    }
  }
};

template <typename T, bool BWD, bool AFFINE>
__device__ __forceinline__ void load_group(uint32_t row, int g, int lane, f2 sc2, f2 sf2, f2* v) {
  const uint32_t off = static_cast<uint32_t>((g ^ (lane & 7)) << 4);  // SWIZZLE_128B: chunk ^= row % 8
  Elem<T>::unpack(lds128(row + off), v);
  if (BWD) {
    f2 d[Elem<T>::kPairs];
    Elem<T>::unpack(lds128(row + kBoxBytes + off), d);
#pragma unroll
    for (int q = 0; q < Elem<T>::kPairs; ++q) v[q] = mul2(d[q], fma2(v[q], sc2, sf2));
  } else if (AFFINE) {
#pragma unroll
    for (int q = 0; q < Elem<T>::kPairs; ++q) v[q] = fma2(v[q], sc2, sf2);
  }
}

template <typename T, bool BWD, bool AFFINE, bool SHARED_ACC, int STAGES>
__device__ __forceinline__ void process_tile(const LayerDev& L, const CUtensorMap* maps, const int K, const int tile,
                                             unsigned char* smem) {
  constexpr int kBoxPx = kBoxRowBytes / static_cast<int>(sizeof(T));  // 32 (fp32) / 64 (bf16)
  constexpr int kWords = kBoxPx / 4;                                  // packed key words per box
  constexpr int kTens = BWD ? 2 : 1;
  constexpr int kAccCopies = SHARED_ACC ? 1 : kWarps;
  constexpr int kPairs = Elem<T>::kPairs;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = tile / L.n_cg, cg = tile - chunk * L.n_cg;
  const int n_active = min(32, L.C - cg * 32);

  // ---- shared-memory carve-up: [boxes | accumulators | seen flags | mbarriers] ------------------
  unsigned char* bufs = smem;  // [kWarps][STAGES][kTens][kBoxBytes], 1024-B aligned
  float* acc1 = reinterpret_cast<float*>(smem + static_cast<size_t>(kWarps) * STAGES * kTens * kBoxBytes);
  float* acc2 = acc1 + kAccCopies * K * 32;
  unsigned* seen = reinterpret_cast<unsigned*>(acc2 + kAccCopies * K * 32);  // [K]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(seen + ((K + 1) & ~1));

  for (int i = tid; i < 2 * kAccCopies * K * 32 + K; i += kThreads) reinterpret_cast<unsigned*>(acc1)[i] = 0u;
  if (tid < kWarps * STAGES) mbar_init(smem_u32(&bars[tid]), 1);
  mbar_fence_init();
  __syncthreads();

  const int box_begin = chunk * L.boxes_per_chunk;
  const int box_end = min(box_begin + L.boxes_per_chunk, L.n_boxes);
  const uint64_t policy = policy_evict_first();
  const uint32_t my_bufs = smem_u32(bufs + static_cast<size_t>(warp) * STAGES * kTens * kBoxBytes);
  const uint32_t my_bars = smem_u32(&bars[warp * STAGES]);

  auto issue = [&](int it) {  // one elected lane arms the barrier and launches the tile copies
    const int box = box_begin + warp + it * kWarps;
    if (box >= box_end || lane != 0) return;
    const int stage = it % STAGES;
    const int n = box / L.boxes_per_plane;
    const int p0 = (box - n * L.boxes_per_plane) * kBoxPx;
    const uint32_t bar = my_bars + stage * 8;
    const uint32_t dst = my_bufs + stage * (kTens * kBoxBytes);
    mbar_expect_tx(bar, kTens * kBoxBytes);
    tma_load_2d(dst, maps, p0, n * L.C + cg * 32, bar, policy);
    if (BWD) tma_load_2d(dst + kBoxBytes, maps + 1, p0, n * L.C + cg * 32, bar, policy);
  };
  // packed class keys of the box's pixels 4*lane .. 4*lane+3 (lanes < kWords); K = "dropped"
  auto key_word = [&](int it) -> unsigned {
    const int box = box_begin + warp + it * kWarps;
    const unsigned dropped = static_cast<unsigned>(K) * 0x01010101u;
    if (box >= box_end || lane >= kWords) return dropped;
    const int n = box / L.boxes_per_plane;
    const int p = (box - n * L.boxes_per_plane) * kBoxPx + 4 * lane;
    if (p >= L.HW) return dropped;  // HW % 4 == 0: a word is entirely inside or outside the plane
    if (L.keys == nullptr) return 0u;
    return __ldg(reinterpret_cast<const unsigned*>(L.keys + static_cast<size_t>(n) * L.HW + p));
  };

  float sc = 1.f, sf = 0.f;
  if (lane < n_active) {
    if (L.scale) sc = L.scale[cg * 32 + lane];
    if (L.shift) sf = L.shift[cg * 32 + lane];
  }
  const f2 sc2 = pack2(sc, sc), sf2 = pack2(sf, sf);

  RunAcc<SHARED_ACC> ra;
  ra.K = K;
  ra.lane = lane;
  ra.cur = K;
  ra.curw = static_cast<unsigned>(K) * 0x01010101u;
  ra.acc1 = acc1 + (SHARED_ACC ? 0 : warp * K * 32);
  ra.acc2 = acc2 + (SHARED_ACC ? 0 : warp * K * 32);
  ra.seen = seen;

#pragma unroll
  for (int s = 0; s < STAGES; ++s) issue(s);
  unsigned lw = key_word(0);

  const int n_my = (box_end - box_begin - warp + kWarps - 1) / kWarps;  // boxes of this warp
  for (int it = 0; it < n_my; ++it) {
    const unsigned lw_next = key_word(it + 1);  // global load overlaps the wait below
    const int stage = it % STAGES;
    mbar_wait(my_bars + stage * 8, (it / STAGES) & 1);
    const uint32_t row = my_bufs + stage * (kTens * kBoxBytes) + lane * kBoxRowBytes;
    const bool uniform = __all_sync(0xffffffffu, lane >= kWords || lw == ra.curw);
    if (uniform) {  // the whole box continues the current run: branch-free
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        f2 v[kPairs];
        load_group<T, BWD, AFFINE>(row, g, lane, sc2, sf2, v);
#pragma unroll
        for (int h = 0; h < kPairs / 2; ++h) ra.add_pair_fast(v[2 * h], v[2 * h + 1]);
      }
    } else {
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        f2 v[kPairs];
        load_group<T, BWD, AFFINE>(row, g, lane, sc2, sf2, v);
#pragma unroll
        for (int h = 0; h < kPairs / 2; ++h)
          ra.add4(v[2 * h], v[2 * h + 1], __shfl_sync(0xffffffffu, lw, g * (kPairs / 2) + h));
      }
    }
    __syncwarp();
    issue(it + STAGES);  // refill the stage just consumed
    lw = lw_next;
  }
  ra.flush();
  __syncthreads();

  // ---- CTA partials -> fp64 arena (coalesced RED.F64; only classes this CTA met) ----------------
  for (int idx = tid; idx < K * 32; idx += kThreads) {
    const int k = idx >> 5, cl = idx & 31;
    if (seen[k] == 0u || cl >= n_active) continue;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int w = 0; w < kAccCopies; ++w) {
      s1 += acc1[w * K * 32 + idx];
      s2 += acc2[w * K * 32 + idx];
    }
    const size_t o = static_cast<size_t>(k) * L.C + cg * 32 + cl;
    atomicAdd(&L.S1[o], static_cast<double>(s1));
    atomicAdd(&L.S2[o], static_cast<double>(s2));
  }
}

template <typename T, bool BWD, bool AFFINE, bool SHARED_ACC, int STAGES, int MAXL>
__global__ void __launch_bounds__(kThreads, 4)
    class_stats_kernel(const __grid_constant__ GroupParams<MAXL, BWD ? 2 : 1> P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // dynamic shared memory is only guaranteed 16-B aligned; SWIZZLE_128B boxes need 1024 B
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tile = blockIdx.x;
  int lo = 0, hi = P.n_layers;  // largest l with tile_prefix[l] <= tile
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (P.tile_prefix[mid] <= tile) lo = mid;
    else hi = mid;
  }
  process_tile<T, BWD, AFFINE, SHARED_ACC, STAGES>(P.L[lo], &P.maps[lo * (BWD ? 2 : 1)], P.K, tile - P.tile_prefix[lo], smem);
}

// Generic path: any extent / alignment / layout (tiny 1x1..6x6 maps, odd crops, NHWC).  One
// thread per channel walks the pixels of one plane chunk; runs are flushed straight to the arena.
struct GenericLayer {
  const void* x;
  const void* dy;
  const uint8_t* keys;
  const float* scale;
  const float* shift;
  double* S1;
  double* S2;
  int32_t C, HW;
};
template <typename T, bool BWD>
__global__ void class_stats_generic_kernel(const GenericLayer L, const int K, const int nhwc, const int px_per_block) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.z;
  const int p_begin = blockIdx.y * px_per_block, p_end = min(p_begin + px_per_block, L.HW);
  if (c >= L.C) return;
  const float sc = L.scale ? L.scale[c] : 1.f, sf = L.shift ? L.shift[c] : 0.f;
  const T* x = reinterpret_cast<const T*>(L.x);
  const T* dy = reinterpret_cast<const T*>(L.dy);
  float a1 = 0.f, a2 = 0.f;
  unsigned cur = K;
  auto flush = [&]() {
    if (cur < static_cast<unsigned>(K)) {
      atomicAdd(&L.S1[static_cast<size_t>(cur) * L.C + c], static_cast<double>(a1));
      atomicAdd(&L.S2[static_cast<size_t>(cur) * L.C + c], static_cast<double>(a2));
    }
    a1 = a2 = 0.f;
  };
  for (int p = p_begin; p < p_end; ++p) {
    const unsigned k = L.keys ? L.keys[static_cast<size_t>(n) * L.HW + p] : 0u;
    if (k != cur) {
      flush();
      cur = k;
    }
    const size_t idx = nhwc ? (static_cast<size_t>(n) * L.HW + p) * L.C + c : (static_cast<size_t>(n) * L.C + c) * L.HW + p;
    float v = fmaf(static_cast<float>(x[idx]), sc, sf);
    if (BWD) v *= static_cast<float>(dy[idx]);
    a1 += v;
    a2 = fmaf(v, v, a2);
  }
  flush();
}

// ---- label keys: in-register legacy-`nearest` down-sampling + per-class pixel counts --------------
__device__ __forceinline__ int load_label(const void* label, int dtype, long long idx) {
  if (dtype == DCFP_LABEL_U8) return static_cast<const unsigned char*>(label)[idx];
  if (dtype == DCFP_LABEL_I32) return static_cast<const int*>(label)[idx];
  const long long v = static_cast<const long long*>(label)[idx];
  return (v < 0 || v > 0x7fffffffLL) ? -1 : static_cast<int>(v);
}

__global__ void __launch_bounds__(256) label_keys_kernel(const void* __restrict__ label, int label_dtype, int N, int H0, int W0,
                                                         int h, int w, int K, float sh, float sw, uint8_t* __restrict__ keys,
                                                         double* __restrict__ cnt) {
  __shared__ unsigned hist[256];
  hist[threadIdx.x] = 0u;
  __syncthreads();
  const long long total = static_cast<long long>(N) * h * w;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(t % w);
    const long long r = t / w;
    const int i = static_cast<int>(r % h), n = static_cast<int>(r / h);
    // ATen nearest_neighbor_compute_source_index: min(floor(dst * (float)(in/out)), in - 1)
    const int si = min(static_cast<int>(floorf(i * sh)), H0 - 1);
    const int sj = min(static_cast<int>(floorf(j * sw)), W0 - 1);
    const int l = load_label(label, label_dtype, (static_cast<long long>(n) * H0 + si) * W0 + sj);
    const unsigned k = (l < 0 || l >= K) ? K : l;
    keys[t] = static_cast<uint8_t>(k);
    if (cnt != nullptr && k < static_cast<unsigned>(K)) atomicAdd(&hist[k], 1u);
  }
  __syncthreads();
  if (cnt != nullptr && threadIdx.x < K && hist[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], static_cast<double>(hist[threadIdx.x]));
}

// ---- host side ------------------------------------------------------------------------------------
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

// [rows = N*C][cols = HW] view of an NCHW tensor, box = [32 rows][128 B], SWIZZLE_128B
int make_map(CUtensorMap* map, const void* base, int dtype, long long rows, long long cols) {
  EncodeTiledFn enc = encode_tiled_fn();
  DCFP_REQUIRE(enc != nullptr, DCFP_EUNSUPPORTED, "class_stats: cuTensorMapEncodeTiled is not available in this driver");
  const size_t es = dtype == DCFP_F32 ? 4 : 2;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * es};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBoxRowBytes / es), 32u};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(map, dtype == DCFP_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                         const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCFP_REQUIRE(r == CUDA_SUCCESS, DCFP_EINVAL, "class_stats: cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld cols=%lld",
               static_cast<int>(r), rows, cols);
  return 0;
}

size_t tile_smem_bytes(int K, bool bwd, bool shared_acc, int stages) {
  const int copies = shared_acc ? 1 : kWarps;
  return static_cast<size_t>(kWarps) * stages * (bwd ? 2 : 1) * kBoxBytes + static_cast<size_t>(2) * copies * K * 32 * 4 +
         static_cast<size_t>((K + 1) & ~1) * 4 + 8 * kWarps * stages + 1024 /* base alignment slack */;
}

int validate(const dcfp_layer_desc& d, int idx) {
  DCFP_REQUIRE(d.x && d.S1 && d.S2, DCFP_EINVAL, "class_stats[%d]: x/S1/S2 must be non-null", idx);
  DCFP_REQUIRE(d.N > 0 && d.C > 0 && d.h > 0 && d.w > 0, DCFP_EINVAL, "class_stats[%d]: bad extent N=%d C=%d h=%d w=%d", idx,
               d.N, d.C, d.h, d.w);
  DCFP_REQUIRE(d.K >= 1 && d.K <= DCFP_MAX_CLASSES, DCFP_ETOOBIG, "class_stats[%d]: K=%d outside [1,%d]", idx, d.K,
               DCFP_MAX_CLASSES);
  DCFP_REQUIRE(d.dtype == DCFP_F32 || d.dtype == DCFP_BF16, DCFP_EINVAL, "class_stats[%d]: unknown dtype %d", idx, d.dtype);
  DCFP_REQUIRE(d.layout == DCFP_NCHW || d.layout == DCFP_NHWC, DCFP_EINVAL, "class_stats[%d]: unknown layout %d", idx, d.layout);
  DCFP_REQUIRE(d.keys != nullptr || d.K == 1, DCFP_EINVAL, "class_stats[%d]: keys == NULL requires K == 1", idx);
  DCFP_REQUIRE(static_cast<long long>(d.h) * d.w < (1LL << 30), DCFP_ETOOBIG, "class_stats[%d]: plane too large", idx);
  DCFP_REQUIRE(static_cast<long long>(d.N) * d.C < (1LL << 31), DCFP_ETOOBIG, "class_stats[%d]: too many planes", idx);
  return 0;
}

// the TMA path needs 16-B aligned planes and word-aligned key rows; everything else is generic
bool tiled_ok(const dcfp_layer_desc& d) {
  const size_t es = d.dtype == DCFP_F32 ? 4 : 2;
  const size_t plane = static_cast<size_t>(d.h) * d.w * es;
  if (d.layout != DCFP_NCHW) return false;
  if (plane % 16 != 0 || plane < 512) return false;
  if (reinterpret_cast<uintptr_t>(d.x) % 16 != 0) return false;
  if (d.dy && reinterpret_cast<uintptr_t>(d.dy) % 16 != 0) return false;
  if (d.keys && reinterpret_cast<uintptr_t>(d.keys) % 4 != 0) return false;
  return true;
}

template <typename T, bool BWD>
int launch_generic(const dcfp_layer_desc& d, cudaStream_t stream) {
  GenericLayer L{d.x, d.dy, d.keys, d.scale, d.shift, d.S1, d.S2, d.C, d.h * d.w};
  const int threads = 128;
  const int px_per_block = 256;
  dim3 grid((d.C + threads - 1) / threads, (L.HW + px_per_block - 1) / px_per_block, d.N);
  class_stats_generic_kernel<T, BWD><<<grid, threads, 0, stream>>>(L, d.K, d.layout == DCFP_NHWC, px_per_block);
  return finish_launch("class_stats_generic");
}

template <typename T, bool BWD, bool AFFINE, bool SHARED_ACC, int STAGES, int MAXL>
int launch_tiled(const GroupParams<MAXL, BWD ? 2 : 1>& P, int n_tiles, cudaStream_t stream) {
  const size_t smem = tile_smem_bytes(P.K, BWD, SHARED_ACC, STAGES);
  auto kern = class_stats_kernel<T, BWD, AFFINE, SHARED_ACC, STAGES, MAXL>;
  int rc = ensure_smem(reinterpret_cast<const void*>(kern), static_cast<int>(smem));
  if (rc) return rc;
  kern<<<n_tiles, kThreads, smem, stream>>>(P);
  return finish_launch("class_stats");
}

template <typename T, bool BWD, int MAXL>
int run_tiled(const dcfp_layer_desc* descs, const int* which, int n, int boxes_per_chunk, cudaStream_t stream) {
  constexpr int kTens = BWD ? 2 : 1;
  constexpr int kBoxPx = kBoxRowBytes / static_cast<int>(sizeof(T));
  const int K = descs[which[0]].K;
  GroupParams<MAXL, kTens> P;
  P.n_layers = n;
  P.K = K;
  P.tile_prefix[0] = 0;
  bool affine = false;
  for (int i = 0; i < n; ++i) {
    const dcfp_layer_desc& d = descs[which[i]];
    LayerDev& L = P.L[i];
    L.keys = d.keys;
    L.scale = d.scale;
    L.shift = d.shift;
    L.S1 = d.S1;
    L.S2 = d.S2;
    L.C = d.C;
    L.HW = d.h * d.w;
    L.n_cg = (d.C + 31) / 32;
    L.boxes_per_plane = (L.HW + kBoxPx - 1) / kBoxPx;
    L.n_boxes = L.boxes_per_plane * d.N;
    L.boxes_per_chunk = boxes_per_chunk;
    affine = affine || d.scale || d.shift;
    const long long tiles = static_cast<long long>((L.n_boxes + boxes_per_chunk - 1) / boxes_per_chunk) * L.n_cg;
    DCFP_REQUIRE(P.tile_prefix[i] + tiles < (1LL << 31), DCFP_ETOOBIG, "class_stats: too many tiles");
    P.tile_prefix[i + 1] = P.tile_prefix[i] + static_cast<int>(tiles);
    int rc = make_map(&P.maps[i * kTens], d.x, d.dtype, static_cast<long long>(d.N) * d.C, L.HW);
    if (rc == 0 && BWD) rc = make_map(&P.maps[i * kTens + 1], d.dy, d.dtype, static_cast<long long>(d.N) * d.C, L.HW);
    if (rc) return rc;
  }
  const int n_tiles = P.tile_prefix[n];
  if (n_tiles == 0) return 0;
  if (K > kPrivateAccMaxK) {  // one CTA-wide accumulator copy, shared atomics; [K x 32 x 2] floats
    if (BWD || affine) return launch_tiled<T, BWD, true, true, 2, MAXL>(P, n_tiles, stream);
    return launch_tiled<T, BWD, false, true, 2, MAXL>(P, n_tiles, stream);
  }
  if (BWD || affine) return launch_tiled<T, BWD, true, false, 2, MAXL>(P, n_tiles, stream);
  return launch_tiled<T, BWD, false, false, 2, MAXL>(P, n_tiles, stream);
}

int run(const dcfp_layer_desc* descs, int n_layers, cudaStream_t stream) {
  DCFP_REQUIRE(descs != nullptr && n_layers > 0, DCFP_EINVAL, "class_stats: no layers");
  DCFP_REQUIRE(n_layers <= DCFP_MAX_GROUP_LAYERS, DCFP_ETOOBIG, "class_stats: %d layers > %d per call", n_layers,
               DCFP_MAX_GROUP_LAYERS);
  const int K = descs[0].K, dtype = descs[0].dtype;
  const bool bwd = descs[0].dy != nullptr;
  int tiled[DCFP_MAX_GROUP_LAYERS];
  int n_tiled = 0;
  long long total_boxes = 0;  // boxes x channel groups over the whole call
  const int box_px = kBoxRowBytes / (dtype == DCFP_F32 ? 4 : 2);
  for (int i = 0; i < n_layers; ++i) {
    const dcfp_layer_desc& d = descs[i];
    int rc = validate(d, i);
    if (rc) return rc;
    DCFP_REQUIRE(d.K == K && d.dtype == dtype && (d.dy != nullptr) == bwd, DCFP_EINVAL,
                 "class_stats[%d]: K / dtype / functor differ inside one group", i);
    if (tiled_ok(d)) {
      tiled[n_tiled++] = i;
      total_boxes += ((static_cast<long long>(d.h) * d.w + box_px - 1) / box_px) * d.N * ((d.C + 31) / 32);
    } else {
      if (dtype == DCFP_F32) rc = bwd ? launch_generic<float, true>(d, stream) : launch_generic<float, false>(d, stream);
      else rc = bwd ? launch_generic<__nv_bfloat16, true>(d, stream) : launch_generic<__nv_bfloat16, false>(d, stream);
      if (rc) return rc;
    }
  }
  if (n_tiled == 0) return 0;
  // chunk length: ~512 KB per CTA, shortened while the call cannot fill ~4 waves of 4 CTAs/SM
  int chunk = kTargetBoxesPerChunk;
  while (chunk > 16 && total_boxes / chunk < 4LL * 4 * kNumSMs) chunk >>= 1;

  const int big = bwd ? kBigGroupBwd : kBigGroupFwd;
  for (int first = 0; first < n_tiled;) {
    const int m = std::min(n_tiled - first, big);
    int rc;
    if (m <= kSmallGroup) {
      if (dtype == DCFP_F32)
        rc = bwd ? run_tiled<float, true, kSmallGroup>(descs, tiled + first, m, chunk, stream)
                 : run_tiled<float, false, kSmallGroup>(descs, tiled + first, m, chunk, stream);
      else
        rc = bwd ? run_tiled<__nv_bfloat16, true, kSmallGroup>(descs, tiled + first, m, chunk, stream)
                 : run_tiled<__nv_bfloat16, false, kSmallGroup>(descs, tiled + first, m, chunk, stream);
    } else if (bwd) {
      rc = dtype == DCFP_F32 ? run_tiled<float, true, kBigGroupBwd>(descs, tiled + first, m, chunk, stream)
                             : run_tiled<__nv_bfloat16, true, kBigGroupBwd>(descs, tiled + first, m, chunk, stream);
    } else {
      rc = dtype == DCFP_F32 ? run_tiled<float, false, kBigGroupFwd>(descs, tiled + first, m, chunk, stream)
                             : run_tiled<__nv_bfloat16, false, kBigGroupFwd>(descs, tiled + first, m, chunk, stream);
    }
    if (rc) return rc;
    first += m;
  }
  return 0;
}

}  // namespace
}  // namespace dcfp

extern "C" int dcfp_class_stats(const dcfp_layer_desc* desc_host, void* stream) {
  return dcfp::run(desc_host, 1, static_cast<cudaStream_t>(stream));
}

extern "C" int dcfp_class_stats_grouped(const dcfp_layer_desc* descs_host, int n_layers, void* stream) {
  return dcfp::run(descs_host, n_layers, static_cast<cudaStream_t>(stream));
}

extern "C" int dcfp_label_keys(const void* label, int label_dtype, int N, int H0, int W0, int h, int w, int K, uint8_t* keys,
                               double* cnt, void* stream) {
  using namespace dcfp;
  DCFP_REQUIRE(label && keys, DCFP_EINVAL, "label_keys: null pointer");
  DCFP_REQUIRE(N > 0 && H0 > 0 && W0 > 0 && h > 0 && w > 0, DCFP_EINVAL, "label_keys: bad extent");
  DCFP_REQUIRE(K >= 1 && K <= DCFP_MAX_CLASSES, DCFP_ETOOBIG, "label_keys: K=%d outside [1,%d]", K, DCFP_MAX_CLASSES);
  DCFP_REQUIRE(label_dtype >= DCFP_LABEL_U8 && label_dtype <= DCFP_LABEL_I64, DCFP_EINVAL, "label_keys: unknown label dtype %d",
               label_dtype);
  const long long total = static_cast<long long>(N) * h * w;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 4LL * kNumSMs));
  label_keys_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      label, label_dtype, N, H0, W0, h, w, K, static_cast<float>(H0) / static_cast<float>(h),
      static_cast<float>(W0) / static_cast<float>(w), keys, cnt);
  return finish_launch("label_keys");
}
