// K1 -- label-keyed segmented reduction over conv/BN feature maps (sm_100a).
//
//   S1[k][c] += sum_{pixels p of class k} v(p, c),  S2[k][c] += sum v^2,  cnt[k] += #pixels
//
// HBM-bound: every feature-map byte is read exactly once; there is no dense contraction, so no
// tensor cores.  Design:
//
//  * NCHW planes are pixel-contiguous, but the class key varies along pixels and is identical
//    across channels.  Each warp therefore stages a [32 channels x 256 B] tile into shared
//    memory with 32 one-dimensional bulk async copies (TMA engine, cp.async.bulk + mbarrier;
//    one fully coalesced 256-B row per lane) and then reads it back TRANSPOSED: lane == channel,
//    so the label of every pixel is warp-uniform.  Rows are padded by 16 B, which makes the
//    lane-per-row 128-bit shared loads bank-conflict free.
//  * With a warp-uniform label the reduction is a run-length accumulate in registers: while the
//    (nearest-down-sampled, in-register) label does not change, a1 += v, a2 += v*v; on a label
//    change the run is flushed to a [K x 32] accumulator in shared memory.  A whole 4-pixel
//    group is handled by one compare when its packed labels equal the current run's.
//  * Every warp owns a private 2-stage pipeline (its own mbarriers), so the main loop has no
//    CTA-wide synchronisation.  A CTA covers 32 channels x one pixel chunk; its 4 warps split
//    the chunk's 256-B segments round-robin.  Shared accumulators are per-warp copies (plain
//    read-modify-write) when K is small, one CTA-wide copy updated with shared atomics otherwise.
//  * At the end the CTA adds its [K x 32] partials into the fp64 arena with coalesced RED.F64
//    (only classes it actually met), so the cross-CTA / cross-image combine is done in fp64.
//  * `dcfp_class_stats_grouped` runs any number of resident layers in ONE launch: the layer
//    table travels in kernel parameter space and each CTA binary-searches its layer.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"

namespace dcfp {
namespace {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kStages = 2;
constexpr int kSegBytes = 256;                    // bytes of one channel row per stage
constexpr int kPitch = kSegBytes + 16;            // padded row pitch in shared memory
constexpr int kBufBytes = 32 * kPitch;            // one [32 x 256 B] tile
constexpr int kGroups = kSegBytes / 16;           // 128-bit groups per row
constexpr int kTargetSegsPerChunk = 64;           // 64 x 8 KB = 512 KB of input per CTA
constexpr int kPrivateAccMaxK = 24;               // per-warp accumulator copies up to this K

struct LayerDev {
  const char* x;
  const char* dy;
  const float* scale;
  const float* shift;
  const void* label;
  double* S1;
  double* S2;
  double* cnt;
  int32_t N, C, HW, w;
  int32_t H0, W0;
  float sh, sw;  // (float)H0 / h, (float)W0 / w  -- legacy `nearest` scale
  int32_t n_cg, segs_per_plane, n_segs, segs_per_chunk;
  int32_t label_dtype, same_res;
};

// Layer table in kernel parameter space (no H2D copy, no workspace).  Two sizes: per-layer hook
// launches use the small one so the launch does not copy ~20 KB of parameters.
template <int MAXL>
struct GroupParams {
  int32_t n_layers;
  int32_t K;
  int32_t tile_prefix[MAXL + 1];
  LayerDev L[MAXL];
};
constexpr int kSmallGroup = 4;

__device__ __forceinline__ int load_label(const void* label, int dtype, long long idx) {
  if (dtype == DCFP_LABEL_U8) return static_cast<const unsigned char*>(label)[idx];
  if (dtype == DCFP_LABEL_I32) return static_cast<const int*>(label)[idx];
  const long long v = static_cast<const long long*>(label)[idx];
  return (v < 0 || v > 0x7fffffffLL) ? -1 : static_cast<int>(v);
}

// class key of pixel p of plane n at this layer's resolution (K == "dropped")
__device__ __forceinline__ unsigned class_of(const LayerDev& L, int K, int n, int p) {
  if (p >= L.HW) return K;
  if (L.label == nullptr) return 0;
  int i = p / L.w;
  int j = p - i * L.w;
  if (!L.same_res) {
    i = min(static_cast<int>(floorf(i * L.sh)), L.H0 - 1);
    j = min(static_cast<int>(floorf(j * L.sw)), L.W0 - 1);
  }
  const int l = load_label(L.label, L.label_dtype, (static_cast<long long>(n) * L.H0 + i) * L.W0 + j);
  return (l < 0 || l >= K) ? K : l;
}

template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static constexpr int kPerGroup = 4;  // values per 128-bit group
  __device__ static __forceinline__ void unpack(const uint4& r, float* v) {
    v[0] = __uint_as_float(r.x);
    v[1] = __uint_as_float(r.y);
    v[2] = __uint_as_float(r.z);
    v[3] = __uint_as_float(r.w);
  }
};
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int kPerGroup = 8;
  __device__ static __forceinline__ void unpack(const uint4& r, float* v) {
    v[0] = __uint_as_float(r.x << 16);
    v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16);
    v[3] = __uint_as_float(r.y & 0xffff0000u);
    v[4] = __uint_as_float(r.z << 16);
    v[5] = __uint_as_float(r.z & 0xffff0000u);
    v[6] = __uint_as_float(r.w << 16);
    v[7] = __uint_as_float(r.w & 0xffff0000u);
  }
};

// Flush one finished run of one lane (= one channel) into the shared accumulators.  Kept out of
// line: it sits on the rare path (label change) and would otherwise be replicated 64x in the
// unrolled main loop.  Everything is passed by value so the caller's state stays in registers.
template <bool SHARED_ACC>
__device__ __noinline__ void flush_run(float* acc1, float* acc2, unsigned* cnt_s, unsigned cur, int lane, float a1, float a2,
                                       int run_px) {
  if (SHARED_ACC) {
    atomicAdd(&acc1[cur * 32 + lane], a1);
    atomicAdd(&acc2[cur * 32 + lane], a2);
  } else {
    acc1[cur * 32 + lane] += a1;
    acc2[cur * 32 + lane] += a2;
  }
  if (lane == 0) atomicAdd(&cnt_s[cur], static_cast<unsigned>(run_px));
}

// Run-length accumulator of one lane; every field except a1/a2 is warp-uniform.
template <bool SHARED_ACC>
struct RunAcc {
  float a1 = 0.f, a2 = 0.f;
  unsigned cur, curw;
  int run_px = 0;
  float* acc1;
  float* acc2;
  unsigned* cnt_s;
  int K, lane;

  __device__ __forceinline__ void flush() {
    if (cur < static_cast<unsigned>(K) && run_px > 0) flush_run<SHARED_ACC>(acc1, acc2, cnt_s, cur, lane, a1, a2, run_px);
    a1 = 0.f;
    a2 = 0.f;
    run_px = 0;
  }
  // four consecutive pixels whose packed class keys are `wv` (one byte each, warp-uniform)
  __device__ __forceinline__ void add4(const float* v, unsigned wv) {
    if (wv == curw) {
      a1 += (v[0] + v[1]) + (v[2] + v[3]);
      a2 = fmaf(v[0], v[0], a2);
      a2 = fmaf(v[1], v[1], a2);
      a2 = fmaf(v[2], v[2], a2);
      a2 = fmaf(v[3], v[3], a2);
      run_px += 4;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const unsigned l = (wv >> (8 * q)) & 0xffu;
        if (l != cur) {
          flush();
          cur = l;
          curw = l * 0x01010101u;
        }
        a1 += v[q];
        a2 = fmaf(v[q], v[q], a2);
        run_px += 1;
      }
    }
  }
};

template <typename T, bool BWD, bool SHARED_ACC>
__device__ __forceinline__ void process_tile(const LayerDev& L, const int K, const int tile, unsigned char* smem) {
  constexpr int kSegPx = kSegBytes / static_cast<int>(sizeof(T));
  constexpr int kWords = kSegPx / 4;  // packed label words per segment (<= 32)
  constexpr int kTens = BWD ? 2 : 1;
  constexpr int kAccCopies = SHARED_ACC ? 1 : kWarps;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = tile / L.n_cg, cg = tile - chunk * L.n_cg;
  const int c = cg * 32 + lane;
  const bool active = c < L.C;
  const int n_active = min(32, L.C - cg * 32);

  // ---- shared-memory carve-up --------------------------------------------------------------
  float* acc1 = reinterpret_cast<float*>(smem);          // [copies][K][32]
  float* acc2 = acc1 + kAccCopies * K * 32;               // [copies][K][32]
  unsigned* cnt_s = reinterpret_cast<unsigned*>(acc2 + kAccCopies * K * 32);  // [K]
  uintptr_t off = reinterpret_cast<uintptr_t>(cnt_s + K) - reinterpret_cast<uintptr_t>(smem);
  off = (off + 7) & ~uintptr_t(7);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + off);  // [kWarps][kStages]
  off += sizeof(unsigned long long) * kWarps * kStages;
  off = (off + 127) & ~uintptr_t(127);
  unsigned char* bufs = smem + off;  // [kWarps][kStages][kTens][kBufBytes]

  for (int i = tid; i < 2 * kAccCopies * K * 32 + K; i += kThreads) reinterpret_cast<unsigned*>(smem)[i] = 0u;
  if (tid < kWarps * kStages) mbar_init(smem_u32(&bars[tid]), 1);
  mbar_fence_init();
  __syncthreads();

  const int seg_begin = chunk * L.segs_per_chunk;
  const int seg_end = min(seg_begin + L.segs_per_chunk, L.n_segs);
  const uint64_t policy = policy_evict_first();
  const uint32_t my_bufs = smem_u32(bufs + static_cast<size_t>(warp) * kStages * kTens * kBufBytes);
  const uint32_t my_bars = smem_u32(&bars[warp * kStages]);

  auto issue = [&](int it) {
    const int seg = seg_begin + warp + it * kWarps;
    if (seg >= seg_end) return;
    const int stage = it % kStages;
    const int n = seg / L.segs_per_plane;
    const int p0 = (seg - n * L.segs_per_plane) * kSegPx;
    const uint32_t bytes = static_cast<uint32_t>(min(kSegPx, L.HW - p0)) * sizeof(T);
    const uint32_t bar = my_bars + stage * 8;
    if (lane == 0) mbar_expect_tx(bar, bytes * n_active * kTens);
    __syncwarp();
    if (active) {
      const size_t goff = ((static_cast<size_t>(n) * L.C + c) * L.HW + p0) * sizeof(T);
      const uint32_t dst = my_bufs + stage * (kTens * kBufBytes) + lane * kPitch;
      bulk_g2s(dst, L.x + goff, bytes, bar, policy);
      if (BWD) bulk_g2s(dst + kBufBytes, L.dy + goff, bytes, bar, policy);
    }
  };
  // packed class keys of the segment's pixels 4*lane .. 4*lane+3 (lanes < kWords)
  auto label_word = [&](int it) -> unsigned {
    const int seg = seg_begin + warp + it * kWarps;
    if (seg >= seg_end || lane >= kWords) return 0u;
    const int n = seg / L.segs_per_plane;
    const int p = (seg - n * L.segs_per_plane) * kSegPx + 4 * lane;
    return class_of(L, K, n, p) | (class_of(L, K, n, p + 1) << 8) | (class_of(L, K, n, p + 2) << 16) |
           (class_of(L, K, n, p + 3) << 24);
  };

  float sc = 1.f, sf = 0.f;
  if (active) {
    if (L.scale) sc = L.scale[c];
    if (L.shift) sf = L.shift[c];
  }

  RunAcc<SHARED_ACC> ra;
  ra.K = K;
  ra.lane = lane;
  ra.cur = K;
  ra.curw = static_cast<unsigned>(K) * 0x01010101u;
  ra.acc1 = acc1 + (SHARED_ACC ? 0 : warp * K * 32);
  ra.acc2 = acc2 + (SHARED_ACC ? 0 : warp * K * 32);
  ra.cnt_s = cnt_s;

#pragma unroll
  for (int s = 0; s < kStages; ++s) issue(s);
  unsigned lw = label_word(0);

  const int n_my = (seg_end - seg_begin - warp + kWarps - 1) / kWarps;  // segments of this warp
  for (int it = 0; it < n_my; ++it) {
    const unsigned lw_next = label_word(it + 1);  // global loads overlap the wait below
    const int stage = it % kStages;
    mbar_wait(my_bars + stage * 8, (it / kStages) & 1);
    const uint32_t row = my_bufs + stage * (kTens * kBufBytes) + lane * kPitch;
    // all shared loads of the stage first (ILP), then the run-length accumulate from registers
    uint4 rx[kGroups];
#pragma unroll
    for (int g = 0; g < kGroups; ++g) rx[g] = lds128(row + g * 16);
    if (BWD) {
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        float v[Elem<T>::kPerGroup], d[Elem<T>::kPerGroup];
        Elem<T>::unpack(rx[g], v);
        Elem<T>::unpack(lds128(row + kBufBytes + g * 16), d);
#pragma unroll
        for (int q = 0; q < Elem<T>::kPerGroup; ++q) v[q] = d[q] * fmaf(v[q], sc, sf);
#pragma unroll
        for (int h = 0; h < Elem<T>::kPerGroup / 4; ++h)
          ra.add4(v + 4 * h, __shfl_sync(0xffffffffu, lw, g * (Elem<T>::kPerGroup / 4) + h));
      }
    } else {
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        float v[Elem<T>::kPerGroup];
        Elem<T>::unpack(rx[g], v);
#pragma unroll
        for (int q = 0; q < Elem<T>::kPerGroup; ++q) v[q] = fmaf(v[q], sc, sf);
#pragma unroll
        for (int h = 0; h < Elem<T>::kPerGroup / 4; ++h)
          ra.add4(v + 4 * h, __shfl_sync(0xffffffffu, lw, g * (Elem<T>::kPerGroup / 4) + h));
      }
    }
    __syncwarp();
    issue(it + kStages);  // refill the stage just consumed
    lw = lw_next;
  }
  ra.flush();
  __syncthreads();

  // ---- CTA partials -> fp64 arena (coalesced RED.F64; only classes this CTA met) --------------
  for (int idx = tid; idx < K * 32; idx += kThreads) {
    const int k = idx >> 5, cl = idx & 31;
    if (cnt_s[k] == 0u || cl >= n_active) continue;
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
  if (cg == 0 && L.cnt != nullptr)
    for (int k = tid; k < K; k += kThreads)
      if (cnt_s[k]) atomicAdd(&L.cnt[k], static_cast<double>(cnt_s[k]));
}

template <typename T, bool BWD, bool SHARED_ACC, int MAXL>
__global__ void __launch_bounds__(kThreads, 2) class_stats_kernel(const __grid_constant__ GroupParams<MAXL> P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tile = blockIdx.x;
  int lo = 0, hi = P.n_layers;  // largest l with tile_prefix[l] <= tile
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (P.tile_prefix[mid] <= tile) lo = mid;
    else hi = mid;
  }
  process_tile<T, BWD, SHARED_ACC>(P.L[lo], P.K, tile - P.tile_prefix[lo], smem);
}

// Generic path: any extent / alignment / layout (tiny 1x1..6x6 maps, odd crops, NHWC).  One
// thread per channel walks the pixels of one plane chunk; runs are flushed straight to the arena.
template <typename T, bool BWD>
__global__ void class_stats_generic_kernel(const LayerDev L, const int K, const int nhwc, const int px_per_block) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.z;
  const int p_begin = blockIdx.y * px_per_block, p_end = min(p_begin + px_per_block, L.HW);
  if (c >= L.C) return;
  const float sc = L.scale ? L.scale[c] : 1.f, sf = L.shift ? L.shift[c] : 0.f;
  const T* x = reinterpret_cast<const T*>(L.x);
  const T* dy = reinterpret_cast<const T*>(L.dy);
  float a1 = 0.f, a2 = 0.f;
  int run = 0;
  unsigned cur = K;
  auto flush = [&]() {
    if (cur < static_cast<unsigned>(K) && run > 0) {
      atomicAdd(&L.S1[static_cast<size_t>(cur) * L.C + c], static_cast<double>(a1));
      atomicAdd(&L.S2[static_cast<size_t>(cur) * L.C + c], static_cast<double>(a2));
      if (c == 0 && L.cnt) atomicAdd(&L.cnt[cur], static_cast<double>(run));
    }
    a1 = a2 = 0.f;
    run = 0;
  };
  for (int p = p_begin; p < p_end; ++p) {
    const unsigned k = class_of(L, K, n, p);
    if (k != cur) {
      flush();
      cur = k;
    }
    const size_t idx = nhwc ? (static_cast<size_t>(n) * L.HW + p) * L.C + c : (static_cast<size_t>(n) * L.C + c) * L.HW + p;
    float v = fmaf(static_cast<float>(x[idx]), sc, sf);
    if (BWD) v *= static_cast<float>(dy[idx]);
    a1 += v;
    a2 = fmaf(v, v, a2);
    ++run;
  }
  flush();
}

size_t tile_smem_bytes(int K, bool bwd, bool shared_acc) {
  const int copies = shared_acc ? 1 : kWarps;
  size_t off = static_cast<size_t>(2) * copies * K * 32 * 4 + static_cast<size_t>(K) * 4;
  off = (off + 7) & ~size_t(7);
  off += 8 * kWarps * kStages;
  off = (off + 127) & ~size_t(127);
  return off + static_cast<size_t>(kWarps) * kStages * (bwd ? 2 : 1) * kBufBytes;
}

int validate(const dcfp_layer_desc& d, int idx) {
  DCFP_REQUIRE(d.x && d.S1 && d.S2, DCFP_EINVAL, "class_stats[%d]: x/S1/S2 must be non-null", idx);
  DCFP_REQUIRE(d.N > 0 && d.C > 0 && d.h > 0 && d.w > 0, DCFP_EINVAL, "class_stats[%d]: bad extent N=%d C=%d h=%d w=%d", idx,
               d.N, d.C, d.h, d.w);
  DCFP_REQUIRE(d.K >= 1 && d.K <= DCFP_MAX_CLASSES, DCFP_ETOOBIG, "class_stats[%d]: K=%d outside [1,%d]", idx, d.K,
               DCFP_MAX_CLASSES);
  DCFP_REQUIRE(d.dtype == DCFP_F32 || d.dtype == DCFP_BF16, DCFP_EINVAL, "class_stats[%d]: unknown dtype %d", idx, d.dtype);
  DCFP_REQUIRE(d.layout == DCFP_NCHW || d.layout == DCFP_NHWC, DCFP_EINVAL, "class_stats[%d]: unknown layout %d", idx, d.layout);
  if (d.label) {
    DCFP_REQUIRE(d.H0 > 0 && d.W0 > 0, DCFP_EINVAL, "class_stats[%d]: bad label extent %dx%d", idx, d.H0, d.W0);
    DCFP_REQUIRE(d.label_dtype >= DCFP_LABEL_U8 && d.label_dtype <= DCFP_LABEL_I64, DCFP_EINVAL,
                 "class_stats[%d]: unknown label dtype %d", idx, d.label_dtype);
  } else {
    DCFP_REQUIRE(d.K == 1, DCFP_EINVAL, "class_stats[%d]: label == NULL requires K == 1", idx);
  }
  DCFP_REQUIRE(static_cast<long long>(d.h) * d.w < (1LL << 30), DCFP_ETOOBIG, "class_stats[%d]: plane too large", idx);
  return 0;
}

LayerDev to_dev(const dcfp_layer_desc& d) {
  LayerDev L{};
  L.x = static_cast<const char*>(d.x);
  L.dy = static_cast<const char*>(d.dy);
  L.scale = d.scale;
  L.shift = d.shift;
  L.label = d.label;
  L.S1 = d.S1;
  L.S2 = d.S2;
  L.cnt = d.cnt;
  L.N = d.N;
  L.C = d.C;
  L.HW = d.h * d.w;
  L.w = d.w;
  L.H0 = d.label ? d.H0 : d.h;
  L.W0 = d.label ? d.W0 : d.w;
  L.sh = static_cast<float>(L.H0) / static_cast<float>(d.h);
  L.sw = static_cast<float>(L.W0) / static_cast<float>(d.w);
  L.label_dtype = d.label_dtype;
  L.same_res = (L.H0 == d.h && L.W0 == d.w);
  return L;
}

// the bulk-copy path needs 16-B aligned planes; everything else takes the generic kernel
bool tiled_ok(const dcfp_layer_desc& d) {
  const size_t es = d.dtype == DCFP_F32 ? 4 : 2;
  const size_t plane = static_cast<size_t>(d.h) * d.w * es;
  if (d.layout != DCFP_NCHW) return false;
  if (plane % 16 != 0 || plane < 512) return false;
  if (reinterpret_cast<uintptr_t>(d.x) % 16 != 0) return false;
  if (d.dy && reinterpret_cast<uintptr_t>(d.dy) % 16 != 0) return false;
  return true;
}

template <typename T, bool BWD>
int launch_generic(const dcfp_layer_desc& d, cudaStream_t stream) {
  const LayerDev L = to_dev(d);
  const int threads = 128;
  const int px_per_block = 256;
  dim3 grid((d.C + threads - 1) / threads, (L.HW + px_per_block - 1) / px_per_block, d.N);
  class_stats_generic_kernel<T, BWD><<<grid, threads, 0, stream>>>(L, d.K, d.layout == DCFP_NHWC, px_per_block);
  return finish_launch("class_stats_generic");
}

template <typename T, bool BWD, bool SHARED_ACC, int MAXL>
int launch_tiled(const GroupParams<MAXL>& P, int n_tiles, cudaStream_t stream) {
  const size_t smem = tile_smem_bytes(P.K, BWD, SHARED_ACC);
  auto kern = class_stats_kernel<T, BWD, SHARED_ACC, MAXL>;
  int rc = ensure_smem(reinterpret_cast<const void*>(kern), static_cast<int>(smem));
  if (rc) return rc;
  kern<<<n_tiles, kThreads, smem, stream>>>(P);
  return finish_launch("class_stats");
}

int run(const dcfp_layer_desc* descs, int n_layers, cudaStream_t stream) {
  DCFP_REQUIRE(descs != nullptr && n_layers > 0, DCFP_EINVAL, "class_stats: no layers");
  DCFP_REQUIRE(n_layers <= DCFP_MAX_GROUP_LAYERS, DCFP_ETOOBIG, "class_stats: %d layers > %d per call", n_layers,
               DCFP_MAX_GROUP_LAYERS);
  const int K = descs[0].K, dtype = descs[0].dtype;
  const bool bwd = descs[0].dy != nullptr;
  for (int i = 0; i < n_layers; ++i) {
    int rc = validate(descs[i], i);
    if (rc) return rc;
    DCFP_REQUIRE(descs[i].K == K && descs[i].dtype == dtype && (descs[i].dy != nullptr) == bwd, DCFP_EINVAL,
                 "class_stats[%d]: K / dtype / functor differ inside one group", i);
  }
  GroupParams<DCFP_MAX_GROUP_LAYERS> P;  // ~20 KB on the host stack
  P.n_layers = 0;
  P.K = K;
  P.tile_prefix[0] = 0;
  const int seg_px = kSegBytes / (dtype == DCFP_F32 ? 4 : 2);
  long long total_segs = 0;
  for (int i = 0; i < n_layers; ++i)
    if (tiled_ok(descs[i])) {
      const long long spp = (static_cast<long long>(descs[i].h) * descs[i].w + seg_px - 1) / seg_px;
      total_segs += spp * descs[i].N * ((descs[i].C + 31) / 32);
    }
  // chunk length: ~512 KB per CTA, shortened when the whole call cannot otherwise fill 4 waves
  int chunk_target = kTargetSegsPerChunk;
  while (chunk_target > 8 && total_segs / chunk_target < 4LL * 2 * kNumSMs) chunk_target >>= 1;

  for (int i = 0; i < n_layers; ++i) {
    const dcfp_layer_desc& d = descs[i];
    if (!tiled_ok(d)) {
      int rc;
      if (dtype == DCFP_F32) rc = bwd ? launch_generic<float, true>(d, stream) : launch_generic<float, false>(d, stream);
      else rc = bwd ? launch_generic<__nv_bfloat16, true>(d, stream) : launch_generic<__nv_bfloat16, false>(d, stream);
      if (rc) return rc;
      continue;
    }
    LayerDev L = to_dev(d);
    L.n_cg = (d.C + 31) / 32;
    L.segs_per_plane = (L.HW + seg_px - 1) / seg_px;
    L.n_segs = L.segs_per_plane * d.N;
    L.segs_per_chunk = chunk_target;
    const int n_chunks = (L.n_segs + L.segs_per_chunk - 1) / L.segs_per_chunk;
    const long long tiles = static_cast<long long>(n_chunks) * L.n_cg;
    DCFP_REQUIRE(P.tile_prefix[P.n_layers] + tiles < (1LL << 31), DCFP_ETOOBIG, "class_stats: too many tiles");
    P.L[P.n_layers] = L;
    P.tile_prefix[P.n_layers + 1] = P.tile_prefix[P.n_layers] + static_cast<int>(tiles);
    ++P.n_layers;
  }
  if (P.n_layers == 0) return 0;
  const int n_tiles = P.tile_prefix[P.n_layers];
  const bool shared_acc = K > kPrivateAccMaxK;
#define DCFP_K1_DISPATCH(T, PP)                                                                          \
  (bwd ? (shared_acc ? launch_tiled<T, true, true>(PP, n_tiles, stream) : launch_tiled<T, true, false>(PP, n_tiles, stream)) \
       : (shared_acc ? launch_tiled<T, false, true>(PP, n_tiles, stream) : launch_tiled<T, false, false>(PP, n_tiles, stream)))
  if (P.n_layers <= kSmallGroup) {
    GroupParams<kSmallGroup> Q;
    Q.n_layers = P.n_layers;
    Q.K = P.K;
    for (int i = 0; i < P.n_layers; ++i) Q.L[i] = P.L[i];
    for (int i = 0; i <= P.n_layers; ++i) Q.tile_prefix[i] = P.tile_prefix[i];
    return dtype == DCFP_F32 ? DCFP_K1_DISPATCH(float, Q) : DCFP_K1_DISPATCH(__nv_bfloat16, Q);
  }
  return dtype == DCFP_F32 ? DCFP_K1_DISPATCH(float, P) : DCFP_K1_DISPATCH(__nv_bfloat16, P);
#undef DCFP_K1_DISPATCH
}

}  // namespace
}  // namespace dcfp

extern "C" int dcfp_class_stats(const dcfp_layer_desc* desc_host, void* stream) {
  return dcfp::run(desc_host, 1, static_cast<cudaStream_t>(stream));
}

extern "C" int dcfp_class_stats_grouped(const dcfp_layer_desc* descs_host, int n_layers, void* stream) {
  return dcfp::run(descs_host, n_layers, static_cast<cudaStream_t>(stream));
}
