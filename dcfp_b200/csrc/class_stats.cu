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
#include <cstdlib>

#include "common.cuh"

namespace dcfp {
namespace {

constexpr int kWarpsPrivate = 4;  // warps per CTA when every warp owns an accumulator table (small K)
constexpr int kWarpsShared = 8;   // warps per CTA sharing one table through shared atomics (large K)
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
  int32_t C, HW, n_cg, ld;
  int32_t boxes_per_plane, n_boxes, boxes_per_chunk;
  int32_t centered;  // DCFP_AFFINE_INVSTD_MEAN: shift holds the batch mean
};

// Layer table + TMA descriptors in kernel parameter space (no H2D copy, no workspace).
template <int MAXL, int TENS>
struct GroupParams {
  alignas(64) CUtensorMap maps[MAXL * TENS];  // [layer][x, dy]
  LayerDev L[MAXL];
  int32_t tile_prefix[MAXL + 1];
  int32_t n_layers;
  int32_t K;
  int32_t stages;
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

// acc[key][lane] += (a1, a2): the shared accumulator is an interleaved float2 [K][32] table, so one
// 64-bit load / FADD2 / 64-bit store updates both moments of (class, channel).  Lanes touch
// consecutive 8-byte slots: conflict-free.
template <bool SHARED_ACC>
__device__ __forceinline__ void acc_add(uint32_t acc_lane, unsigned key, float a1, float a2) {
  const uint32_t addr = acc_lane + key * 256u;
  if (SHARED_ACC) {  // one CTA-wide table (large K): shared-memory atomics
    float* p = reinterpret_cast<float*>(__cvta_shared_to_generic(addr));
    atomicAdd(p, a1);
    atomicAdd(p + 1, a2);
  } else {  // per-warp table: plain read-modify-write
    f2 cur;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(cur) : "r"(addr));
    cur = add2(cur, pack2(a1, a2));
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(cur) : "memory");
  }
}

// value of one pixel re-read from the staged box (per-pixel path of a quad that straddles a class
// boundary; a rolled loop keeps this code to one site)
template <typename T, bool BWD, bool AFFINE>
__device__ __forceinline__ float load_px(uint32_t addr, float sc, float sf) {
  float x, d = 1.f;
  if (sizeof(T) == 4) {
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr));
    if (BWD) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(d) : "r"(addr + kBoxBytes));
  } else {
    unsigned short hx, hd = 0;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hx) : "r"(addr));
    if (BWD) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hd) : "r"(addr + kBoxBytes));
    x = __uint_as_float(static_cast<unsigned>(hx) << 16);
    d = __uint_as_float(static_cast<unsigned>(hd) << 16);
  }
  if (BWD) return d * fmaf(x, sc, sf);
  return AFFINE ? fmaf(x, sc, sf) : x;
}

template <typename T, bool BWD, bool AFFINE>
__device__ __forceinline__ void load_group(uint32_t addr, f2 sc2, f2 sf2, f2* v) {
  Elem<T>::unpack(lds128(addr), v);
  if (BWD) {
    f2 d[Elem<T>::kPairs];
    Elem<T>::unpack(lds128(addr + kBoxBytes), d);
#pragma unroll
    for (int q = 0; q < Elem<T>::kPairs; ++q) v[q] = mul2(d[q], fma2(v[q], sc2, sf2));
  } else if (AFFINE) {
#pragma unroll
    for (int q = 0; q < Elem<T>::kPairs; ++q) v[q] = fma2(v[q], sc2, sf2);
  }
}

// position of a warp's it-th box inside the layer, advanced without divisions
struct BoxCursor {
  int n, b;  // plane, box inside the plane
  __device__ __forceinline__ void advance(int step, int boxes_per_plane) {
    b += step;
    while (b >= boxes_per_plane) {
      b -= boxes_per_plane;
      ++n;
    }
  }
};

// RUNLEN: keep the current class run (key, sum, sum of squares) in registers and touch the shared table only
// when the class changes -- the table of the large-K variant is updated with shared atomics (2 x ~64 cycles
// per warp-wide update), so the number of updates, not of pixels, is what it can afford.
template <typename T, bool BWD, bool AFFINE, bool SHARED_ACC, int WARPS, bool RUNLEN>
__device__ __forceinline__ void process_tile(const LayerDev& L, const CUtensorMap* maps, const int K, const int stages,
                                             const int tile, unsigned char* smem) {
  constexpr int kBoxPx = kBoxRowBytes / static_cast<int>(sizeof(T));  // 32 (fp32) / 64 (bf16)
  constexpr int kWords = kBoxPx / 4;                                  // packed key words (quads) per box
  constexpr int kQuadsPerGroup = kWords / kGroups;                    // 1 (fp32) / 2 (bf16)
  constexpr int kTens = BWD ? 2 : 1;
  constexpr int kAccCopies = SHARED_ACC ? 1 : WARPS;
  constexpr int kPairs = Elem<T>::kPairs;
  constexpr int kStageBytes = kTens * kBoxBytes;
  constexpr int kThreadsT = WARPS * 32;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = tile / L.n_cg, cg = tile - chunk * L.n_cg;
  const int n_active = min(32, L.C - cg * 32);

  // ---- shared-memory carve-up: [boxes | accumulators (float2 [copies][K][32]) | mbarriers] -------
  unsigned char* bufs = smem;  // [WARPS][stages][kTens][kBoxBytes], 1024-B aligned
  float2* acc = reinterpret_cast<float2*>(smem + static_cast<size_t>(WARPS) * stages * kStageBytes);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(acc + kAccCopies * K * 32);

  for (int i = tid; i < kAccCopies * K * 32; i += kThreadsT) acc[i] = make_float2(0.f, 0.f);
  if (tid < WARPS * stages) mbar_init(smem_u32(&bars[tid]), 1);
  mbar_fence_init();
  __syncthreads();

  const int box_begin = chunk * L.boxes_per_chunk;
  const int box_end = min(box_begin + L.boxes_per_chunk, L.n_boxes);
  const int n_my = (box_end - box_begin - warp + WARPS - 1) / WARPS;  // boxes of this warp
  const uint64_t policy = policy_evict_first();
  const uint32_t my_bufs = smem_u32(bufs + static_cast<size_t>(warp) * stages * kStageBytes);
  const uint32_t my_bars = smem_u32(&bars[warp * stages]);
  const uint32_t acc_lane = smem_u32(acc + (SHARED_ACC ? 0 : warp * K * 32) + lane);
  const int row0 = cg * 32;

  BoxCursor issue_at, key_at;
  issue_at.n = (box_begin + warp) / L.boxes_per_plane;
  issue_at.b = (box_begin + warp) - issue_at.n * L.boxes_per_plane;
  key_at = issue_at;

  int issue_it = 0, issue_stage = 0;
  auto issue = [&]() {  // one elected lane arms the barrier and launches the tile copies
    if (issue_it < n_my) {
      if (lane == 0) {
        const uint32_t bar = my_bars + issue_stage * 8;
        const uint32_t dst = my_bufs + issue_stage * kStageBytes;
        mbar_expect_tx(bar, kStageBytes);
        tma_load_2d(dst, maps, issue_at.b * kBoxPx, issue_at.n * L.C + row0, bar, policy);
        if (BWD) tma_load_2d(dst + kBoxBytes, maps + 1, issue_at.b * kBoxPx, issue_at.n * L.C + row0, bar, policy);
      }
      issue_at.advance(WARPS, L.boxes_per_plane);
    }
    ++issue_it;
    if (++issue_stage == stages) issue_stage = 0;
  };
  // packed class keys of the box's pixels 4*lane .. 4*lane+3 (lanes < kWords); K = "dropped"
  const unsigned dropped = static_cast<unsigned>(K) * 0x01010101u;
  int key_it = 0;
  auto key_word = [&]() -> unsigned {
    unsigned w = dropped;
    if (key_it < n_my) {
      const int p = key_at.b * kBoxPx + 4 * lane;  // HW % 4 == 0: a word is entirely inside or outside the plane
      if (lane < kWords && p < L.HW)
        w = L.keys ? __ldg(reinterpret_cast<const unsigned*>(L.keys + static_cast<size_t>(key_at.n) * L.HW + p)) : 0u;
      key_at.advance(WARPS, L.boxes_per_plane);
    }
    ++key_it;
    return w;
  };

  float sc = 1.f, sf = 0.f;
  if (lane < n_active) {
    if (L.scale) sc = L.scale[row0 + lane];
    if (L.shift) sf = L.shift[row0 + lane];
    if (L.centered) sf = -sf * sc;  // (x - mean) * invstd == x * invstd + (-mean * invstd)
  }
  const f2 sc2 = pack2(sc, sc), sf2 = pack2(sf, sf);

  for (int s = 0; s < stages; ++s) issue();
  unsigned lw = key_word();

  // SWIZZLE_128B: the 16-B chunk index of row r is XORed with r % 8 (lane == row)
  const uint32_t row_off = static_cast<uint32_t>(lane * kBoxRowBytes);
  const uint32_t l7 = static_cast<uint32_t>(lane & 7);

  // run state (RUNLEN only): class of the open run (K = none / dropped) and its partial sums
  unsigned run_key = static_cast<unsigned>(K);
  float run1 = 0.f, run2 = 0.f;
  auto run_add = [&](unsigned key, float a1, float a2) {  // key is warp-uniform: no divergence
    if (key != run_key) {
      if (run_key < static_cast<unsigned>(K)) acc_add<SHARED_ACC>(acc_lane, run_key, run1, run2);
      run_key = key;
      run1 = a1;
      run2 = a2;
    } else {
      run1 += a1;
      run2 += a2;
    }
  };

  int stage = 0;
  uint32_t parity = 0;
  for (int it = 0; it < n_my; ++it) {
    const unsigned lw_next = key_word();  // global load overlaps the wait below
    mbar_wait(my_bars + stage * 8, parity);
    const uint32_t box = my_bufs + stage * kStageBytes + row_off;
    const unsigned w0 = __shfl_sync(0xffffffffu, lw, 0);
    const unsigned key0 = w0 & 0xffu;
    const bool uniform = __all_sync(0xffffffffu, lane >= kWords || lw == key0 * 0x01010101u);
    if (uniform) {
      // every pixel of the box has the same class: branch-free packed accumulate, one table update
      if (key0 < static_cast<unsigned>(K)) {
        f2 s1a = 0, s1b = 0, s2a = 0, s2b = 0;
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
          f2 v[kPairs];
          load_group<T, BWD, AFFINE>(box + ((g ^ l7) << 4), sc2, sf2, v);
#pragma unroll
          for (int h = 0; h < kPairs / 2; ++h) {
            s1a = add2(s1a, v[2 * h]);
            s1b = add2(s1b, v[2 * h + 1]);
            s2a = fma2(v[2 * h], v[2 * h], s2a);
            s2b = fma2(v[2 * h + 1], v[2 * h + 1], s2b);
          }
        }
        s1a = add2(s1a, s1b);
        s2a = add2(s2a, s2b);
        if (RUNLEN) run_add(key0, lo2(s1a) + hi2(s1a), lo2(s2a) + hi2(s2a));
        else acc_add<SHARED_ACC>(acc_lane, key0, lo2(s1a) + hi2(s1a), lo2(s2a) + hi2(s2a));
      } else if (RUNLEN) {
        run_add(static_cast<unsigned>(K), 0.f, 0.f);  // dropped pixels close the open run
      }
    } else {
      // a class boundary crosses the box: per quad (4 px, one packed key word) -- a quad with one
      // class is summed in registers and added to the table; a straddling quad goes pixel by pixel
#pragma unroll 1
      for (int g = 0; g < kGroups; ++g) {
        const uint32_t gaddr = box + ((static_cast<uint32_t>(g) ^ l7) << 4);
        f2 v[kPairs];
        load_group<T, BWD, AFFINE>(gaddr, sc2, sf2, v);
#pragma unroll
        for (int h = 0; h < kQuadsPerGroup; ++h) {
          const unsigned wv = __shfl_sync(0xffffffffu, lw, g * kQuadsPerGroup + h);
          const unsigned key = wv & 0xffu;
          if (wv == key * 0x01010101u) {
            if (key < static_cast<unsigned>(K)) {
              const f2 t1 = add2(v[2 * h], v[2 * h + 1]);
              const f2 t2 = fma2(v[2 * h + 1], v[2 * h + 1], mul2(v[2 * h], v[2 * h]));
              if (RUNLEN) run_add(key, lo2(t1) + hi2(t1), lo2(t2) + hi2(t2));
              else acc_add<SHARED_ACC>(acc_lane, key, lo2(t1) + hi2(t1), lo2(t2) + hi2(t2));
            } else if (RUNLEN) {
              run_add(static_cast<unsigned>(K), 0.f, 0.f);
            }
          } else {
#pragma unroll 1
            for (int e = 0; e < 4; ++e) {
              const unsigned ke = (wv >> (8 * e)) & 0xffu;
              if (ke < static_cast<unsigned>(K)) {
                const float x = load_px<T, BWD, AFFINE>(gaddr + (h * 4 + e) * static_cast<int>(sizeof(T)), sc, sf);
                if (RUNLEN) run_add(ke, x, x * x);
                else acc_add<SHARED_ACC>(acc_lane, ke, x, x * x);
              } else if (RUNLEN) {
                run_add(static_cast<unsigned>(K), 0.f, 0.f);
              }
            }
          }
        }
      }
    }
    __syncwarp();
    issue();  // refill the stage just consumed
    lw = lw_next;
    if (++stage == stages) {
      stage = 0;
      parity ^= 1u;
    }
  }
  if (RUNLEN) run_add(static_cast<unsigned>(K), 0.f, 0.f);  // close the last run
  __syncthreads();

  // ---- CTA partials -> fp64 arena (coalesced RED.F64; zero partials are skipped) ----------------
  for (int idx = tid; idx < K * 32; idx += kThreadsT) {
    const int k = idx >> 5, cl = idx & 31;
    if (cl >= n_active) continue;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int w = 0; w < kAccCopies; ++w) {
      const float2 a = acc[w * K * 32 + idx];
      s1 += a.x;
      s2 += a.y;
    }
    if (s1 == 0.f && s2 == 0.f) continue;  // class not met by this CTA (or all-zero values): nothing to add
    const size_t o = static_cast<size_t>(k) * L.ld + row0 + cl;
    atomicAdd(&L.S1[o], static_cast<double>(s1));
    atomicAdd(&L.S2[o], static_cast<double>(s2));
  }
}

template <typename T, bool BWD, bool AFFINE, bool SHARED_ACC, int WARPS, int MAXL, bool RUNLEN>
__global__ void __launch_bounds__(WARPS * 32, SHARED_ACC ? 2 : 4)
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
  process_tile<T, BWD, AFFINE, SHARED_ACC, WARPS, RUNLEN>(P.L[lo], &P.maps[lo * (BWD ? 2 : 1)], P.K, P.stages, tile - P.tile_prefix[lo],
                                                  smem);
}

// ---- channels_last (NHWC) path ---------------------------------------------------------------------
// x[n][p][c] with C contiguous: every pixel has ONE class for all its channels, so with lane = 4 consecutive
// channels the class key is warp-uniform by construction and a pixel's [128-channel slab] is read straight out of
// a TMA-staged box with one conflict-free LDS per lane -- no transposition.
//  * Each warp owns a slab and a pixel phase and runs a private 2-stage pipeline of [G pixels x 128 channels]
//    tensor-tile copies (cp.async.bulk.tensor.2d, mbarrier completion, evict-first): the bytes in flight live in
//    shared memory, not in registers or L1 miss queues (a register-staged LDG version of this kernel stalled at
//    56 % of the roofline with the same nominal bytes in flight).
//  * Accumulation is STATE-FREE straight-line code: the 8 pixel rows of one 64-bit key word (or the 4 of a quad, or
//    a single straddling pixel) are summed in registers and added to the class's row.  Carrying an open run in
//    registers across pixels made ptxas shuffle the accumulators at every possible run boundary (29 instructions
//    per pixel row, 23 % of them moves) and left 8 warps/SM issue-latency bound at half the roofline.
//  * Rows live in the warp's private SLOT CACHE: 12 rows of [2][128] fp32; the class held by row i is a register
//    of lane i, looked up with one ballot (fully associative).  Labels are spatially coherent, so a warp meets few
//    classes at a time; a 13th class evicts a row to the fp64 arena.  12 KB per warp for ANY K <= 255: no table
//    sized by K, no shared atomics, no CTA-wide barrier anywhere.
//  * Class keys are fetched 32 iterations at a time (lane j holds iteration block + j) and broadcast by shuffle:
//    a per-iteration global load put ~1 us on every iteration's critical path.
//  * The kernel is PERSISTENT (one CTA per SM walks a contiguous range of tiles ordered (layer, slab group,
//    chunk)); rows are folded into the arena only when the (layer, slab group) changes.
// cuDNN's tensor-core convolutions are NHWC-native: running the feature-map producer in channels_last removes its
// layout transposes (measured 52 -> 36 ms per c2 step) -- this is the layout bench.py scores.
constexpr int kNhwcWarps = 8;
constexpr int kNhwcSlab = 128;       // channels per warp row: 32 lanes x 4
constexpr int kNhwcSlots = 12;       // class rows per warp (tag of row i lives in lane i)
// one staged box = [G px][128 ch]: 8 KB forward (G = 16 fp32 / 32 bf16), 4 KB per tensor backward (x and dy boxes):
// 8 KB per pipeline stage either way, so the fixed per-iteration work (wait, key broadcast, refill) is paid per 8 KB
__host__ __device__ constexpr int nhwc_box_bytes(bool bwd) { return bwd ? 4096 : 8192; }

struct NhwcLayer {
  const uint8_t* keys;  // [N*HW]
  const float* scale;
  const float* shift;
  double* S1;
  double* S2;
  int32_t C, ld, centered;  // C: columns of the [rows][C] view the tensor map describes
  int32_t n_px;           // rows of that view: N * HW, or N * HW / 2 when fold2
  int32_t fold2;          // 64-channel layer viewed as [N*HW/2][128]: a row = 2 pixels, lanes 16-31 hold the odd one
  int32_t spc;            // slabs per CTA: 1, 2, 4 or 8
  int32_t n_slab_groups;  // ceil(ceil(C / 128) / spc)
  int32_t px_per_chunk;   // multiple of G * (8 / spc)
  int32_t n_chunks;
};
template <int MAXL, int TENS>
struct NhwcParams {
  alignas(64) CUtensorMap maps[MAXL * TENS];  // [layer][x, dy]: [n_px rows][C cols], box [G][128]
  NhwcLayer L[MAXL];
  int32_t tile_prefix[MAXL + 1];
  int32_t n_layers;
  int32_t K;
  int32_t stages;  // boxes in flight per warp
};
constexpr int kNhwcBigGroupFwd = 128;  // 128 * (128 + 72) B = 25.0 KB of kernel parameters
constexpr int kNhwcBigGroupBwd = 80;   //  80 * (256 + 72) B = 25.6 KB

// the lane's 4 channels of pixel `row` inside a staged box, as two packed fp32 pairs
template <typename T>
struct BoxRow;
template <>
struct BoxRow<float> {
  static constexpr int kRowBytes = kNhwcSlab * 4;
  __device__ static __forceinline__ void load(uint32_t box_lane, int row, f2& a, f2& b) {
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "r"(box_lane + row * (kNhwcSlab * 4)));
  }
  static constexpr int kLaneBytes = 16;
};
template <>
struct BoxRow<__nv_bfloat16> {
  static constexpr int kRowBytes = kNhwcSlab * 2;
  __device__ static __forceinline__ void load(uint32_t box_lane, int row, f2& a, f2& b) {
    unsigned lo, hi;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(box_lane + row * (kNhwcSlab * 2)));
    a = Elem<__nv_bfloat16>::widen(lo);
    b = Elem<__nv_bfloat16>::widen(hi);
  }
  static constexpr int kLaneBytes = 8;
};

constexpr int kNhwcStageBudget = 16 << 10;  // bytes of staging per warp: 4 x-boxes, or 2 (x, dy) pairs

template <typename T, bool BWD, bool AFFINE, int MAXL>
__global__ void __launch_bounds__(kNhwcWarps * 32, 1)
    class_stats_nhwc_kernel(const __grid_constant__ NhwcParams<MAXL, BWD ? 2 : 1> P) {
  constexpr int kNhwcBoxBytes = nhwc_box_bytes(BWD);
  constexpr int G = kNhwcBoxBytes / BoxRow<T>::kRowBytes;
  constexpr int Q = G / 4;  // packed key words (4 pixels each) per group
  constexpr int kTens = BWD ? 2 : 1;
  constexpr int kStageBytes = kTens * kNhwcBoxBytes;
  const int kStages = P.stages;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // [warp][stage][x | dy] boxes | [warp][slot][2][128] fp32 | [warp][stage] mbarriers
  float* slots = reinterpret_cast<float*>(smem + kNhwcWarps * kStages * kStageBytes);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(slots + kNhwcWarps * kNhwcSlots * 256);

  const unsigned K = static_cast<unsigned>(P.K);
  const int n_tiles = P.tile_prefix[P.n_layers];
  const int t_first = static_cast<int>(static_cast<long long>(blockIdx.x) * n_tiles / gridDim.x);
  const int t_last = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * n_tiles / gridDim.x);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* const mine = reinterpret_cast<float4*>(slots + static_cast<size_t>(warp) * kNhwcSlots * 256) + lane;
  const uint32_t my_bufs = smem_u32(smem + static_cast<size_t>(warp) * kStages * kStageBytes);
  const uint32_t my_bars = smem_u32(&bars[warp * kStages]);
  const uint32_t lane_off = static_cast<uint32_t>(lane * BoxRow<T>::kLaneBytes);
  const unsigned dropped = K * 0x01010101u;
  const bool direct = K <= static_cast<unsigned>(kNhwcSlots);  // slot == class, no tags
  const uint64_t policy = policy_evict_first();

  if (lane < kStages) mbar_init(my_bars + lane * 8, 1);
  for (int i = 0; i < kNhwcSlots * 2; ++i) mine[i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
  mbar_fence_init();
  __syncwarp();
  unsigned parity_bits = 0;  // bit s = parity of the next completion of stage s

  int layer = 0, cur_layer = -1, cur_sg = -1;
  int phases = kNhwcWarps, phase = 0, c0 = 0, col0 = 0;
  bool lane_on = false, fold2 = false;
  double* out1 = nullptr;  // &S1[c0], &S2[c0] of the current layer
  double* out2 = nullptr;
  size_t ld = 0;
  f2 sc01 = pack2(1.f, 1.f), sc23 = sc01, sf01 = 0, sf23 = 0;
  // slot cache: lane i (< kNhwcSlots) holds the class of row i (kFree = none); round-robin victim; last hit
  constexpr unsigned kFree = 0xffffffffu;
  unsigned my_tag = kFree, used = 0;  // `used`: CLOCK reference bits of the rows (warp-uniform)
  int victim = 0;
  unsigned last_key = 0xffffffffu;
  int last_slot = 0;
  const uint32_t mine_u32 = smem_u32(mine);

  auto row_to_arena = [&](int slot, unsigned cls) {  // fp32 row -> fp64 arena, row zeroed
    const float4 a = mine[slot * 64], b = mine[slot * 64 + 32];
    if (lane_on) {
      double* d1 = out1 + cls * ld;
      double* d2 = out2 + cls * ld;
      if (a.x != 0.f) atomicAdd(d1 + 0, static_cast<double>(a.x));
      if (a.y != 0.f) atomicAdd(d1 + 1, static_cast<double>(a.y));
      if (a.z != 0.f) atomicAdd(d1 + 2, static_cast<double>(a.z));
      if (a.w != 0.f) atomicAdd(d1 + 3, static_cast<double>(a.w));
      if (b.x != 0.f) atomicAdd(d2 + 0, static_cast<double>(b.x));
      if (b.y != 0.f) atomicAdd(d2 + 1, static_cast<double>(b.y));
      if (b.z != 0.f) atomicAdd(d2 + 2, static_cast<double>(b.z));
      if (b.w != 0.f) atomicAdd(d2 + 3, static_cast<double>(b.w));
    }
    mine[slot * 64] = make_float4(0.f, 0.f, 0.f, 0.f);
    mine[slot * 64 + 32] = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto slot_of = [&](unsigned cls) -> int {  // warp-uniform
    if (direct) return static_cast<int>(cls);
    if (cls == last_key) return last_slot;
    const unsigned hit = __ballot_sync(0xffffffffu, my_tag == cls);  // fully associative lookup in one vote
    int slot;
    if (hit) {
      slot = __ffs(hit) - 1;
      used |= 1u << slot;
    } else {
      // CLOCK replacement: take the first row (from the hand) not referenced since the hand last passed it, so the
      // classes of the rows being streamed stay resident while stale ones leave (round-robin evicted hot rows)
      constexpr unsigned kAll = (1u << kNhwcSlots) - 1u;
      if ((used & kAll) == kAll) used = 0;
      const unsigned cand = ~used & kAll;
      const unsigned ahead = cand & ~((1u << victim) - 1u);
      slot = __ffs(ahead ? ahead : cand) - 1;
      victim = slot + 1 == kNhwcSlots ? 0 : slot + 1;
      used |= 1u << slot;
      const unsigned old = __shfl_sync(0xffffffffu, my_tag, slot);
      if (old != kFree) row_to_arena(slot, old);
      if (lane == slot) my_tag = cls;
    }
    last_key = cls;
    last_slot = slot;
    return slot;
  };
  // row[class] += (sum, sum of squares) of this lane's 4 channels over a few pixels of one class.  No run state is
  // carried between pixel groups: every path below is straight-line code over values that die at the row update.
  auto row_add = [&](unsigned key, f2 s1a, f2 s1b, f2 s2a, f2 s2b) {
    if (key >= K) return;  // dropped pixels (warp-uniform)
    const uint32_t addr = mine_u32 + static_cast<uint32_t>(slot_of(key)) * 1024u;
    f2 a, b, c, d;
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(c), "=l"(d) : "r"(addr + 512u));
    a = add2(a, s1a);
    b = add2(b, s1b);
    c = add2(c, s2a);
    d = add2(d, s2b);
    asm volatile("st.shared.v2.b64 [%0], {%1,%2};" ::"r"(addr), "l"(a), "l"(b) : "memory");
    asm volatile("st.shared.v2.b64 [%0], {%1,%2};" ::"r"(addr + 512u), "l"(c), "l"(d) : "memory");
  };
  auto load_px = [&](uint32_t box_lane, int row, f2& a, f2& b) {  // value functor of one pixel row
    BoxRow<T>::load(box_lane, row, a, b);
    if (BWD) {
      f2 da, db;
      BoxRow<T>::load(box_lane + kNhwcBoxBytes, row, da, db);
      a = mul2(da, fma2(a, sc01, sf01));
      b = mul2(db, fma2(b, sc23, sf23));
    } else if (AFFINE) {
      a = fma2(a, sc01, sf01);
      b = fma2(b, sc23, sf23);
    }
  };
  auto fold = [&]() {  // every row of this warp -> arena (end of a (layer, slab group))
    if (direct) {
      for (unsigned k = 0; k < K; ++k) row_to_arena(static_cast<int>(k), k);
    } else {
      for (int slot = 0; slot < kNhwcSlots; ++slot) {
        const unsigned cls = __shfl_sync(0xffffffffu, my_tag, slot);
        if (cls != kFree) row_to_arena(slot, cls);
      }
      my_tag = kFree;
      victim = 0;
      used = 0;
      last_key = 0xffffffffu;
    }
  };

  for (int tile = t_first; tile < t_last; ++tile) {
    while (tile >= P.tile_prefix[layer + 1]) ++layer;
    const NhwcLayer& L = P.L[layer];
    const CUtensorMap* maps = &P.maps[layer * kTens];
    const int t = tile - P.tile_prefix[layer];
    const int sg = t / L.n_chunks;
    const int chunk = t - sg * L.n_chunks;
    if (layer != cur_layer || sg != cur_sg) {
      if (cur_layer >= 0) fold();
      cur_layer = layer;
      cur_sg = sg;
      const int spc = L.spc;
      phases = kNhwcWarps / spc;
      phase = warp / spc;
      col0 = (sg * spc + warp % spc) * kNhwcSlab;
      c0 = col0 + lane * 4;
      lane_on = c0 < L.C;  // C % 4 == 0: a lane's 4 channels are all inside or all outside (TMA zero-fills outside)
      fold2 = L.fold2 != 0;
      if (fold2) c0 &= 63;  // columns 64..127 of a pixel-pair row are channels 0..63 of the odd pixel
      ld = static_cast<size_t>(L.ld);
      out1 = L.S1 + c0;
      out2 = L.S2 + c0;
      sc01 = sc23 = pack2(1.f, 1.f);
      sf01 = sf23 = 0;
      if (lane_on && (BWD || AFFINE)) {
        float sc[4], sf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sc[j] = L.scale ? L.scale[c0 + j] : 1.f;
          sf[j] = L.shift ? L.shift[c0 + j] : 0.f;
          if (L.centered) sf[j] = -sf[j] * sc[j];
        }
        sc01 = pack2(sc[0], sc[1]);
        sc23 = pack2(sc[2], sc[3]);
        sf01 = pack2(sf[0], sf[1]);
        sf23 = pack2(sf[2], sf[3]);
      }
    }

    const int p_begin = chunk * L.px_per_chunk;
    const int p_end = min(p_begin + L.px_per_chunk, L.n_px);
    const int n_groups = (p_end - p_begin + G - 1) / G;
    const int n_my = col0 < L.C ? (n_groups - phase + phases - 1) / phases : 0;  // a warp whose slab is empty idles

    int issue_it = 0, issue_stage = 0;
    auto issue = [&]() {  // one elected lane arms the barrier and launches the tile copies of the warp's next group
      if (issue_it < n_my) {
        if (lane == 0) {
          const uint32_t bar = my_bars + issue_stage * 8;
          const uint32_t dst = my_bufs + issue_stage * kStageBytes;
          const int p = p_begin + (phase + issue_it * phases) * G;
          mbar_expect_tx(bar, kStageBytes);
          tma_load_2d(dst, maps, col0, p, bar, policy);  // rows past n_px / columns past C arrive as zeros
          if (BWD) tma_load_2d(dst + kNhwcBoxBytes, maps + 1, col0, p, bar, policy);
        }
      }
      ++issue_it;
      if (++issue_stage == kStages) issue_stage = 0;
    };
    // Class keys: lane j holds the packed keys of the warp's iteration (block + j), fetched 32 iterations at a time
    // and one block ahead, then broadcast with a shuffle.  (A per-iteration global load -- even issued one iteration
    // early -- put its full latency on every iteration's critical path: 8 warps/SM cannot hide ~1 us per 4 KB box.)
    const int kf = fold2 ? 2 : 1;  // key bytes per row
    auto fetch_keys = [&](int it0, unsigned* dst) {
      const int it = it0 + lane;
#pragma unroll
      for (int q = 0; q < 2 * Q; ++q) dst[q] = dropped;
      if (it < n_my) {
        const int p = p_begin + (phase + it * phases) * G;
        if (p + G <= p_end) {
#pragma unroll
          for (int q = 0; q < 2 * Q; ++q)
            if (q < Q * kf) dst[q] = L.keys ? __ldg(reinterpret_cast<const unsigned*>(L.keys + static_cast<size_t>(p) * kf) + q) : 0u;
        } else {  // ragged tail of the chunk: missing pixels are "dropped"
#pragma unroll
          for (int i = 0; i < 2 * G; ++i) {
            if (i < G * kf && p * kf + i < p_end * kf) {
              const unsigned k = L.keys ? L.keys[static_cast<size_t>(p) * kf + i] : 0u;
              const int sh = 8 * (i & 3);
              dst[i >> 2] = (dst[i >> 2] & ~(0xffu << sh)) | (k << sh);
            }
          }
        }
      }
    };
    unsigned kcur[2 * Q], knxt[2 * Q], kw[2 * Q];
    fetch_keys(0, kcur);
    fetch_keys(32, knxt);

    for (int s = 0; s < kStages; ++s) issue();
    int stage = 0;
    for (int it = 0; it < n_my; ++it) {
      const int j = it & 31;
      if (j == 0 && it > 0) {
#pragma unroll
        for (int q = 0; q < 2 * Q; ++q) kcur[q] = knxt[q];
        fetch_keys(it + 32, knxt);
      }
#pragma unroll
      for (int q = 0; q < 2 * Q; ++q)
        if (q < Q * kf) kw[q] = __shfl_sync(0xffffffffu, kcur[q], j);
      mbar_wait(my_bars + stage * 8, (parity_bits >> stage) & 1u);
      parity_bits ^= 1u << stage;
      const uint32_t box_lane = my_bufs + stage * kStageBytes + lane_off;
      if (fold2) {
        // pixel-pair rows: row r = pixels 2r (lanes 0-15) and 2r+1 (lanes 16-31); 8 rows = 16 key bytes = 4 words
        const bool upper = lane >= 16;
#pragma unroll
        for (int hf = 0; hf < G / 8; ++hf) {
          const unsigned w0 = kw[4 * hf], w1 = kw[4 * hf + 1], w2 = kw[4 * hf + 2], w3 = kw[4 * hf + 3];
          const unsigned k0 = w0 & 0xffu;
          if (w0 == k0 * 0x01010101u && w1 == w0 && w2 == w0 && w3 == w0) {  // 16 pixels of one class
            f2 s1a = 0, s1b = 0, s2a = 0, s2b = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              f2 a, b;
              load_px(box_lane, 8 * hf + e, a, b);
              s1a = add2(s1a, a);
              s1b = add2(s1b, b);
              s2a = fma2(a, a, s2a);
              s2b = fma2(b, b, s2b);
            }
            row_add(k0, s1a, s1b, s2a, s2b);
          } else {
#pragma unroll 1
            for (int q = 0; q < 2; ++q) {  // 4 rows = 8 pixels = 2 key words
              const unsigned wa = q ? w2 : w0, wb = q ? w3 : w1;
              const unsigned kq = wa & 0xffu;
              const int row0 = 8 * hf + 4 * q;
              if (wa == kq * 0x01010101u && wb == wa) {
                f2 s1a = 0, s1b = 0, s2a = 0, s2b = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  f2 a, b;
                  load_px(box_lane, row0 + e, a, b);
                  s1a = add2(s1a, a);
                  s1b = add2(s1b, b);
                  s2a = fma2(a, a, s2a);
                  s2b = fma2(b, b, s2b);
                }
                row_add(kq, s1a, s1b, s2a, s2b);
              } else {
#pragma unroll 1
                for (int e = 0; e < 4; ++e) {  // one row: its two pixels may belong to two classes
                  const unsigned pair = ((e < 2 ? wa : wb) >> (16 * (e & 1))) & 0xffffu;
                  const unsigned ke = pair & 0xffu, ko = pair >> 8;
                  f2 a, b;
                  load_px(box_lane, row0 + e, a, b);
                  if (ke == ko) {
                    row_add(ke, a, b, mul2(a, a), mul2(b, b));
                  } else {  // even pixel: lanes 0-15 contribute, odd pixel: lanes 16-31
                    const f2 ae = upper ? 0 : a, be = upper ? 0 : b, ao = upper ? a : 0, bo = upper ? b : 0;
                    row_add(ke, ae, be, mul2(ae, ae), mul2(be, be));
                    row_add(ko, ao, bo, mul2(ao, ao), mul2(bo, bo));
                  }
                }
              }
            }
          }
        }
      } else {
#pragma unroll
      for (int hf = 0; hf < G / 8; ++hf) {  // 8 pixels = one 64-bit key word at a time
        const unsigned w_lo = kw[2 * hf], w_hi = kw[2 * hf + 1];
        const unsigned k0 = w_lo & 0xffu;
        if (w_lo == k0 * 0x01010101u && w_hi == w_lo) {  // one class: 8 rows summed in registers, one row update
          f2 s1a = 0, s1b = 0, s2a = 0, s2b = 0;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            f2 a, b;
            load_px(box_lane, 8 * hf + e, a, b);
            s1a = add2(s1a, a);
            s1b = add2(s1b, b);
            s2a = fma2(a, a, s2a);
            s2b = fma2(b, b, s2b);
          }
          row_add(k0, s1a, s1b, s2a, s2b);
        } else {  // a class boundary inside: per quad, and pixel by pixel only in the quad that straddles it
#pragma unroll 1
          for (int q = 0; q < 2; ++q) {
            const unsigned w = q ? w_hi : w_lo;
            const unsigned kq = w & 0xffu;
            const int row0 = 8 * hf + 4 * q;
            if (w == kq * 0x01010101u) {
              f2 s1a = 0, s1b = 0, s2a = 0, s2b = 0;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                f2 a, b;
                load_px(box_lane, row0 + e, a, b);
                s1a = add2(s1a, a);
                s1b = add2(s1b, b);
                s2a = fma2(a, a, s2a);
                s2b = fma2(b, b, s2b);
              }
              row_add(kq, s1a, s1b, s2a, s2b);
            } else {
#pragma unroll 1
              for (int e = 0; e < 4; ++e) {
                f2 a, b;
                load_px(box_lane, row0 + e, a, b);
                row_add((w >> (8 * e)) & 0xffu, a, b, mul2(a, a), mul2(b, b));
              }
            }
          }
        }
      }
      }  // !fold2
      __syncwarp();
      issue();  // refill the stage just consumed
      if (++stage == kStages) stage = 0;
    }
  }
  if (cur_layer >= 0) fold();
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
  int32_t C, HW, ld, centered;
};
template <typename T, bool BWD>
__global__ void class_stats_generic_kernel(const GenericLayer L, const int K, const int nhwc, const int px_per_block) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.z;
  const int p_begin = blockIdx.y * px_per_block, p_end = min(p_begin + px_per_block, L.HW);
  if (c >= L.C) return;
  const float sc = L.scale ? L.scale[c] : 1.f;
  float sf = L.shift ? L.shift[c] : 0.f;
  if (L.centered) sf = -sf * sc;
  const T* x = reinterpret_cast<const T*>(L.x);
  const T* dy = reinterpret_cast<const T*>(L.dy);
  float a1 = 0.f, a2 = 0.f;
  unsigned cur = K;
  auto flush = [&]() {
    if (cur < static_cast<unsigned>(K)) {
      atomicAdd(&L.S1[static_cast<size_t>(cur) * L.ld + c], static_cast<double>(a1));
      atomicAdd(&L.S2[static_cast<size_t>(cur) * L.ld + c], static_cast<double>(a2));
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

// NHWC: [rows = N*HW][cols = C], box = [G px][128 channels], no swizzle (a pixel row is read with one LDS per lane)
int make_map_nhwc(CUtensorMap* map, const void* base, int dtype, long long rows, long long cols, int box_bytes) {
  EncodeTiledFn enc = encode_tiled_fn();
  DCFP_REQUIRE(enc != nullptr, DCFP_EUNSUPPORTED, "class_stats: cuTensorMapEncodeTiled is not available in this driver");
  const size_t es = dtype == DCFP_F32 ? 4 : 2;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols) * es};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kNhwcSlab), static_cast<cuuint32_t>(box_bytes / (kNhwcSlab * es))};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(map, dtype == DCFP_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                         const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCFP_REQUIRE(r == CUDA_SUCCESS, DCFP_EINVAL, "class_stats: cuTensorMapEncodeTiled (NHWC) failed (CUresult %d) rows=%lld cols=%lld",
               static_cast<int>(r), rows, cols);
  return 0;
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
  const int warps = shared_acc ? kWarpsShared : kWarpsPrivate;
  const int copies = shared_acc ? 1 : warps;
  return static_cast<size_t>(warps) * stages * (bwd ? 2 : 1) * kBoxBytes + static_cast<size_t>(copies) * K * 32 * 8 +
         8 * warps * stages + 1024 /* base alignment slack */;
}

int validate(const dcfp_layer_desc& d, int idx) {
  DCFP_REQUIRE(d.x && d.S1 && d.S2, DCFP_EINVAL, "class_stats[%d]: x/S1/S2 must be non-null", idx);
  DCFP_REQUIRE(d.N > 0 && d.C > 0 && d.h > 0 && d.w > 0, DCFP_EINVAL, "class_stats[%d]: bad extent N=%d C=%d h=%d w=%d", idx,
               d.N, d.C, d.h, d.w);
  DCFP_REQUIRE(d.K >= 1 && d.K <= DCFP_MAX_CLASSES, DCFP_ETOOBIG, "class_stats[%d]: K=%d outside [1,%d]", idx, d.K,
               DCFP_MAX_CLASSES);
  DCFP_REQUIRE(d.dtype == DCFP_F32 || d.dtype == DCFP_BF16, DCFP_EINVAL, "class_stats[%d]: unknown dtype %d", idx, d.dtype);
  DCFP_REQUIRE(d.layout == DCFP_NCHW || d.layout == DCFP_NHWC, DCFP_EINVAL, "class_stats[%d]: unknown layout %d", idx, d.layout);
  DCFP_REQUIRE(d.ld == 0 || d.ld >= d.C, DCFP_EINVAL, "class_stats[%d]: ld=%d < C=%d", idx, d.ld, d.C);
  DCFP_REQUIRE(d.affine_mode == DCFP_AFFINE_SCALE_SHIFT || d.affine_mode == DCFP_AFFINE_INVSTD_MEAN, DCFP_EINVAL,
               "class_stats[%d]: unknown affine_mode %d", idx, d.affine_mode);
  DCFP_REQUIRE(d.affine_mode == DCFP_AFFINE_SCALE_SHIFT || (d.scale && d.shift), DCFP_EINVAL,
               "class_stats[%d]: DCFP_AFFINE_INVSTD_MEAN needs both scale (invstd) and shift (mean)", idx);
  DCFP_REQUIRE(d.reserved == 0, DCFP_EINVAL, "class_stats[%d]: reserved field must be 0", idx);
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

// the NHWC fast path (TMA): 16-B aligned base and row pitch, whole 4-channel vectors
bool nhwc_ok(const dcfp_layer_desc& d) {
  if (d.layout != DCFP_NHWC) return false;
  const size_t es = d.dtype == DCFP_F32 ? 4 : 2;
  if (d.C % 4 != 0 || (static_cast<size_t>(d.C) * es) % 16 != 0) return false;
  if (reinterpret_cast<uintptr_t>(d.x) % 16 != 0) return false;
  if (d.dy && reinterpret_cast<uintptr_t>(d.dy) % 16 != 0) return false;
  if (d.keys && reinterpret_cast<uintptr_t>(d.keys) % 4 != 0) return false;
  if (static_cast<long long>(d.N) * d.h * d.w < 64) return false;  // tiny pooled maps: generic
  return true;
}

template <typename T, bool BWD, int MAXL>
int run_nhwc(const dcfp_layer_desc* descs, const int* which, int n, long long target_bytes, cudaStream_t stream) {
  constexpr int kTens = BWD ? 2 : 1;
  constexpr int kNhwcBoxBytes = nhwc_box_bytes(BWD);
  constexpr int G = kNhwcBoxBytes / BoxRow<T>::kRowBytes;
  const int K = descs[which[0]].K;
  NhwcParams<MAXL, kTens> P;
  P.n_layers = n;
  P.K = K;
  P.tile_prefix[0] = 0;
  bool affine = false;
  for (int i = 0; i < n; ++i) {
    const dcfp_layer_desc& d = descs[which[i]];
    NhwcLayer& L = P.L[i];
    affine = affine || d.scale || d.shift;
    L.keys = d.keys;
    L.scale = d.scale;
    L.shift = d.shift;
    L.S1 = d.S1;
    L.S2 = d.S2;
    L.C = d.C;
    L.ld = d.ld > 0 ? d.ld : d.C;
    L.centered = d.affine_mode == DCFP_AFFINE_INVSTD_MEAN;
    L.n_px = d.N * d.h * d.w;
    // a 64-channel layer fills only half of a 128-channel slab: view it as [n_px / 2][128] (pixel-pair rows)
    L.fold2 = (d.C == 64 && L.n_px % 2 == 0) ? 1 : 0;
    if (L.fold2) {
      L.C = 128;
      L.n_px /= 2;
    }
    const int n_slabs = (L.C + kNhwcSlab - 1) / kNhwcSlab;
    int spc = 1;
    while (spc < kNhwcWarps && spc < n_slabs) spc <<= 1;
    L.spc = spc;
    L.n_slab_groups = (n_slabs + spc - 1) / spc;
    const int gran = G * (kNhwcWarps / spc);  // every phase gets whole pixel groups
    const long long row_bytes = static_cast<long long>(std::min(L.C, spc * kNhwcSlab)) * sizeof(T);
    long long px = std::max<long long>(target_bytes / row_bytes, gran);
    px = (px + gran - 1) / gran * gran;
    L.px_per_chunk = static_cast<int>(std::min<long long>(px, (static_cast<long long>(L.n_px) + gran - 1) / gran * gran));
    L.n_chunks = (L.n_px + L.px_per_chunk - 1) / L.px_per_chunk;
    const long long tiles = static_cast<long long>(L.n_chunks) * L.n_slab_groups;
    DCFP_REQUIRE(P.tile_prefix[i] + tiles < (1LL << 31), DCFP_ETOOBIG, "class_stats: too many tiles");
    P.tile_prefix[i + 1] = P.tile_prefix[i] + static_cast<int>(tiles);
    int rc = make_map_nhwc(&P.maps[i * kTens], d.x, d.dtype, L.n_px, L.C, kNhwcBoxBytes);
    if (rc == 0 && BWD) rc = make_map_nhwc(&P.maps[i * kTens + 1], d.dy, d.dtype, L.n_px, L.C, kNhwcBoxBytes);
    if (rc) return rc;
  }
  const int n_tiles = P.tile_prefix[n];
  if (n_tiles == 0) return 0;
  static const int forced = []() {
    const char* e = getenv("DCFP_K1_NHWC_STAGES");
    return e ? atoi(e) : 0;
  }();
  P.stages = kNhwcStageBudget / (kTens * kNhwcBoxBytes);
  if (forced >= 1 && forced * kTens * kNhwcBoxBytes <= (20 << 10)) P.stages = forced;
  const size_t smem = static_cast<size_t>(kNhwcWarps) * P.stages * kTens * kNhwcBoxBytes +
                      static_cast<size_t>(kNhwcWarps) * kNhwcSlots * 256 * sizeof(float) + 8 * kNhwcWarps * P.stages +
                      1024 /* base alignment slack */;
  void (*kern)(NhwcParams<MAXL, kTens>) = class_stats_nhwc_kernel<T, BWD, true, MAXL>;
  if (!BWD && !affine) kern = class_stats_nhwc_kernel<T, BWD, false, MAXL>;
  int rc = ensure_smem(reinterpret_cast<const void*>(kern), static_cast<int>(smem));
  if (rc) return rc;
  kern<<<std::min(n_tiles, kNumSMs), kNhwcWarps * 32, smem, stream>>>(P);  // persistent: one CTA per SM
  return finish_launch("class_stats_nhwc");
}

template <typename T, bool BWD>
int launch_generic(const dcfp_layer_desc& d, cudaStream_t stream) {
  GenericLayer L{d.x, d.dy, d.keys, d.scale, d.shift, d.S1, d.S2, d.C, d.h * d.w, d.ld > 0 ? d.ld : d.C,
                 d.affine_mode == DCFP_AFFINE_INVSTD_MEAN};
  const int threads = 128;
  const int px_per_block = 256;
  dim3 grid((d.C + threads - 1) / threads, (L.HW + px_per_block - 1) / px_per_block, d.N);
  class_stats_generic_kernel<T, BWD><<<grid, threads, 0, stream>>>(L, d.K, d.layout == DCFP_NHWC, px_per_block);
  return finish_launch("class_stats_generic");
}

// pipeline depth: 2 stages per warp keeps the most warps resident (measured best); override for tuning
int pick_stages() {
  static const int forced = []() {
    const char* e = getenv("DCFP_K1_STAGES");
    return e ? atoi(e) : 0;
  }();
  return (forced >= 2 && forced <= 8) ? forced : 2;
}

template <typename T, bool BWD, bool AFFINE, bool SHARED_ACC, int MAXL>
int launch_tiled(GroupParams<MAXL, BWD ? 2 : 1>& P, int n_tiles, cudaStream_t stream) {
  constexpr int kWarpsT = SHARED_ACC ? kWarpsShared : kWarpsPrivate;
  P.stages = pick_stages();
  const size_t smem = tile_smem_bytes(P.K, BWD, SHARED_ACC, P.stages);
  auto kern = class_stats_kernel<T, BWD, AFFINE, SHARED_ACC, kWarpsT, MAXL, SHARED_ACC>;
  int rc = ensure_smem(reinterpret_cast<const void*>(kern), static_cast<int>(smem));
  if (rc) return rc;
  kern<<<n_tiles, kWarpsT * 32, smem, stream>>>(P);
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
    L.ld = d.ld > 0 ? d.ld : d.C;
    L.centered = d.affine_mode == DCFP_AFFINE_INVSTD_MEAN;
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
    if (BWD || affine) return launch_tiled<T, BWD, true, true, MAXL>(P, n_tiles, stream);
    return launch_tiled<T, BWD, false, true, MAXL>(P, n_tiles, stream);
  }
  if (BWD || affine) return launch_tiled<T, BWD, true, false, MAXL>(P, n_tiles, stream);
  return launch_tiled<T, BWD, false, false, MAXL>(P, n_tiles, stream);
}

int run(const dcfp_layer_desc* descs, int n_layers, cudaStream_t stream) {
  DCFP_REQUIRE(descs != nullptr && n_layers > 0, DCFP_EINVAL, "class_stats: no layers");
  DCFP_REQUIRE(n_layers <= DCFP_MAX_GROUP_LAYERS, DCFP_ETOOBIG, "class_stats: %d layers > %d per call", n_layers,
               DCFP_MAX_GROUP_LAYERS);
  const int K = descs[0].K, dtype = descs[0].dtype;
  const bool bwd = descs[0].dy != nullptr;
  int tiled[DCFP_MAX_GROUP_LAYERS], nhwc[DCFP_MAX_GROUP_LAYERS];
  int n_tiled = 0, n_nhwc = 0;
  long long nhwc_bytes = 0;
  long long total_boxes = 0;  // boxes x channel groups over the whole call
  const int box_px = kBoxRowBytes / (dtype == DCFP_F32 ? 4 : 2);
  for (int i = 0; i < n_layers; ++i) {
    const dcfp_layer_desc& d = descs[i];
    int rc = validate(d, i);
    if (rc) return rc;
    DCFP_REQUIRE(d.K == K && d.dtype == dtype && (d.dy != nullptr) == bwd, DCFP_EINVAL,
                 "class_stats[%d]: K / dtype / functor differ inside one group", i);
    if (nhwc_ok(d)) {
      nhwc[n_nhwc++] = i;
      nhwc_bytes += static_cast<long long>(d.N) * d.C * d.h * d.w * (dtype == DCFP_F32 ? 4 : 2);
    } else if (tiled_ok(d)) {
      tiled[n_tiled++] = i;
      total_boxes += ((static_cast<long long>(d.h) * d.w + box_px - 1) / box_px) * d.N * ((d.C + 31) / 32);
    } else {
      if (dtype == DCFP_F32) rc = bwd ? launch_generic<float, true>(d, stream) : launch_generic<float, false>(d, stream);
      else rc = bwd ? launch_generic<__nv_bfloat16, true>(d, stream) : launch_generic<__nv_bfloat16, false>(d, stream);
      if (rc) return rc;
    }
  }
  if (n_nhwc > 0) {
    // persistent CTAs (one per SM) each walk ~8 tiles: ~1/8 of an SM's share of the call per tile, 128 KB .. 2 MB
    long long target = std::min<long long>(2 << 20, std::max<long long>(128 << 10, nhwc_bytes / (8LL * kNumSMs)));
    const int nhwc_big = bwd ? kNhwcBigGroupBwd : kNhwcBigGroupFwd;
    for (int first = 0; first < n_nhwc;) {
      const int m = std::min(n_nhwc - first, nhwc_big);
      int rc;
      if (m <= kSmallGroup) {
        if (dtype == DCFP_F32)
          rc = bwd ? run_nhwc<float, true, kSmallGroup>(descs, nhwc + first, m, target, stream)
                   : run_nhwc<float, false, kSmallGroup>(descs, nhwc + first, m, target, stream);
        else
          rc = bwd ? run_nhwc<__nv_bfloat16, true, kSmallGroup>(descs, nhwc + first, m, target, stream)
                   : run_nhwc<__nv_bfloat16, false, kSmallGroup>(descs, nhwc + first, m, target, stream);
      } else if (dtype == DCFP_F32) {
        rc = bwd ? run_nhwc<float, true, kNhwcBigGroupBwd>(descs, nhwc + first, m, target, stream)
                 : run_nhwc<float, false, kNhwcBigGroupFwd>(descs, nhwc + first, m, target, stream);
      } else {
        rc = bwd ? run_nhwc<__nv_bfloat16, true, kNhwcBigGroupBwd>(descs, nhwc + first, m, target, stream)
                 : run_nhwc<__nv_bfloat16, false, kNhwcBigGroupFwd>(descs, nhwc + first, m, target, stream);
      }
      if (rc) return rc;
      first += m;
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
