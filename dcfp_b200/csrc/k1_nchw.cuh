// K1, NCHW feature maps.
//
//  * NCHW planes are pixel-contiguous, but the class key varies along pixels and is identical across channels.  Each
//    warp therefore pulls [32 channels x 128 B] boxes into shared memory with ONE TMA tensor-tile copy
//    (cp.async.bulk.tensor.2d, SWIZZLE_128B, mbarrier completion) and reads them back TRANSPOSED -- lane == channel --
//    with conflict-free 128-bit loads.  The class key of every pixel is then warp-uniform.
//  * Accumulation is state-free (packed FADD2/FFMA2): a box whose 32/64 keys are all equal takes a branch-free path and
//    one table update; otherwise per quad (4 px, one packed key word), pixel by pixel only in a straddling quad.
//  * Every warp owns a private 2-stage pipeline (its own mbarriers): no CTA-wide synchronisation in the main loop.  A
//    CTA covers 32 channels x one pixel chunk (<= 512 KB); its 4 warps take the chunk's boxes round-robin.
//  * Tables: per-warp [K x 32] float2 copies (plain RMW) for K <= 24; for larger K the CTA scans its tile's keys once,
//    builds a class -> row remap (rank among the classes present) and every warp keeps a private 24-row table.
//  * At the end the CTA adds its partials into the fp64 arena with coalesced RED.F64 (only classes it met).
#pragma once
#include <algorithm>
#include <cstdlib>

#include "k1_common.cuh"

namespace dcfp {
namespace {

constexpr int kWarpsPrivate = 4;  // warps per CTA when every warp owns an accumulator table (small K)
constexpr int kRemapRows = 24;     // table rows per warp in the large-K mode (classes met by one tile)
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

// acc[key][lane] += (a1, a2): the shared accumulator is an interleaved float2 [K][32] table, so one
// 64-bit load / FADD2 / 64-bit store updates both moments of (class, channel).  Lanes touch
// consecutive 8-byte slots: conflict-free.
__device__ __forceinline__ void acc_add_row(uint32_t acc_lane, unsigned row, float a1, float a2) {
  const uint32_t addr = acc_lane + row * 256u;  // per-warp table: plain read-modify-write
  f2 cur;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(cur) : "r"(addr));
  cur = add2(cur, pack2(a1, a2));
  asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(cur) : "memory");
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

// SHARED_ACC (historic name; K > 24): a table with one row per class does not fit per warp (K x 256 B), but a TILE only
// meets a few classes (labels are spatially coherent).  The CTA first scans the tile's class keys into a presence
// bitmap and builds a class -> row REMAP in shared memory (rank of the class among those present); every warp then
// keeps a private 24-row table addressed through it.  Classes beyond the 24th present one (rare) go straight to the
// arena.  (Earlier large-K variants: one CTA-wide [K x 32] table behind shared atomics, 43-57 % of the roofline -- 2 x
// ~64 LSU cycles per update; a per-warp slot cache with ballot lookups, 44-54 % -- per-quad lookup cost.)
template <typename T, bool BWD, bool AFFINE, bool SHARED_ACC, int WARPS>
__device__ __forceinline__ void process_tile(const LayerDev& L, const CUtensorMap* maps, const int K, const int stages,
                                             const int tile, unsigned char* smem) {
  constexpr int kBoxPx = kBoxRowBytes / static_cast<int>(sizeof(T));  // 32 (fp32) / 64 (bf16)
  constexpr int kWords = kBoxPx / 4;                                  // packed key words (quads) per box
  constexpr int kQuadsPerGroup = kWords / kGroups;                    // 1 (fp32) / 2 (bf16)
  constexpr int kTens = BWD ? 2 : 1;
  constexpr int kAccCopies = WARPS;
  constexpr int kPairs = Elem<T>::kPairs;
  constexpr int kStageBytes = kTens * kBoxBytes;
  constexpr int kThreadsT = WARPS * 32;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = tile / L.n_cg, cg = tile - chunk * L.n_cg;
  const int n_active = min(32, L.C - cg * 32);

  // ---- shared-memory carve-up: [boxes | accumulators (float2 [copies][K][32]) | mbarriers] -------
  unsigned char* bufs = smem;  // [WARPS][stages][kTens][kBoxBytes], 1024-B aligned
  float2* acc = reinterpret_cast<float2*>(smem + static_cast<size_t>(WARPS) * stages * kStageBytes);
  const int rows = SHARED_ACC ? kRemapRows : K;  // table rows per warp
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(acc + kAccCopies * rows * 32);
  // large-K mode: class -> row remap of this tile, its inverse, the presence bitmap
  unsigned char* remap = reinterpret_cast<unsigned char*>(bars + WARPS * stages);  // [256]
  unsigned char* row_class = remap + 256;                                          // [kRemapRows]
  unsigned* present = reinterpret_cast<unsigned*>(row_class + kRemapRows);         // [8]
  constexpr unsigned kOverflow = 0xfeu;

  for (int i = tid; i < kAccCopies * rows * 32; i += kThreadsT) acc[i] = make_float2(0.f, 0.f);
  if (SHARED_ACC && tid < 8) present[tid] = 0u;
  if (tid < WARPS * stages) mbar_init(smem_u32(&bars[tid]), 1);
  mbar_fence_init();
  __syncthreads();

  const int box_begin = chunk * L.boxes_per_chunk;
  const int box_end = min(box_begin + L.boxes_per_chunk, L.n_boxes);
  const int n_my = (box_end - box_begin - warp + WARPS - 1) / WARPS;  // boxes of this warp
  const uint64_t policy = policy_evict_first();
  const uint32_t my_bufs = smem_u32(bufs + static_cast<size_t>(warp) * stages * kStageBytes);
  const uint32_t my_bars = smem_u32(&bars[warp * stages]);
  const uint32_t acc_lane = smem_u32(acc + warp * rows * 32 + lane);
  const int row0 = cg * 32;

  int n_rows = 0;
  unsigned last_key = 0xffffffffu, last_row = 0;
  auto acc_add = [&](unsigned key, float a1, float a2) {
    if (!SHARED_ACC) {
      acc_add_row(acc_lane, key, a1, a2);
      return;
    }
    if (key != last_key) {
      last_key = key;
      last_row = remap[key];
    }
    if (last_row == kOverflow) {  // more than 32 classes in this tile: the rest goes straight to the arena
      if (lane < n_active) {
        const size_t o = static_cast<size_t>(key) * L.ld + row0 + lane;
        atomicAdd(&L.S1[o], static_cast<double>(a1));
        atomicAdd(&L.S2[o], static_cast<double>(a2));
      }
      return;
    }
    acc_add_row(acc_lane, last_row, a1, a2);
  };

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
  // the first boxes are already in flight while the CTA scans its tile's class keys
  if (SHARED_ACC) {
    // (1) presence bitmap of the tile's keys (consecutive duplicates skipped), (2) rank -> remap
    const int n_words = (box_end - box_begin) * kWords;
    const unsigned none = static_cast<unsigned>(K) * 0x01010101u;
    unsigned prev = 0xffffffffu;
    for (int base = 0; base < n_words; base += kThreadsT * 8) {
      unsigned w[8];  // 8 independent loads in flight per thread: the scan costs ~one L2 latency per tile
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int wi = base + j * kThreadsT + tid;
        w[j] = none;
        if (wi < n_words) {
          const int box = box_begin + wi / kWords, word = wi % kWords;
          const int n = box / L.boxes_per_plane, b = box - n * L.boxes_per_plane;
          const int p = b * kBoxPx + 4 * word;
          if (p < L.HW) w[j] = L.keys ? __ldg(reinterpret_cast<const unsigned*>(L.keys + static_cast<size_t>(n) * L.HW + p)) : 0u;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const unsigned k = (w[j] >> (8 * e)) & 0xffu;
          if (k != prev && k < static_cast<unsigned>(K)) atomicOr(&present[k >> 5], 1u << (k & 31));
          prev = k;
        }
      }
    }
    __syncthreads();
    for (int k = tid; k < 256; k += kThreadsT) {
      const unsigned word = present[k >> 5], bit = 1u << (k & 31);
      unsigned char r = 0xffu;
      if (word & bit) {
        int rank = __popc(word & (bit - 1u));
        for (int j = 0; j < (k >> 5); ++j) rank += __popc(present[j]);
        r = rank < kRemapRows ? static_cast<unsigned char>(rank) : static_cast<unsigned char>(kOverflow);
        if (rank < kRemapRows) row_class[rank] = static_cast<unsigned char>(k);
      }
      remap[k] = r;
    }
    int total = 0;
    for (int j = 0; j < 8; ++j) total += __popc(present[j]);
    n_rows = min(total, kRemapRows);
    __syncthreads();
  }
  unsigned lw = key_word();

  // SWIZZLE_128B: the 16-B chunk index of row r is XORed with r % 8 (lane == row)
  const uint32_t row_off = static_cast<uint32_t>(lane * kBoxRowBytes);
  const uint32_t l7 = static_cast<uint32_t>(lane & 7);

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
        acc_add(key0, lo2(s1a) + hi2(s1a), lo2(s2a) + hi2(s2a));
      }
    } else {
      // a class boundary crosses the box: per quad (4 px, one packed key word) -- a quad with one
      // class is summed in registers and added to the table; a straddling quad goes pixel by pixel.
      // Large-K mode: every lane first translates ITS key word to table rows through the tile's remap (4 byte
      // look-ups per lane per box instead of one look-up per quad in the loop below).
      unsigned rw = lw;
      if (SHARED_ACC) {
        rw = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) rw |= static_cast<unsigned>(remap[(lw >> (8 * e)) & 0xffu]) << (8 * e);
      }
      const unsigned row_limit = SHARED_ACC ? static_cast<unsigned>(kRemapRows) : static_cast<unsigned>(K);
#pragma unroll 1
      for (int g = 0; g < kGroups; ++g) {
        const uint32_t gaddr = box + ((static_cast<uint32_t>(g) ^ l7) << 4);
        f2 v[kPairs];
        load_group<T, BWD, AFFINE>(gaddr, sc2, sf2, v);
#pragma unroll
        for (int h = 0; h < kQuadsPerGroup; ++h) {
          const unsigned wv = __shfl_sync(0xffffffffu, rw, g * kQuadsPerGroup + h);
          const unsigned key = wv & 0xffu;  // a table row
          if (wv == key * 0x01010101u && (!SHARED_ACC || key != kOverflow)) {
            if (key < row_limit) {
              const f2 t1 = add2(v[2 * h], v[2 * h + 1]);
              const f2 t2 = fma2(v[2 * h + 1], v[2 * h + 1], mul2(v[2 * h], v[2 * h]));
              acc_add_row(acc_lane, key, lo2(t1) + hi2(t1), lo2(t2) + hi2(t2));
            }
          } else {  // straddling quad, or (large-K mode) a quad of overflow classes: pixel by pixel, by class
            const unsigned kv = SHARED_ACC ? __shfl_sync(0xffffffffu, lw, g * kQuadsPerGroup + h) : wv;
#pragma unroll 1
            for (int e = 0; e < 4; ++e) {
              const unsigned ke = (kv >> (8 * e)) & 0xffu;
              if (ke < static_cast<unsigned>(K)) {
                const float x = load_px<T, BWD, AFFINE>(gaddr + (h * 4 + e) * static_cast<int>(sizeof(T)), sc, sf);
                acc_add(ke, x, x * x);
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
  __syncthreads();

  // ---- CTA partials -> fp64 arena (coalesced RED.F64; zero partials are skipped) ----------------
  const int used_rows = SHARED_ACC ? n_rows : K;
  for (int idx = tid; idx < used_rows * 32; idx += kThreadsT) {
    const int r = idx >> 5, cl = idx & 31;
    if (cl >= n_active) continue;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int w = 0; w < kAccCopies; ++w) {
      const float2 a = acc[w * rows * 32 + idx];
      s1 += a.x;
      s2 += a.y;
    }
    if (s1 == 0.f && s2 == 0.f) continue;  // class not met by this CTA (or all-zero values): nothing to add
    const int k = SHARED_ACC ? row_class[r] : r;
    const size_t o = static_cast<size_t>(k) * L.ld + row0 + cl;
    atomicAdd(&L.S1[o], static_cast<double>(s1));
    atomicAdd(&L.S2[o], static_cast<double>(s2));
  }
}

template <typename T, bool BWD, bool AFFINE, bool SHARED_ACC, int WARPS, int MAXL>
__global__ void __launch_bounds__(WARPS * 32, 4)
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
  process_tile<T, BWD, AFFINE, SHARED_ACC, WARPS>(P.L[lo], &P.maps[lo * (BWD ? 2 : 1)], P.K, P.stages, tile - P.tile_prefix[lo],
                                                  smem);
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
  const int warps = kWarpsPrivate;
  const int rows = shared_acc ? kRemapRows : K;  // tile-local remap vs one row per class
  return static_cast<size_t>(warps) * stages * (bwd ? 2 : 1) * kBoxBytes + static_cast<size_t>(warps) * rows * 32 * 8 +
         8 * warps * stages + (shared_acc ? 256 + kRemapRows + 32 : 0) + 1024 /* base alignment slack */;
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
  constexpr int kWarpsT = kWarpsPrivate;
  P.stages = pick_stages();
  const size_t smem = tile_smem_bytes(P.K, BWD, SHARED_ACC, P.stages);
  auto kern = class_stats_kernel<T, BWD, AFFINE, SHARED_ACC, kWarpsT, MAXL>;
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
  if (K > kPrivateAccMaxK) {  // tile-local class remap, 32 table rows per warp
    if (BWD || affine) return launch_tiled<T, BWD, true, true, MAXL>(P, n_tiles, stream);
    return launch_tiled<T, BWD, false, true, MAXL>(P, n_tiles, stream);
  }
  if (BWD || affine) return launch_tiled<T, BWD, true, false, MAXL>(P, n_tiles, stream);
  return launch_tiled<T, BWD, false, false, MAXL>(P, n_tiles, stream);
}

}  // namespace
}  // namespace dcfp
