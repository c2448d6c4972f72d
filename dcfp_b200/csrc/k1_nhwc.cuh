// K1, channels_last feature maps: TMA [G pixels x 128 channels] boxes, lane = 4 channels, per-warp slot cache.
#pragma once
#include <algorithm>
#include <cstdlib>

#include "bn_common.cuh"
#include "k1_common.cuh"

namespace dcfp {
namespace {

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
// one staged box = [G px][128 ch]: 8 KB fp32 forward (G = 16), 4 KB bf16 forward (G = 16), 4 KB per tensor backward:
// 8 KB per pipeline stage either way, so the fixed per-iteration work (wait, key broadcast, refill) is paid per 8 KB
__host__ __device__ constexpr int nhwc_box_bytes(bool bwd, int elem_size = 4, int warps = 8) {
  return (bwd || elem_size == 2 || warps > 8) ? 4096 : 8192;
}
// 16-warp variant (forward functor): 16 warps x (2 x 4 KB stages + 6 class rows) -- twice the warps per scheduler to hide the
// instruction latency of the accumulation chain (ncu on the 8-warp bf16 forward kernel: 2 warps per scheduler at ~5.5 cycles
// per instruction each, issue active 33 %, 23 instructions per 128-channel pixel row).  The 96 row tags a CTA can publish at
// a fold are the same (16 x 6 = 8 x 12).  (Outlining the cold eviction path with __noinline__ was tried as well: hook-path
// kernels 8 000 -> 6 500 instructions, but the per-layer kernels of the fused BN backward grew and every variant needed
// 16-25 more registers for the call ABI -- dropped.)
constexpr int kNhwcTagsPerCta = kNhwcWarps * kNhwcSlots;
constexpr int kNhwcFwdWarpsBf16 = 16;  // forward functor, bf16 maps (fp32 maps: kNhwcWarps)

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
  int32_t stages;   // boxes in flight per warp
  int32_t keep_l2;  // 0: read-once stream (evict-first); 1: a second pass re-reads these maps (normal L2 priority)
  int32_t debug_skip_rows;  // DCFP_K1_DEBUG_SKIP_ROWS=1: drop the class-row atomics (timing experiments only; results wrong)
  int32_t contig;   // 1: every phase of a CTA walks one contiguous run of its chunk's pixel groups (per-layer launches)
};
// BN-backward fusion (FUSED != 0; one layer per launch): the value functor becomes v = dz * xhat with the ReLU gate
// recomputed from the forward's own z = fma(x, zscale, zshift) (FUSED == 2), and the launch also yields the two
// per-channel totals every BN backward needs before it can write dx:  sum dz and sum dz * xhat, into the scratch stripes.
struct NhwcFused {
  const float* gamma;  // [C]  z = fma(x, zs, zt) with zs = gamma * invstd, zt = fma(-mean, zs, beta): the SAME fp32
  const float* beta;   // [C]  expressions the forward evaluated, so the gate equals "forward output > 0" bit for bit
  BnFinal fin;         // scratch: the per-channel totals go to its stripes (bn_dx_kernel turns them into dgamma / dbeta / dx)
  float* S1f;          // non-null: the class rows go to this fp32 arena with 128-bit vector reductions instead of L.S1 / L.S2
  float* S2f;
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

constexpr int kNhwcStageBudget = 16 << 10;  // bytes of staging per warp (8-warp CTAs): 4 x-boxes, or 2 (x, dy) pairs; 16-warp CTAs: half

// -DDCFP_K1_TRACE (DCFP_K1_TRACE=1 python -m dcfp_b200.build): thread 0 of every CTA of the fused instantiation stamps
// %globaltimer at seven points of the kernel (scripts/k1_trace.py prints the per-phase times: this is how the merge tree
// and the prologue were found to be 30 % + 12 % of a 17 MB launch).  Compiled out otherwise.
#ifdef DCFP_K1_TRACE
__device__ unsigned long long g_k1_trace[160 * 8];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define K1_TRACE(i) do { if (FUSED && threadIdx.x == 0) g_k1_trace[blockIdx.x * 8 + (i)] = gtime(); } while (0)
#else
#define K1_TRACE(i) do { } while (0)
#endif

// FOLD2: -1 = per layer at run time (grouped launches); 0 / 1 = known at compile time (the per-layer launches of the fused BN
// backward: half of the accumulation code disappears -- the kernel is ~6 000 SASS instructions, and a launch that runs for
// 12 us starts with a cold instruction cache every time)
template <typename T, bool BWD, bool AFFINE, int MAXL, int FUSED = 0, int FOLD2 = -1, int WARPS = kNhwcWarps>
__global__ void __launch_bounds__(WARPS * 32, 1)
    class_stats_nhwc_kernel(const __grid_constant__ NhwcParams<MAXL, BWD ? 2 : 1> P, const NhwcFused F) {
  static_assert(FUSED == 0 || BWD, "the fused BN-backward functor reads x and dy");
  static_assert(WARPS == 8 || WARPS == 16, "8 or 16 warps per CTA");
  constexpr int kNhwcWarps = WARPS;                    // shadow the namespace-scope defaults inside the kernel
  constexpr int kNhwcSlots = kNhwcTagsPerCta / WARPS;  // 12 or 6 class rows per warp
  constexpr int kNhwcBoxBytes = nhwc_box_bytes(BWD, static_cast<int>(sizeof(T)), WARPS);
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

  K1_TRACE(0);
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
  const uint64_t policy = P.keep_l2 ? policy_evict_normal() : policy_evict_first();

  if (lane < kStages) mbar_init(my_bars + lane * 8, 1);
  for (int i = 0; i < kNhwcSlots * 2; ++i) mine[i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
  mbar_fence_init();
  __syncwarp();
  unsigned parity_bits = 0;  // bit s = parity of the next completion of stage s

  int layer = 0, cur_layer = -1, cur_sg = -1;
  int phases = kNhwcWarps, phase = 0, c0 = 0, col0 = 0;
  bool lane_on = false, fold2 = FOLD2 > 0;
  double* out1 = nullptr;  // &S1[c0], &S2[c0] of the current layer
  double* out2 = nullptr;
  size_t ld = 0;
  f2 sc01 = pack2(1.f, 1.f), sc23 = sc01, sf01 = 0, sf23 = 0;
  f2 zs01 = 0, zs23 = 0, zt01 = 0, zt23 = 0;  // FUSED == 2: z = x * zs + zt (the forward's pre-ReLU output)
  f2 tb01 = 0, tb23 = 0, tg01 = 0, tg23 = 0;  // FUSED: per-lane totals  sum dz,  sum dz * xhat  of the current slab
  // slot cache: lane i (< kNhwcSlots) holds the class of row i (kFree = none); round-robin victim; last hit
  constexpr unsigned kFree = 0xffffffffu;
  unsigned my_tag = kFree, used = 0;  // `used`: CLOCK reference bits of the rows (warp-uniform)
  int victim = 0;
  unsigned last_key = 0xffffffffu;
  int last_slot = 0;
  const uint32_t mine_u32 = smem_u32(mine);

  // one class row (the lane's 4 channels: S1 in a, S2 in b) -> arena.  fp32 vector reductions into the per-step fp32 arena
  // of the fused BN backward, fp64 scalar atomics otherwise.  Called by all lanes of a warp.
  auto emit_row = [&](unsigned cls, float4 a, float4 b) {
    if (fold2) {  // lanes 16-31 hold the same 64 channels (odd pixels): one atomic per channel, not two
      a.x += __shfl_xor_sync(0xffffffffu, a.x, 16);
      a.y += __shfl_xor_sync(0xffffffffu, a.y, 16);
      a.z += __shfl_xor_sync(0xffffffffu, a.z, 16);
      a.w += __shfl_xor_sync(0xffffffffu, a.w, 16);
      b.x += __shfl_xor_sync(0xffffffffu, b.x, 16);
      b.y += __shfl_xor_sync(0xffffffffu, b.y, 16);
      b.z += __shfl_xor_sync(0xffffffffu, b.z, 16);
      b.w += __shfl_xor_sync(0xffffffffu, b.w, 16);
    }
    if (!lane_on || (fold2 && lane >= 16) || P.debug_skip_rows) return;
    if (FUSED && F.S1f != nullptr) {
      const size_t o = cls * ld + static_cast<size_t>(c0);
      if (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f)
        asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(F.S1f + o), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
      if (b.x != 0.f || b.y != 0.f || b.z != 0.f || b.w != 0.f)
        asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(F.S2f + o), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
      return;
    }
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
  };
  auto row_to_arena = [&](int slot, unsigned cls) {  // the warp's own row -> arena, row zeroed
    emit_row(cls, mine[slot * 64], mine[slot * 64 + 32]);
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
      if (FUSED == 2) {  // dz = dy where the forward's ReLU let the value through (z > 0, strict -- threshold_backward)
        const f2 za = fma2(a, zs01, zt01), zb = fma2(b, zs23, zt23);
        da = pack2(lo2(za) > 0.f ? lo2(da) : 0.f, hi2(za) > 0.f ? hi2(da) : 0.f);
        db = pack2(lo2(zb) > 0.f ? lo2(db) : 0.f, hi2(zb) > 0.f ? hi2(db) : 0.f);
      }
      a = mul2(da, fma2(a, sc01, sf01));
      b = mul2(db, fma2(b, sc23, sf23));
      if (FUSED) {
        tb01 = add2(tb01, da);
        tb23 = add2(tb23, db);
        tg01 = add2(tg01, a);
        tg23 = add2(tg23, b);
      }
    } else if (AFFINE) {
      a = fma2(a, sc01, sf01);
      b = fma2(b, sc23, sf23);
    }
  };
  // End of a (layer, slab group): rows -> arena.  The warps of a CTA that share a slab (phases > 1) combine their rows
  // first, so the arena sees ONE reduction per (class, channel) and CTA instead of one per warp (same-address atomics
  // serialise in the L2 slice; a per-layer launch pays them in full at its tail).  ONE round: every warp publishes the
  // class tags of its 12 rows, and after a single CTA barrier warp `phase` owns the classes c = phase (mod phases): it finds
  // them among the <= 96 published tags with three ballots per class, sums those rows straight out of the other warps'
  // slot memory and sends the result to the arena -- all warps emit in parallel.  (A pairwise tree over the phases -- three
  // rounds of barrier + row_add + re-zero -- was 30 % of a 17 MB launch: scripts/k1_trace.py.)  Every warp of the CTA
  // reaches fold() at the same tile boundaries, so the barriers are uniform.  Publishing area: the warp's own staging
  // buffer (idle here).  last: no tile follows -- the rows need not be zeroed again.
  auto fold = [&](bool last) {
    K1_TRACE(3);
    if (phases > 1) {
      const int spc_ = kNhwcWarps / phases;
      auto pub = [&](int w) { return reinterpret_cast<unsigned*>(smem + static_cast<size_t>(w) * kStages * kStageBytes); };
      if (lane < kNhwcSlots) pub(warp)[lane] = direct ? (static_cast<unsigned>(lane) < K ? static_cast<unsigned>(lane) : kFree) : my_tag;
      if (FUSED) {
        f2* t = reinterpret_cast<f2*>(pub(warp) + 32) + lane * 4;
        t[0] = tb01, t[1] = tb23, t[2] = tg01, t[3] = tg23;
      }
      __syncthreads();
      const int base_w = warp % spc_;
      unsigned tg[3];  // position p = member * 12 + slot of the group's published tags; lane holds p = lane, lane + 32, lane + 64
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int p = lane + 32 * j;
        tg[j] = p < phases * kNhwcSlots ? pub(base_w + (p / kNhwcSlots) * spc_)[p % kNhwcSlots] : kFree;
      }
      for (unsigned c = static_cast<unsigned>(phase); c < K; c += static_cast<unsigned>(phases)) {
        unsigned m[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) m[j] = __ballot_sync(0xffffffffu, tg[j] == c);
        if ((m[0] | m[1] | m[2]) == 0u) continue;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          unsigned mask = m[j];
          while (mask) {
            const int p = __ffs(mask) - 1 + 32 * j;
            mask &= mask - 1;
            const float4* row = reinterpret_cast<const float4*>(slots + static_cast<size_t>(base_w + (p / kNhwcSlots) * spc_) * kNhwcSlots * 256) +
                                (p % kNhwcSlots) * 64 + lane;
            const float4 u = row[0], v = row[32];
            a.x += u.x, a.y += u.y, a.z += u.z, a.w += u.w;
            b.x += v.x, b.y += v.y, b.z += v.z, b.w += v.w;
          }
        }
        emit_row(c, a, b);
      }
      if (FUSED) {  // the totals of the group: its first warp sums them
        if (phase == 0) {
          for (int mbr = 1; mbr < phases; ++mbr) {
            const f2* t = reinterpret_cast<const f2*>(pub(base_w + mbr * spc_) + 32) + lane * 4;
            tb01 = add2(tb01, t[0]);
            tb23 = add2(tb23, t[1]);
            tg01 = add2(tg01, t[2]);
            tg23 = add2(tg23, t[3]);
          }
        } else {
          tb01 = tb23 = tg01 = tg23 = 0;
        }
      }
      K1_TRACE(4);
    }
    if (FUSED) {
      if (fold2) {
        tb01 = add2(tb01, __shfl_xor_sync(0xffffffffu, tb01, 16));
        tb23 = add2(tb23, __shfl_xor_sync(0xffffffffu, tb23, 16));
        tg01 = add2(tg01, __shfl_xor_sync(0xffffffffu, tg01, 16));
        tg23 = add2(tg23, __shfl_xor_sync(0xffffffffu, tg23, 16));
      }
      if (lane_on && !(fold2 && lane >= 16) && (tb01 | tb23 | tg01 | tg23) != 0) {
        double* t0 = bn_stripes(F.fin.scratch) + static_cast<size_t>(blockIdx.x % kBnStripes) * 2 * F.fin.C + c0;
        double* t1 = t0 + F.fin.C;
        atomicAdd(t0 + 0, static_cast<double>(lo2(tb01)));
        atomicAdd(t0 + 1, static_cast<double>(hi2(tb01)));
        atomicAdd(t0 + 2, static_cast<double>(lo2(tb23)));
        atomicAdd(t0 + 3, static_cast<double>(hi2(tb23)));
        atomicAdd(t1 + 0, static_cast<double>(lo2(tg01)));
        atomicAdd(t1 + 1, static_cast<double>(hi2(tg01)));
        atomicAdd(t1 + 2, static_cast<double>(lo2(tg23)));
        atomicAdd(t1 + 3, static_cast<double>(hi2(tg23)));
      }
      tb01 = tb23 = tg01 = tg23 = 0;
    }
    K1_TRACE(5);
    if (phases > 1) {
      if (last) return;
      __syncthreads();  // the other warps of the group have read my rows
      for (int i = 0; i < kNhwcSlots * 2; ++i) mine[i * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes to the staging buffer, TMA writes next
    } else if (direct) {
      for (unsigned k = 0; k < K; ++k) row_to_arena(static_cast<int>(k), k);
    } else {
      for (int slot = 0; slot < kNhwcSlots; ++slot) {
        const unsigned cls = __shfl_sync(0xffffffffu, my_tag, slot);
        if (cls != kFree) row_to_arena(slot, cls);
      }
    }
    my_tag = kFree;
    victim = 0;
    used = 0;
    last_key = 0xffffffffu;
  };

  for (int tile = t_first; tile < t_last; ++tile) {
    while (tile >= P.tile_prefix[layer + 1]) ++layer;
    const NhwcLayer& L = P.L[layer];
    const CUtensorMap* maps = &P.maps[layer * kTens];
    const int t = tile - P.tile_prefix[layer];
    const int sg = t / L.n_chunks;
    const int chunk = t - sg * L.n_chunks;
    const bool new_slab = layer != cur_layer || sg != cur_sg;
    if (new_slab) {
      if (cur_layer >= 0) fold(false);
      cur_layer = layer;
      cur_sg = sg;
      const int spc = L.spc;
      phases = kNhwcWarps / spc;
      phase = warp / spc;
      col0 = (sg * spc + warp % spc) * kNhwcSlab;
      c0 = col0 + lane * 4;
      lane_on = c0 < L.C;  // C % 4 == 0: a lane's 4 channels are all inside or all outside (TMA zero-fills outside)
      fold2 = FOLD2 < 0 ? (L.fold2 != 0) : (FOLD2 > 0);
      if (fold2) c0 &= 63;  // columns 64..127 of a pixel-pair row are channels 0..63 of the odd pixel
      ld = static_cast<size_t>(L.ld);
      out1 = L.S1 + c0;
      out2 = L.S2 + c0;
    }

    const int p_begin = chunk * L.px_per_chunk;
    const int p_end = min(p_begin + L.px_per_chunk, L.n_px);
    const int n_groups = (p_end - p_begin + G - 1) / G;
    // Which pixel groups of the chunk this warp takes.  Grouped launches interleave the phases (group = phase + it * phases:
    // the warps of a CTA read adjacent boxes at the same time).  Per-layer launches (P.contig) give every phase ONE
    // contiguous run of groups instead: a warp then sees the few classes of a short stretch of one image row, not every
    // class of the chunk -- with fragmented label maps or K = 150 / 171 that is the difference between no slot-cache
    // evictions and one per few pixels (each eviction = 256 fp64 atomics).
    const int per_phase = (n_groups + phases - 1) / phases;
    const int g_first = P.contig ? min(phase * per_phase, n_groups) : phase;
    const int g_stride = P.contig ? 1 : phases;
    const int n_my = col0 >= L.C ? 0 : (P.contig ? min(per_phase, n_groups - g_first) : (n_groups - phase + phases - 1) / phases);  // a warp whose slab is empty idles

    int issue_it = 0, issue_stage = 0;
    auto issue = [&]() {  // one elected lane arms the barrier and launches the tile copies of the warp's next group
      if (issue_it < n_my) {
        if (lane == 0) {
          const uint32_t bar = my_bars + issue_stage * 8;
          const uint32_t dst = my_bufs + issue_stage * kStageBytes;
          const int p = p_begin + (g_first + issue_it * g_stride) * G;
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
        const int p = p_begin + (g_first + it * g_stride) * G;
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
    K1_TRACE(1);
    for (int s = 0; s < kStages; ++s) issue();  // the first boxes go out before anything else of the tile is fetched
    fetch_keys(0, kcur);
    fetch_keys(32, knxt);
    if (new_slab) {  // per-channel coefficients of the slab: loaded AFTER the first boxes are in flight (their global-load
                     // latency used to sit in front of the first TMA issue: ~1 us of a 17 us launch)
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
      if (FUSED == 2 && lane_on) {  // needs scale = invstd, shift = mean (DCFP_AFFINE_INVSTD_MEAN)
        float zs[4], zt[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          zs[j] = __fmul_rn(F.gamma[c0 + j], L.scale[c0 + j]);
          zt[j] = __fmaf_rn(-L.shift[c0 + j], zs[j], F.beta[c0 + j]);
        }
        zs01 = pack2(zs[0], zs[1]);
        zs23 = pack2(zs[2], zs[3]);
        zt01 = pack2(zt[0], zt[1]);
        zt23 = pack2(zt[2], zt[3]);
      }
    }
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
      if (it == 0) K1_TRACE(2);
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
  if (cur_layer >= 0) fold(true);
  K1_TRACE(6);
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

// Launch plan knobs of one NHWC call.  single_wave: every layer is cut into <= (#SMs / slab groups) chunks so that the
// whole call is ONE tile per persistent CTA (per-layer launches of the fused BN path: a second tile would restart the
// TMA pipeline and a 149th tile would double the runtime).  keep_l2: a second pass re-reads the maps.
struct NhwcPlan {
  long long target_bytes = 1 << 20;
  bool single_wave = false;
  bool keep_l2 = false;
};

template <typename T, bool BWD, int MAXL, int FUSED = 0, int WARPS = kNhwcWarps>
int run_nhwc(const dcfp_layer_desc* descs, const int* which, int n, const NhwcPlan& plan, cudaStream_t stream,
             const NhwcFused* fused = nullptr) {
  static_assert(WARPS == kNhwcWarps || (!BWD && FUSED == 0), "16-warp CTAs: forward functor only (one 4 KB box per stage)");
  constexpr int kNhwcWarps = WARPS;
  constexpr int kNhwcSlots = kNhwcTagsPerCta / WARPS;
  constexpr int kTens = BWD ? 2 : 1;
  constexpr int kNhwcBoxBytes = nhwc_box_bytes(BWD, static_cast<int>(sizeof(T)), WARPS);
  constexpr int G = kNhwcBoxBytes / BoxRow<T>::kRowBytes;
  const int K = descs[which[0]].K;
  const int sms = num_sms();
  NhwcParams<MAXL, kTens> P;
  P.n_layers = n;
  P.K = K;
  P.keep_l2 = plan.keep_l2 ? 1 : 0;
  static const int no_contig = []() {
    const char* e = getenv("DCFP_K1_NO_CONTIG");
    return e ? atoi(e) : 0;
  }();
  static const int force_contig = []() {
    const char* e = getenv("DCFP_K1_FORCE_CONTIG");
    return e ? atoi(e) : 0;
  }();
  P.contig = ((plan.single_wave && !no_contig) || force_contig) ? 1 : 0;
  static const int skip_rows = []() {
    const char* e = getenv("DCFP_K1_DEBUG_SKIP_ROWS");
    return e ? atoi(e) : 0;
  }();
  P.debug_skip_rows = skip_rows;
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
    // per-layer launches put all 8 warps of a CTA on ONE slab (they merge their rows before touching the arena: 8x fewer
    // of the scarce fp64 atomics, nothing to overlap them with at the tail of a short launch); grouped launches spread the
    // warps over up to 8 slabs (whole 4 KB rows per CTA, folds are rare there)
    static const int wide = []() {
      const char* e = getenv("DCFP_K1_SINGLE_WAVE_WIDE");
      return e ? atoi(e) : 0;
    }();
    if (!plan.single_wave || wide)
      while (spc < kNhwcWarps && spc < n_slabs) spc <<= 1;
    L.spc = spc;
    L.n_slab_groups = (n_slabs + spc - 1) / spc;
    static const int no_single = []() {
      const char* e = getenv("DCFP_K1_NO_SINGLE_WAVE");
      return e ? atoi(e) : 0;
    }();
    if (plan.single_wave && !no_single) {
      // chunks of whole boxes (G rows); the phases of a chunk may differ by one box
      const int chunks = std::max(1, sms / (L.n_slab_groups * n));
      long long px = (static_cast<long long>(L.n_px) + chunks - 1) / chunks;
      px = std::max<long long>((px + G - 1) / G * G, G);
      L.px_per_chunk = static_cast<int>(px);
    } else {
      const int gran = G * (kNhwcWarps / spc);  // every phase gets whole pixel groups
      const long long row_bytes = static_cast<long long>(std::min(L.C, spc * kNhwcSlab)) * sizeof(T);
      long long px = std::max<long long>(plan.target_bytes / row_bytes, gran);
      px = (px + gran - 1) / gran * gran;
      L.px_per_chunk = static_cast<int>(std::min<long long>(px, (static_cast<long long>(L.n_px) + gran - 1) / gran * gran));
    }
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
  constexpr int kBudget = kNhwcStageBudget * 8 / WARPS;  // staging bytes per warp
  P.stages = kBudget / (kTens * kNhwcBoxBytes);
  if (forced >= 1 && forced * kTens * kNhwcBoxBytes <= kBudget + (kBudget >> 2)) P.stages = forced;
  static const int force_keep = []() {
    const char* e = getenv("DCFP_K1_KEEP_L2");
    return e ? atoi(e) : -1;
  }();
  if (force_keep >= 0) P.keep_l2 = force_keep;
  const size_t smem = static_cast<size_t>(kNhwcWarps) * P.stages * kTens * kNhwcBoxBytes +
                      static_cast<size_t>(kNhwcWarps) * kNhwcSlots * 256 * sizeof(float) + 8 * kNhwcWarps * P.stages +
                      1024 /* base alignment slack */;
  void (*kern)(NhwcParams<MAXL, kTens>, NhwcFused) = class_stats_nhwc_kernel<T, BWD, true, MAXL, FUSED, -1, WARPS>;
  if (!BWD && !affine) kern = class_stats_nhwc_kernel<T, BWD, false, MAXL, 0, -1, WARPS>;
  if (FUSED) kern = P.L[0].fold2 ? class_stats_nhwc_kernel<T, BWD, true, MAXL, FUSED, (FUSED ? 1 : -1)>
                                 : class_stats_nhwc_kernel<T, BWD, true, MAXL, FUSED, (FUSED ? 0 : -1)>;
  NhwcFused F{};
  if (FUSED) {
    DCFP_REQUIRE(fused != nullptr && n == 1 && fused->fin.scratch != nullptr, DCFP_EINVAL, "class_stats: fused BN backward needs one layer");
    F = *fused;
  }
  int rc = ensure_smem(reinterpret_cast<const void*>(kern), static_cast<int>(smem));
  if (rc) return rc;
  kern<<<std::min(n_tiles, sms), kNhwcWarps * 32, smem, stream>>>(P, F);  // persistent: one CTA per SM
  return finish_launch("class_stats_nhwc");
}

}  // namespace
}  // namespace dcfp
