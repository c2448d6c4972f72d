// K1 -- label-keyed segmented reduction over conv/BN feature maps (sm_100a): dispatcher, generic kernel, label keys.
//
//   S1[k][c] += sum_{pixels p of class k} v(p, c),      S2[k][c] += sum v(p, c)^2
//
// HBM-bound: every feature-map byte is read exactly once; there is no dense contraction, so no tensor cores.
// Design:
//
//  * `dcfp_label_keys` nearest-down-samples the label map in registers ONCE per label resolution into a compact uint8
//    class-key plane (and counts pixels per class).  Doing it inside the reduction was measured at 27 % of all issued
//    instructions, repeated by every channel group of every layer.
//  * channels_last maps -> k1_nhwc.cuh (TMA [G px x 128 ch] boxes, lane = 4 channels, state-free accumulation into a
//    per-warp slot cache, persistent CTAs);  NCHW maps -> k1_nchw.cuh (TMA [32 ch x 128 B] boxes read back transposed,
//    lane = channel, per-warp class tables);  tiny / unaligned maps -> the generic kernel below.
//  * `dcfp_class_stats_grouped` runs many resident layers in ONE launch per path: layer table and tensor maps travel in
//    kernel parameter space.  CTA partials reach the fp64 arena with RED.F64: the cross-CTA / cross-image combine is fp64.
#include <math.h>

#include <algorithm>
#include <cstdlib>

#include "k1_nchw.cuh"
#include "k1_nhwc.cuh"

namespace dcfp {
namespace {

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
  DCFP_REQUIRE((d.hints & ~DCFP_HINT_KEEP_L2) == 0, DCFP_EINVAL, "class_stats[%d]: unknown bits in hints (%d)", idx, d.hints);
  DCFP_REQUIRE(d.keys != nullptr || d.K == 1, DCFP_EINVAL, "class_stats[%d]: keys == NULL requires K == 1", idx);
  DCFP_REQUIRE(static_cast<long long>(d.h) * d.w < (1LL << 30), DCFP_ETOOBIG, "class_stats[%d]: plane too large", idx);
  DCFP_REQUIRE(static_cast<long long>(d.N) * d.C < (1LL << 31), DCFP_ETOOBIG, "class_stats[%d]: too many planes", idx);
  return 0;
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
    static const int tiles_per_cta = []() {
      const char* e = getenv("DCFP_K1_NHWC_TILES_PER_CTA");
      const int v = e ? atoi(e) : 0;
      return v >= 1 && v <= 64 ? v : 8;
    }();
    NhwcPlan plan;
    plan.target_bytes = std::min<long long>(8 << 20, std::max<long long>(128 << 10, nhwc_bytes / (static_cast<long long>(tiles_per_cta) * num_sms())));
    // a lone layer (per-layer launches): exactly one tile per persistent CTA
    plan.single_wave = n_nhwc == 1;
    plan.keep_l2 = (descs[nhwc[0]].hints & DCFP_HINT_KEEP_L2) != 0;
    const int nhwc_big = bwd ? kNhwcBigGroupBwd : kNhwcBigGroupFwd;
    // forward functor on bf16 maps: 16-warp CTAs (k1_nhwc.cuh) -- measured [B200], all c2 layers of 2 images in one grouped
    // launch: bf16 51 % -> 80 % of the HBM roofline, fp32 95 % -> 94 % (already bandwidth-bound: stays on 8 warps); at
    // K = 150 neither kernel is bound by its streaming loop (35 % both).  DCFP_K1_FWD_WARPS=8|16 forces one kernel for
    // both dtypes (A/B measurements, and the GPU suite's second pass over the forward cases).
    static const int forced_warps = []() {
      const char* e = getenv("DCFP_K1_FWD_WARPS");
      return e ? atoi(e) : 0;
    }();
    const int fwd_warps = forced_warps ? forced_warps : (dtype == DCFP_BF16 ? kNhwcFwdWarpsBf16 : kNhwcWarps);
    for (int first = 0; first < n_nhwc;) {
      const int m = std::min(n_nhwc - first, nhwc_big);
      int rc;
      if (!bwd && fwd_warps == 16) {
        if (m <= kSmallGroup)
          rc = dtype == DCFP_F32 ? run_nhwc<float, false, kSmallGroup, 0, 16>(descs, nhwc + first, m, plan, stream)
                                 : run_nhwc<__nv_bfloat16, false, kSmallGroup, 0, 16>(descs, nhwc + first, m, plan, stream);
        else
          rc = dtype == DCFP_F32 ? run_nhwc<float, false, kNhwcBigGroupFwd, 0, 16>(descs, nhwc + first, m, plan, stream)
                                 : run_nhwc<__nv_bfloat16, false, kNhwcBigGroupFwd, 0, 16>(descs, nhwc + first, m, plan, stream);
      } else if (m <= kSmallGroup) {
        if (dtype == DCFP_F32)
          rc = bwd ? run_nhwc<float, true, kSmallGroup>(descs, nhwc + first, m, plan, stream)
                   : run_nhwc<float, false, kSmallGroup>(descs, nhwc + first, m, plan, stream);
        else
          rc = bwd ? run_nhwc<__nv_bfloat16, true, kSmallGroup>(descs, nhwc + first, m, plan, stream)
                   : run_nhwc<__nv_bfloat16, false, kSmallGroup>(descs, nhwc + first, m, plan, stream);
      } else if (dtype == DCFP_F32) {
        rc = bwd ? run_nhwc<float, true, kNhwcBigGroupBwd>(descs, nhwc + first, m, plan, stream)
                 : run_nhwc<float, false, kNhwcBigGroupFwd>(descs, nhwc + first, m, plan, stream);
      } else {
        rc = bwd ? run_nhwc<__nv_bfloat16, true, kNhwcBigGroupBwd>(descs, nhwc + first, m, plan, stream)
                 : run_nhwc<__nv_bfloat16, false, kNhwcBigGroupFwd>(descs, nhwc + first, m, plan, stream);
      }
      if (rc) return rc;
      first += m;
    }
  }
  if (n_tiled == 0) return 0;
  // chunk length: ~512 KB per CTA, shortened while the call cannot fill ~4 waves of 4 CTAs/SM
  int chunk = kTargetBoxesPerChunk;
  while (chunk > 16 && total_boxes / chunk < 4LL * 4 * num_sms()) chunk >>= 1;

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

// ---- internal entry points used by bn_fused.cu ---------------------------------------------------------------------
// BN backward, first pass: class-keyed S1/S2 of v = dz * xhat, the per-channel totals (sum dz, sum dz * xhat) into the
// scratch stripes (bn_dx_kernel's prologue turns those into dgamma / dbeta and the dx coefficients, bn_common.cuh)
int k1_run_bn_backward(const dcfp_layer_desc& d0, const BnFinal& fin, bool relu, float* S1f, float* S2f, cudaStream_t stream) {
  dcfp_layer_desc d = d0;
  if (S1f != nullptr) {  // fp32 rows: the fp64 pointers are unused (validate() wants them non-null)
    DCFP_REQUIRE(S2f != nullptr && reinterpret_cast<uintptr_t>(S1f) % 16 == 0 && reinterpret_cast<uintptr_t>(S2f) % 16 == 0 &&
                     (d.ld > 0 ? d.ld : d.C) % 4 == 0,
                 DCFP_EINVAL, "bn_backward: fp32 class rows need 16-byte aligned S1 / S2 and ld %% 4 == 0");
    d.S1 = reinterpret_cast<double*>(S1f);
    d.S2 = reinterpret_cast<double*>(S2f);
  }
  int rc = validate(d, 0);
  if (rc) return rc;
  DCFP_REQUIRE(d.dy != nullptr && d.affine_mode == DCFP_AFFINE_INVSTD_MEAN && nhwc_ok(d), DCFP_EUNSUPPORTED,
               "bn_backward: needs a channels_last fp32/bf16 map with C %% 4 == 0, 16-byte aligned, >= 64 pixels");
  DCFP_REQUIRE(fin.scratch != nullptr && fin.gamma && fin.beta, DCFP_EINVAL, "bn_backward: null scratch / gamma / beta");
  NhwcPlan plan;
  plan.single_wave = true;
  plan.keep_l2 = (d.hints & DCFP_HINT_KEEP_L2) != 0;
  const NhwcFused F{fin.gamma, fin.beta, fin, S1f, S2f};
  const int which = 0;
  if (d.dtype == DCFP_F32)
    return relu ? run_nhwc<float, true, kSmallGroup, 2>(&d, &which, 1, plan, stream, &F)
                : run_nhwc<float, true, kSmallGroup, 1>(&d, &which, 1, plan, stream, &F);
  return relu ? run_nhwc<__nv_bfloat16, true, kSmallGroup, 2>(&d, &which, 1, plan, stream, &F)
              : run_nhwc<__nv_bfloat16, true, kSmallGroup, 1>(&d, &which, 1, plan, stream, &F);
}

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
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 4LL * num_sms()));
  label_keys_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      label, label_dtype, N, H0, W0, h, w, K, static_cast<float>(H0) / static_cast<float>(h),
      static_cast<float>(W0) / static_cast<float>(w), keys, cnt);
  return finish_launch("label_keys");
}

#ifdef DCFP_K1_TRACE
extern "C" int dcfp_debug_k1_trace(unsigned long long* host_out) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, dcfp::g_k1_trace, sizeof(unsigned long long) * 160 * 8));
}
#endif
