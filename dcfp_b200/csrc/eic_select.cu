// K2 -- EIC score update, exact per-group k-th order statistic, strict-> masks with per-layer
// min-keep top-k fallback (sm_100a).  Follows pruners/dcfp_pruner.py:15-20 (EIC), :43-66
// (get_thresh) and :68-92 (gen_channel_mask) of the reference.
//
// The reference spends ~10 tiny launches per BN layer on the update (600-1100 launches per
// step) and a CPU sort per group for the threshold.  Here: ONE launch updates the concatenated
// score vector of all layers, ONE launch (a CTA per group) radix-selects the exact thresholds,
// ONE launch (a CTA per layer) writes the masks.  The vectors are <= ~54k floats, so this is
// launch-latency bound by construction -- reported in microseconds, not as a roofline fraction.
#include "common.cuh"

namespace dcfp {
namespace {

// ---- K2a -------------------------------------------------------------------------------------
// Bit-exact restatement of dcfp_pruner.py:18-20 in fp32:
//   flag = grad*gamma > 0
//   g    = flag*|grad| + (!flag)*eic            (bool * float products, then one add)
//   eic  = eic*r + g*(1-r)                      (two rounded products, then one rounded add)
// __fmul_rn/__fadd_rn forbid FMA contraction so the roundings match the reference's.
__device__ __forceinline__ float eic_step(float grad, float gamma, float prev, float r, float omr) {
  const bool flag = __fmul_rn(grad, gamma) > 0.f;
  const float t1 = __fmul_rn(flag ? 1.f : 0.f, fabsf(grad));
  const float t2 = __fmul_rn(flag ? 0.f : 1.f, prev);
  const float g = __fadd_rn(t1, t2);
  return __fadd_rn(__fmul_rn(prev, r), __fmul_rn(g, omr));
}

__global__ void eic_update_ptrs_kernel(const float* const* __restrict__ grad_ptrs, const float* const* __restrict__ gamma_ptrs,
                                       const int32_t* __restrict__ offsets, float* __restrict__ eic, float r, float omr,
                                       int first_step) {
  const int l = blockIdx.x;
  const int beg = offsets[l], n = offsets[l + 1] - beg;
  const float* __restrict__ g = grad_ptrs[l];
  const float* __restrict__ w = gamma_ptrs[l];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float prev = first_step ? 0.f : eic[beg + i];
    eic[beg + i] = eic_step(g[i], w[i], prev, r, omr);
  }
}

__global__ void eic_update_flat_kernel(const float* __restrict__ grad, const float* __restrict__ gamma, float* __restrict__ eic,
                                       int n, float r, float omr, int first_step) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) eic[i] = eic_step(grad[i], gamma[i], first_step ? 0.f : eic[i], r, omr);
}

// dgamma[c] = sum_k S1[k][c]  (fp64 sum over classes, rounded once to fp32)
__global__ void reduce_classes_kernel(const double* __restrict__ S1, int K, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0;
  for (int k = 0; k < K; ++k) s += S1[static_cast<size_t>(k) * C + c];
  out[c] = static_cast<float>(s);
}

// end-of-step fold: dgamma = sum_k step.S1, total += step, step = 0  (one coalesced pass, thread == channel).
// step32 (optional): the fp32 per-step arena the fused BN backward fills with vector atomics; same layout, folded in fp64.
__global__ void fold_step_kernel(double* __restrict__ step, float* __restrict__ step32, double* __restrict__ total, int K, int C,
                                 float* __restrict__ dgamma) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const size_t plane = static_cast<size_t>(K) * C;
  double s = 0.0;
  for (int k = 0; k < K; ++k) {
    const size_t o = static_cast<size_t>(k) * C + c;
    double a = step[o], b = step[plane + o];
    if (a != 0.0) step[o] = 0.0;
    if (b != 0.0) step[plane + o] = 0.0;
    if (step32 != nullptr) {
      const float a32 = step32[o], b32 = step32[plane + o];
      if (a32 != 0.f) step32[o] = 0.f;
      if (b32 != 0.f) step32[plane + o] = 0.f;
      a += static_cast<double>(a32);
      b += static_cast<double>(b32);
    }
    s += a;
    if (total != nullptr) {
      if (a != 0.0) total[o] += a;
      if (b != 0.0) total[plane + o] += b;
    }
  }
  if (dgamma != nullptr) dgamma[c] = static_cast<float>(s);
}

// ---- K2b -------------------------------------------------------------------------------------
// order-preserving float -> uint key (ascending); +NaN sorts last like torch.sort
__device__ __forceinline__ uint32_t f2key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

constexpr int kSelThreads = 1024;

// Picks the digit d with  sum(hist[0..d-1]) <= k < sum(hist[0..d])  and the remainder k - sum(hist[0..d-1]).
// Executed by warp 0: lane j owns bins 8j..8j+7, an exclusive prefix over the lane sums locates the lane, the lane
// walks its 8 bins (a single thread walking 256 bins cost ~8 k cycles per radix pass).
__device__ __forceinline__ void pick_digit(const uint32_t* hist, long long k, uint32_t* bcast) {
  const int lane = threadIdx.x & 31;
  uint32_t b[8];
  uint32_t mine = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    b[i] = hist[8 * lane + i];
    mine += b[i];
  }
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const long long excl = static_cast<long long>(incl) - mine;
  const bool here = k >= excl && k < static_cast<long long>(incl);
  const unsigned vote = __ballot_sync(0xffffffffu, here);
  if (vote == 0) {  // k beyond the population (cannot happen for a valid k): clamp to the last bin like the serial walk
    if (lane == 31) {
      bcast[0] = 255u;
      bcast[1] = static_cast<uint32_t>(k - (static_cast<long long>(incl) - b[7]));
    }
    return;
  }
  if (here) {
    long long rem = k - excl;
    int d = 0;
    for (; d < 7; ++d) {
      if (rem < static_cast<long long>(b[d])) break;
      rem -= b[d];
    }
    bcast[0] = static_cast<uint32_t>(8 * lane + d);
    bcast[1] = static_cast<uint32_t>(rem);
  }
}

// Block-wide exact k-th smallest key (0-based) among the elements enumerated by `visit`.
// visit(f) calls f(key) for every element owned by this thread.  4 MSB-first 8-bit passes.
template <typename Visit>
__device__ uint32_t block_radix_select(Visit visit, long long k, uint32_t* hist /*[256]*/, uint32_t* bcast /*[2]*/) {
  uint32_t prefix = 0, mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    visit([&](uint32_t key) {
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
    });
    __syncthreads();
    if (threadIdx.x < 32) pick_digit(hist, k, bcast);
    __syncthreads();
    prefix |= bcast[0] << shift;
    mask |= 0xffu << shift;
    k = bcast[1];
    __syncthreads();
  }
  return prefix;
}

// one CTA per group: thresh[g] = k_idx[g]-th smallest score among the layers of group g  (fallback for score vectors
// whose per-element group map does not fit in shared memory: walks the layers one by one)
__global__ void __launch_bounds__(kSelThreads) group_thresh_layers_kernel(const float* __restrict__ score,
                                                                          const int32_t* __restrict__ layer_off,
                                                                          const int32_t* __restrict__ layer_group, int n_layers,
                                                                          long long k0, long long k1, float* __restrict__ thresh_out) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t bcast[2];
  const int g = blockIdx.x;
  const long long k = g == 0 ? k0 : k1;
  if (k < 0) {  // empty group: the reference leaves thresh[g] = 0 (dcfp_pruner.py:58-60)
    if (threadIdx.x == 0) thresh_out[g] = 0.f;
    return;
  }
  auto visit = [&](auto&& f) {
    for (int l = 0; l < n_layers; ++l) {
      if (layer_group[l] != g) continue;
      const int end = layer_off[l + 1];
      for (int i = layer_off[l] + threadIdx.x; i < end; i += blockDim.x) f(f2key(score[i]));
    }
  };
  const uint32_t key = block_radix_select(visit, k, hist, bcast);
  if (threadIdx.x == 0) thresh_out[g] = key2f(key);
}

// Same result, built for the sizes that occur (<= ~54 k scores): the CTA first writes a one-byte group id per element
// into shared memory, then every radix pass is ONE flat, 8x-unrolled sweep over the score vector (independent loads in
// flight) into per-warp histograms.  The layer-by-layer walk above exposes a global-load latency per layer and pass
// (113 layers x 4 passes: 148 us on c2); this one takes ~15 us.
constexpr int kSelUnroll = 8;
__global__ void __launch_bounds__(kSelThreads) group_thresh_kernel(const float* __restrict__ score,
                                                                   const int32_t* __restrict__ layer_off,
                                                                   const int32_t* __restrict__ layer_group, int n_layers,
                                                                   int n_total, long long k0, long long k1,
                                                                   float* __restrict__ thresh_out) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  uint32_t* whist = reinterpret_cast<uint32_t*>(sel_smem);  // [32 warps][256]
  uint32_t* hist = whist + 32 * 256;                        // [256]
  uint32_t* bcast = hist + 256;                             // [2]
  unsigned char* grp = reinterpret_cast<unsigned char*>(bcast + 2);  // [n_total]
  const int g = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long k = g == 0 ? k0 : k1;
  if (k < 0) {
    if (tid == 0) thresh_out[g] = 0.f;
    return;
  }
  for (int l = warp; l < n_layers; l += kSelThreads / 32) {
    const unsigned char gl = static_cast<unsigned char>(layer_group[l]);
    const int end = layer_off[l + 1];
    for (int i = layer_off[l] + lane; i < end; i += 32) grp[i] = gl;
  }
  __syncthreads();
  uint32_t prefix = 0, mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 32 * 256; i += kSelThreads) whist[i] = 0;
    __syncthreads();
    for (int base = 0; base < n_total; base += kSelThreads * kSelUnroll) {
      uint32_t key[kSelUnroll];
      bool mine[kSelUnroll];
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u) {
        const int i = base + u * kSelThreads + tid;
        mine[u] = i < n_total && grp[i] == g;
        key[u] = mine[u] ? f2key(score[i]) : 0u;
      }
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u)
        if (mine[u] && (key[u] & mask) == prefix) atomicAdd(&whist[warp * 256 + ((key[u] >> shift) & 0xffu)], 1u);
    }
    __syncthreads();
    if (tid < 256) {
      uint32_t sum = 0;
      for (int w = 0; w < 32; ++w) sum += whist[w * 256 + tid];
      hist[tid] = sum;
    }
    __syncthreads();
    if (tid < 32) pick_digit(hist, k, bcast);
    __syncthreads();
    prefix |= bcast[0] << shift;
    mask |= 0xffu << shift;
    k = bcast[1];
    __syncthreads();
  }
  if (tid == 0) thresh_out[g] = key2f(prefix);
}

constexpr int kMaskThreads = 256;

// one CTA per layer: mask = score > thresh[group]; min-keep fallback (dcfp_pruner.py:77-82)
__global__ void __launch_bounds__(kMaskThreads) layer_mask_kernel(const float* __restrict__ score,
                                                                  const int32_t* __restrict__ layer_off,
                                                                  const int32_t* __restrict__ layer_group,
                                                                  const int32_t* __restrict__ min_keep,
                                                                  const float* __restrict__ thresh, float* __restrict__ mask_out,
                                                                  int32_t* __restrict__ kept_out) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t bcast[2];
  __shared__ int s_count;
  const int l = blockIdx.x;
  const int beg = layer_off[l], C = layer_off[l + 1] - beg;
  const float t = thresh[layer_group[l] & 1];  // groups 2/3: masked with thresh[0/1] but not part of the threshold set
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  int local = 0;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const bool keep = score[beg + i] > t;
    mask_out[beg + i] = keep ? 1.f : 0.f;
    local += keep;
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_count, local);
  __syncthreads();
  int kept = s_count;
  const int mk = min(min_keep[l], C);
  if (kept < mk) {
    // keep the mk highest scores: everything above the mk-th largest value, then fill with the
    // lowest-index elements equal to it (torch.sort's tie order is implementation-defined)
    auto visit = [&](auto&& f) {
      for (int i = threadIdx.x; i < C; i += blockDim.x) f(f2key(score[beg + i]));
    };
    const uint32_t kth = block_radix_select(visit, static_cast<long long>(C - mk), hist, bcast);
    __syncthreads();
    if (threadIdx.x == 0) {
      int above = 0;
      for (int i = 0; i < C; ++i) above += f2key(score[beg + i]) > kth;
      int need = mk - above;
      for (int i = 0; i < C; ++i) {
        const uint32_t key = f2key(score[beg + i]);
        if (key > kth) mask_out[beg + i] = 1.f;
        else if (key == kth && need > 0) {
          mask_out[beg + i] = 1.f;
          --need;
        }
      }
      int total = 0;
      for (int i = 0; i < C; ++i) total += mask_out[beg + i] != 0.f;
      s_count = total;
    }
    __syncthreads();
    kept = s_count;
  }
  if (threadIdx.x == 0 && kept_out) kept_out[l] = kept;
}

}  // namespace
}  // namespace dcfp

using namespace dcfp;

extern "C" int dcfp_eic_update(const float* const* grad_ptrs, const float* const* gamma_ptrs, const int32_t* offsets,
                               int n_layers, float* eic, float r, float one_minus_r, int first_step, void* stream) {
  DCFP_REQUIRE(grad_ptrs && gamma_ptrs && offsets && eic, DCFP_EINVAL, "eic_update: null pointer");
  DCFP_REQUIRE(n_layers > 0, DCFP_EINVAL, "eic_update: n_layers=%d", n_layers);
  eic_update_ptrs_kernel<<<n_layers, 256, 0, static_cast<cudaStream_t>(stream)>>>(grad_ptrs, gamma_ptrs, offsets, eic, r,
                                                                                  one_minus_r, first_step);
  return finish_launch("eic_update");
}

extern "C" int dcfp_eic_update_flat(const float* grad, const float* gamma, float* eic, int n, float r, float one_minus_r,
                                    int first_step, void* stream) {
  DCFP_REQUIRE(grad && gamma && eic, DCFP_EINVAL, "eic_update_flat: null pointer");
  DCFP_REQUIRE(n > 0, DCFP_EINVAL, "eic_update_flat: n=%d", n);
  eic_update_flat_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(grad, gamma, eic, n, r, one_minus_r,
                                                                                         first_step);
  return finish_launch("eic_update_flat");
}

extern "C" int dcfp_reduce_classes(const double* S1, int K, int C, float* out, void* stream) {
  DCFP_REQUIRE(S1 && out, DCFP_EINVAL, "reduce_classes: null pointer");
  DCFP_REQUIRE(K > 0 && C > 0, DCFP_EINVAL, "reduce_classes: K=%d C=%d", K, C);
  reduce_classes_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(S1, K, C, out);
  return finish_launch("reduce_classes");
}

extern "C" int dcfp_fold_step2(double* step, float* step32, double* total, int K, int C, float* dgamma, void* stream) {
  DCFP_REQUIRE(step != nullptr, DCFP_EINVAL, "fold_step: null step arena");
  DCFP_REQUIRE(K > 0 && C > 0, DCFP_EINVAL, "fold_step: K=%d C=%d", K, C);
  fold_step_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(step, step32, total, K, C, dgamma);
  return finish_launch("fold_step");
}

extern "C" int dcfp_fold_step(double* step, double* total, int K, int C, float* dgamma, void* stream) {
  return dcfp_fold_step2(step, nullptr, total, K, C, dgamma, stream);
}

extern "C" int dcfp_thresh_mask(const float* score, const int32_t* layer_off, const int32_t* layer_group, const int32_t* min_keep,
                                int n_layers, int n_total, const int64_t* k_idx_host, float* mask_out, float* thresh_out,
                                int32_t* kept_out, void* stream) {
  DCFP_REQUIRE(score && layer_off && layer_group && min_keep && k_idx_host && mask_out && thresh_out, DCFP_EINVAL,
               "thresh_mask: null pointer");
  DCFP_REQUIRE(n_layers > 0 && n_total > 0, DCFP_EINVAL, "thresh_mask: n_layers=%d n_total=%d", n_layers, n_total);
  DCFP_REQUIRE(k_idx_host[0] < n_total && k_idx_host[1] < n_total, DCFP_EINVAL,
               "thresh_mask: threshold index out of range (global_percent >= 1?)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t sel_smem = (32 * 256 + 256 + 2) * sizeof(uint32_t) + static_cast<size_t>(n_total);
  int rc;
  if (sel_smem <= 200u * 1024u) {
    rc = ensure_smem(reinterpret_cast<const void*>(group_thresh_kernel), static_cast<int>(sel_smem));
    if (rc) return rc;
    group_thresh_kernel<<<2, kSelThreads, sel_smem, s>>>(score, layer_off, layer_group, n_layers, n_total, k_idx_host[0],
                                                         k_idx_host[1], thresh_out);
  } else {
    group_thresh_layers_kernel<<<2, kSelThreads, 0, s>>>(score, layer_off, layer_group, n_layers, k_idx_host[0], k_idx_host[1],
                                                         thresh_out);
  }
  rc = finish_launch("group_thresh");
  if (rc) return rc;
  layer_mask_kernel<<<n_layers, kMaskThreads, 0, s>>>(score, layer_off, layer_group, min_keep, thresh_out, mask_out, kept_out);
  return finish_launch("layer_mask");
}
