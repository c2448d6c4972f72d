// K3 -- channel gather (deploy_subnet, pruners/channel_pruner.py:907-948) and the bias-compensation
// reduce-GEMV (resize_subnet_bias, :873-905).  Pure data movement: HBM-bound, no tensor cores.
//
//   dst[o', i', e] = src[out_idx[o'], in_idx[i'], e]        e in [0, khw)
//
// The reference does two boolean-mask advanced-indexing passes (+ .contiguous()) per module from a
// Python loop over 130-234 modules.  Here one thread block owns a run of consecutive DESTINATION
// elements (coalesced stores); the kept inputs of a kept row are monotone in memory, so the loads
// touch every 32-byte sector of the row at most once.  `dcfp_channel_gather_grouped` slices every
// module of a model in ONE launch from a device-resident table.
#include "common.cuh"

namespace dcfp {
namespace {

constexpr int kGatherThreads = 256;
constexpr int kGatherPerThread = 8;
constexpr int kGatherTile = kGatherThreads * kGatherPerThread;  // dst elements per CTA

// KHW: kernel footprint known at compile time (1 = 1x1 conv / linear / vectors, 9 = 3x3 conv; 0 = runtime value).
// One 64-bit division per THREAD locates the tile; every element then costs one 32-bit division by the row length and
// one by KHW (a multiply-shift when KHW is a constant) -- the first version divided 64-bit values twice per element and
// ran at a quarter of the HBM roofline.
template <typename E, int KHW>
__device__ __forceinline__ void gather_tile(const dcfp_gather_desc& d, long long first) {
  const int khw = KHW ? KHW : d.khw;
  const unsigned row = static_cast<unsigned>(d.n_in) * static_cast<unsigned>(khw);  // dst elements per output channel
  const long long total = static_cast<long long>(row) * d.n_out;
  const E* __restrict__ src = static_cast<const E*>(d.src);
  E* __restrict__ dst = static_cast<E*>(d.dst);
  const long long src_row = static_cast<long long>(d.I) * khw;
  const long long o_first = first / row;
  const unsigned r_first = static_cast<unsigned>(first - o_first * row);
  // phase 1: all index math and all loads (independent, predicated -- no early exit, so they are in flight together);
  // phase 2: the coalesced stores
  E v[kGatherPerThread];
  bool ok[kGatherPerThread];
#pragma unroll
  for (int u = 0; u < kGatherPerThread; ++u) {
    const unsigned local = u * kGatherThreads + threadIdx.x;
    ok[u] = first + local < total;
    v[u] = E(0);
    if (ok[u]) {
      const unsigned r = r_first + local;  // < row + tile: fits in 32 bits (row < 2^31 - tile, checked on the host)
      const unsigned q = r / row;
      const unsigned rem = r - q * row;
      const int o = static_cast<int>(o_first) + static_cast<int>(q);
      const unsigned i = rem / static_cast<unsigned>(khw), e = rem - i * static_cast<unsigned>(khw);
      const int so = d.out_idx ? __ldg(d.out_idx + o) : o;
      const int si = d.in_idx ? __ldg(d.in_idx + i) : static_cast<int>(i);
      v[u] = src[so * src_row + static_cast<long long>(si) * khw + e];
    }
  }
#pragma unroll
  for (int u = 0; u < kGatherPerThread; ++u)
    if (ok[u]) dst[first + u * kGatherThreads + threadIdx.x] = v[u];
}

template <typename E>
__device__ __forceinline__ void gather_tile_any(const dcfp_gather_desc& d, long long first) {
  if (d.khw == 1) gather_tile<E, 1>(d, first);  // CTA-uniform branch
  else if (d.khw == 9) gather_tile<E, 9>(d, first);
  else gather_tile<E, 0>(d, first);
}

template <typename E>
__global__ void __launch_bounds__(kGatherThreads) gather_kernel(const dcfp_gather_desc d) {
  gather_tile_any<E>(d, static_cast<long long>(blockIdx.x) * kGatherTile);
}

template <typename E>
__global__ void __launch_bounds__(kGatherThreads) gather_grouped_kernel(const dcfp_gather_desc* __restrict__ descs,
                                                                        const long long* __restrict__ tile_prefix, int n) {
  const long long tile = blockIdx.x;
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tile_prefix[mid] <= tile) lo = mid;
    else hi = mid;
  }
  gather_tile_any<E>(descs[lo], (tile - tile_prefix[lo]) * kGatherTile);
}

// offset[o] = sum_i act[i] * sum_e W[o][i][e]; one CTA per output channel streams the row with coalesced loads
// (one WARP per row left a [256, 2048, 3, 3] weight at 157 GB/s: 256 warps cannot cover HBM latency)
constexpr int kBiasThreads = 256;
__global__ void __launch_bounds__(kBiasThreads) bias_comp_kernel(const float* __restrict__ W, int O, int I, int khw,
                                                                 const float* __restrict__ act, float* __restrict__ offset) {
  __shared__ float partial[kBiasThreads / 32];
  const int o = blockIdx.x;
  const int row = I * khw;
  const float* __restrict__ w = W + static_cast<long long>(o) * row;
  float acc = 0.f;
  if (khw == 1) {
    for (int t = threadIdx.x; t < row; t += kBiasThreads) acc = fmaf(w[t], act[t], acc);
  } else {
    for (int t = threadIdx.x; t < row; t += kBiasThreads) acc = fmaf(w[t], act[t / khw], acc);
  }
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) partial[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float sum = 0.f;
    for (int i = 0; i < kBiasThreads / 32; ++i) sum += partial[i];
    offset[o] = sum;
  }
}

int validate_gather(const dcfp_gather_desc& d, int idx) {
  DCFP_REQUIRE(d.n_out >= 0 && d.n_in >= 0 && d.I > 0 && d.khw > 0, DCFP_EINVAL, "channel_gather[%d]: bad extents", idx);
  const bool empty = d.n_out == 0 || d.n_in == 0;
  DCFP_REQUIRE(empty || (d.src && d.dst), DCFP_EINVAL, "channel_gather[%d]: null src/dst", idx);
  // an EMPTY selection arrives as a zero-length index list, whose device pointer may legitimately be NULL
  DCFP_REQUIRE(d.in_idx != nullptr || d.n_in == d.I || d.n_in == 0, DCFP_EINVAL,
               "channel_gather[%d]: in_idx NULL requires n_in == I", idx);
  DCFP_REQUIRE(static_cast<long long>(d.n_in) * d.khw < (1LL << 31) - kGatherTile, DCFP_ETOOBIG,
               "channel_gather[%d]: output row too long", idx);
  return 0;
}

}  // namespace
}  // namespace dcfp

using namespace dcfp;

extern "C" int dcfp_channel_gather(const void* src, void* dst, const int32_t* out_idx, int n_out, const int32_t* in_idx, int n_in,
                                   int I, int khw, int elt_size, void* stream) {
  dcfp_gather_desc d{src, dst, out_idx, in_idx, n_out, n_in, I, khw};
  int rc = validate_gather(d, 0);
  if (rc) return rc;
  DCFP_REQUIRE(elt_size == 4 || elt_size == 2, DCFP_EUNSUPPORTED, "channel_gather: elt_size %d (2 or 4)", elt_size);
  const long long total = static_cast<long long>(n_out) * n_in * khw;
  if (total == 0) return 0;
  const long long blocks = (total + kGatherTile - 1) / kGatherTile;
  DCFP_REQUIRE(blocks < (1LL << 31), DCFP_ETOOBIG, "channel_gather: tensor too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (elt_size == 4) gather_kernel<uint32_t><<<static_cast<unsigned>(blocks), kGatherThreads, 0, s>>>(d);
  else gather_kernel<uint16_t><<<static_cast<unsigned>(blocks), kGatherThreads, 0, s>>>(d);
  return finish_launch("channel_gather");
}

extern "C" size_t dcfp_channel_gather_workspace(int n) {
  if (n < 0) n = 0;
  return static_cast<size_t>(n) * sizeof(dcfp_gather_desc) + static_cast<size_t>(n + 1) * sizeof(long long) + 64;
}

extern "C" int dcfp_channel_gather_grouped(const dcfp_gather_desc* descs_host, int n, int elt_size, void* desc_workspace,
                                           size_t workspace_bytes, void* stream) {
  DCFP_REQUIRE(descs_host && desc_workspace, DCFP_EINVAL, "channel_gather_grouped: null pointer");
  DCFP_REQUIRE(n > 0, DCFP_EINVAL, "channel_gather_grouped: n=%d", n);
  DCFP_REQUIRE(elt_size == 4 || elt_size == 2, DCFP_EUNSUPPORTED, "channel_gather_grouped: elt_size %d (2 or 4)", elt_size);
  DCFP_REQUIRE(workspace_bytes >= dcfp_channel_gather_workspace(n), DCFP_EINVAL, "channel_gather_grouped: workspace too small");
  DCFP_REQUIRE(reinterpret_cast<uintptr_t>(desc_workspace) % 8 == 0, DCFP_EINVAL, "channel_gather_grouped: workspace misaligned");
  // host-side staging lives on the stack / heap of this call only (re-entrant)
  long long* prefix = new (std::nothrow) long long[n + 1];
  DCFP_REQUIRE(prefix != nullptr, DCFP_EINVAL, "channel_gather_grouped: out of host memory");
  prefix[0] = 0;
  for (int i = 0; i < n; ++i) {
    int rc = validate_gather(descs_host[i], i);
    if (rc) {
      delete[] prefix;
      return rc;
    }
    const long long total = static_cast<long long>(descs_host[i].n_out) * descs_host[i].n_in * descs_host[i].khw;
    prefix[i + 1] = prefix[i] + (total + kGatherTile - 1) / kGatherTile;
  }
  const long long tiles = prefix[n];
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(desc_workspace);
  long long* d_prefix = reinterpret_cast<long long*>(ws);
  dcfp_gather_desc* d_descs = reinterpret_cast<dcfp_gather_desc*>(ws + static_cast<size_t>(n + 1) * sizeof(long long));
  cudaError_t e = cudaMemcpyAsync(d_prefix, prefix, static_cast<size_t>(n + 1) * sizeof(long long), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_descs, descs_host, static_cast<size_t>(n) * sizeof(dcfp_gather_desc), cudaMemcpyHostToDevice, s);
  delete[] prefix;  // pageable source: the copy has been staged when cudaMemcpyAsync returns
  if (e != cudaSuccess) return cuda_fail(e, "channel_gather_grouped: table upload");
  if (tiles == 0) return 0;
  DCFP_REQUIRE(tiles < (1LL << 31), DCFP_ETOOBIG, "channel_gather_grouped: too many tiles");
  if (elt_size == 4) gather_grouped_kernel<uint32_t><<<static_cast<unsigned>(tiles), kGatherThreads, 0, s>>>(d_descs, d_prefix, n);
  else gather_grouped_kernel<uint16_t><<<static_cast<unsigned>(tiles), kGatherThreads, 0, s>>>(d_descs, d_prefix, n);
  return finish_launch("channel_gather_grouped");
}

extern "C" int dcfp_bias_comp(const float* W, int O, int I, int khw, const float* act, float* offset_out, void* stream) {
  DCFP_REQUIRE(W && act && offset_out, DCFP_EINVAL, "bias_comp: null pointer");
  DCFP_REQUIRE(O > 0 && I > 0 && khw > 0, DCFP_EINVAL, "bias_comp: O=%d I=%d khw=%d", O, I, khw);
  DCFP_REQUIRE(static_cast<long long>(I) * khw < (1LL << 31), DCFP_ETOOBIG, "bias_comp: row too long");
  bias_comp_kernel<<<O, kBiasThreads, 0, static_cast<cudaStream_t>(stream)>>>(W, O, I, khw, act, offset_out);
  return finish_launch("bias_comp");
}
