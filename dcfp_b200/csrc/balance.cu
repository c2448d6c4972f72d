// Class-balance pixel weights -- datasets/Base.py:73-89 of the reference (`get_label`,
// used by the finetune stage: BALANCE=2, LOSS_TYPE='gsrl').  A second label-keyed reduction on the label path K1 uses:
//
//   class_num[n][k] = #pixels of image n with label k                (np.bincount, ignore label -> extra bin, dropped)
//   mode 1:  w[k] = 1 / (class_num[k] + 1)
//   mode 2:  w[k] = (1 + 1e-8 - beta^class_num[cls_n]) / (1 + 1e-8 - beta^class_num[k])     (effective-number weights)
//   w = clip(w, 0, 1);  weight[n][p] = label == ignore ? 0 : w[label[n][p]]                    (float64, as numpy)
//
// The reference does this per image on the host inside the data loader; here one launch histograms the whole batch
// (shared-memory bins, one row of the count table per image) and one launch maps the pixels.  HBM-bound and tiny:
// reads the label twice, writes 8 B per pixel.
#include <algorithm>

#include "common.cuh"

namespace dcfp {
namespace {

__device__ __forceinline__ int label_at(const void* label, int dtype, long long idx) {
  if (dtype == DCFP_LABEL_U8) return static_cast<const unsigned char*>(label)[idx];
  if (dtype == DCFP_LABEL_I32) return static_cast<const int*>(label)[idx];
  const long long v = static_cast<const long long*>(label)[idx];
  return (v < 0 || v > 0x7fffffffLL) ? -1 : static_cast<int>(v);
}

// grid = (chunks, N); counts[n][k] (K + 1 bins: the last one collects the ignore label)
__global__ void __launch_bounds__(256) image_hist_kernel(const void* __restrict__ label, int dtype, long long hw, int K,
                                                         int ignore_label, long long* __restrict__ counts) {
  __shared__ unsigned hist[257];
  for (int i = threadIdx.x; i < 257; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  const int n = blockIdx.y;
  const long long base = static_cast<long long>(n) * hw;
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < hw;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int l = label_at(label, dtype, base + p);
    const int bin = (l == ignore_label) ? K : l;
    if (bin >= 0 && bin <= K) atomicAdd(&hist[bin], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= K; i += blockDim.x)
    if (hist[i]) atomicAdd(reinterpret_cast<unsigned long long*>(&counts[static_cast<long long>(n) * (K + 1) + i]),
                           static_cast<unsigned long long>(hist[i]));
}

__global__ void __launch_bounds__(256) balance_weight_kernel(const void* __restrict__ label, int dtype, long long hw, int K,
                                                             int ignore_label, const long long* __restrict__ counts,
                                                             const int* __restrict__ sample_class, int mode, double beta,
                                                             double* __restrict__ weight) {
  __shared__ double table[257];
  const int n = blockIdx.y;
  const long long* cnt = counts + static_cast<long long>(n) * (K + 1);
  double numer = 1.0;
  if (mode == 2) {
    // a class outside [0, K] cannot index the count table: it is treated as a class without pixels (numer = 1e-8); the
    // reference (numpy fancy indexing, datasets/Base.py:84) raises IndexError there -- a kernel cannot, and the host does not
    // read device data back to validate it
    const int cls = sample_class[n];
    const double c_cls = (cls >= 0 && cls <= K) ? static_cast<double>(cnt[cls]) : 0.0;
    numer = 1.0 + 1e-8 - pow(beta, c_cls);
  }
  for (int k = threadIdx.x; k <= K; k += blockDim.x) {
    double w = 0.0;  // the ignore bin (k == K) maps to weight 0 (Base.py:84)
    if (k < K) {
      const double c = static_cast<double>(cnt[k]);
      w = mode == 1 ? 1.0 / (c + 1.0) : numer / (1.0 + 1e-8 - pow(beta, c));
      w = fmin(fmax(w, 0.0), 1.0);
    }
    table[k] = w;
  }
  __syncthreads();
  const long long base = static_cast<long long>(n) * hw;
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < hw;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int l = label_at(label, dtype, base + p);
    const int bin = (l == ignore_label) ? K : l;
    weight[base + p] = (bin >= 0 && bin <= K) ? table[bin] : 0.0;
  }
}

}  // namespace
}  // namespace dcfp

using namespace dcfp;

extern "C" int dcfp_class_balance_weights(const void* label, int label_dtype, int N, int H, int W, int K, int ignore_label,
                                          const int32_t* sample_class, int mode, double beta, int64_t* class_num,
                                          double* weight, void* stream) {
  DCFP_REQUIRE(label && class_num && weight, DCFP_EINVAL, "class_balance_weights: null pointer");
  DCFP_REQUIRE(N > 0 && H > 0 && W > 0, DCFP_EINVAL, "class_balance_weights: bad extent");
  DCFP_REQUIRE(K >= 1 && K <= DCFP_MAX_CLASSES, DCFP_ETOOBIG, "class_balance_weights: K=%d outside [1,%d]", K, DCFP_MAX_CLASSES);
  DCFP_REQUIRE(label_dtype >= DCFP_LABEL_U8 && label_dtype <= DCFP_LABEL_I64, DCFP_EINVAL,
               "class_balance_weights: unknown label dtype %d", label_dtype);
  DCFP_REQUIRE(mode == 1 || mode == 2, DCFP_EINVAL, "class_balance_weights: mode %d (1 = inverse count, 2 = effective number)", mode);
  DCFP_REQUIRE(mode == 1 || sample_class != nullptr, DCFP_EINVAL, "class_balance_weights: mode 2 needs sample_class[N]");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long hw = static_cast<long long>(H) * W;
  cudaError_t e = cudaMemsetAsync(class_num, 0, static_cast<size_t>(N) * (K + 1) * sizeof(int64_t), s);
  if (e != cudaSuccess) return cuda_fail(e, "class_balance_weights: memset");
  const int chunks = static_cast<int>(std::min<long long>((hw + 4095) / 4096, 2LL * num_sms()));
  dim3 grid(chunks, N);
  image_hist_kernel<<<grid, 256, 0, s>>>(label, label_dtype, hw, K, ignore_label, reinterpret_cast<long long*>(class_num));
  int rc = finish_launch("image_hist");
  if (rc) return rc;
  balance_weight_kernel<<<grid, 256, 0, s>>>(label, label_dtype, hw, K, ignore_label, reinterpret_cast<const long long*>(class_num),
                                             sample_class, mode, beta, weight);
  return finish_launch("balance_weight");
}
