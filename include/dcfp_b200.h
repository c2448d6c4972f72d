/*
 * dcfp_b200.h -- C ABI of the B200-native DCFP scoring -> keep-mask -> channel-gather path.
 *
 * This is the drop-in boundary.  The reference (wzx99/DCFP) is pure
 * Python: its "FFI" for this path is the set of torch tensor ops issued by
 * pruners/dcfp_pruner.py and pruners/channel_pruner.py.  Each entry point below names the
 * reference lines it replaces.  The Python binding a maintainer adds on the reference side is
 * shown in INTEGRATION.md (ctypes stub and the TORCH_LIBRARY binding in
 * dcfp_b200/csrc/torch_binding.cpp).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it;
 *   - functions are re-entrant and thread-safe (hooks call them from the autograd engine thread).  They never
 *     allocate device memory and never synchronise the device.  Process-wide state is limited to a launch counter
 *     (dcfp_launch_count: statistics only) and a mutex-guarded record of which kernels already had their
 *     shared-memory limit raised (cudaFuncSetAttribute, once per kernel and device).  Host-side work per call:
 *     argument validation, cuTensorMapEncodeTiled for the TMA kernels (up to 2 maps per layer, host only), and --
 *     dcfp_channel_gather_grouped only -- one cudaMemcpyAsync of the descriptor table from a pageable host staging
 *     array (the runtime copies pageable memory before returning: no device synchronisation, but not asynchronous
 *     to the host).  The cooperative one-launch BN forward (workspace != NULL) needs all its CTAs co-resident and
 *     is launched with cudaLaunchCooperativeKernel;
 *   - return 0 on success, a negative DCFP_E* validation code, or a positive cudaError_t;
 *     no C++ exception crosses the boundary; dcfp_last_error() returns a thread-local message.
 */
#ifndef DCFP_B200_H_
#define DCFP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCFP_ABI_VERSION 4

/* element types of feature maps / weights */
#define DCFP_F32 0
#define DCFP_BF16 1
/* label element types (reference: int64, train.py:253; datasets store uint8) */
#define DCFP_LABEL_U8 0
#define DCFP_LABEL_I32 1
#define DCFP_LABEL_I64 2
/* feature-map layouts */
#define DCFP_NCHW 0
#define DCFP_NHWC 1 /* torch.channels_last */
/* meaning of dcfp_layer_desc.scale / .shift */
#define DCFP_AFFINE_SCALE_SHIFT 0 /* xa = x * scale[c] + shift[c]                                     */
#define DCFP_AFFINE_INVSTD_MEAN 1 /* xa = (x - shift[c]) * scale[c]: scale = invstd, shift = batch mean,
                                     i.e. the two vectors autograd's BN node saved -- no host-side prep */

#define DCFP_EINVAL (-1)      /* bad argument (null pointer, non-positive extent, unknown enum) */
#define DCFP_EUNSUPPORTED (-2) /* valid but not implemented combination */
#define DCFP_ETOOBIG (-3)     /* exceeds a documented limit (K > 255, too many layers, ...) */

/* dcfp_layer_desc.hints */
#define DCFP_HINT_KEEP_L2 1 /* another pass re-reads this map soon: load with normal L2 priority instead of evict-first */

#define DCFP_MAX_CLASSES 255
#define DCFP_MAX_GROUP_LAYERS 160

/* ---- label keys: nearest down-sampling of the label map, once per label resolution --------------
 * keys[n][i][j] = label[n][min(floor(i * (float)H0 / h), H0-1)][min(floor(j * (float)W0 / w), W0-1)]
 * (legacy `nearest` of F.interpolate, evaluated in registers -- no float label tensor is ever
 * materialised); a label outside [0, K) -- e.g. the ignore label 255 -- becomes the key K
 * ("dropped").  cnt[k] += number of pixels of class k at this resolution (fp64; may be NULL).
 * Reference counterpart of the count: np.bincount in datasets/Base.py:78.                       */
int dcfp_label_keys(const void* label, int label_dtype, int N, int H0, int W0, int h, int w, int K, uint8_t* keys,
                    double* cnt, void* stream);

/*
 * One scored feature map (one BN layer of one micro-batch).
 *
 * value functor, per element of channel c at pixel p:
 *   forward  (dy == NULL):  v = xa = x * scale[c] + shift[c]       (scale/shift NULL -> 1 / 0;
 *                                                                   see affine_mode)
 *   backward (dy != NULL):  v = dy * xa     with scale = invstd, shift = -mean * invstd (or
 *                           affine_mode = DCFP_AFFINE_INVSTD_MEAN and shift = mean)
 *                           => v = dy * xhat, and
 *                           sum_k S1[k][c] == d(loss)/d(gamma_c)   (BN backward, the quantity
 *                           pruners/dcfp_pruner.py:18 reads as m.weight.grad)
 * class key:  keys[n][p] from dcfp_label_keys at this layer's (h, w); key >= K is dropped.
 *             keys == NULL puts every pixel in class 0 (K must be 1).
 *             NOTE the identity above sums over ALL pixels: autograd's gradient includes the pixels
 *             whose label is outside [0, K) (ignore label 255 -- no loss, but gradient below the
 *             logits).  dcfp_label_keys gives them the key K; pass K + 1 here (K + 1 rows in S1/S2)
 *             to collect them in a row of their own, as dcfp_b200/scorer.py does.
 * accumulates (+=, fp64):  S1[k*ld + c] += sum v,  S2[k*ld + c] += sum v*v
 */
typedef struct dcfp_layer_desc {
  const void* x;       /* [N,C,h,w] (NCHW) or [N,h,w,C] (NHWC), dtype `dtype` */
  const void* dy;      /* same shape/dtype as x, or NULL */
  const float* scale;  /* [C] or NULL */
  const float* shift;  /* [C] or NULL */
  const uint8_t* keys; /* [N,h,w] class keys, or NULL */
  double* S1;          /* [K, ld] rows; this layer owns columns [0, C) */
  double* S2;          /* [K, ld] */
  int32_t N, C, h, w;
  int32_t K;
  int32_t dtype;  /* DCFP_F32 | DCFP_BF16 */
  int32_t layout; /* DCFP_NCHW | DCFP_NHWC */
  int32_t ld;     /* row stride of S1/S2 in elements (0 -> C): lets all layers share one [K, sum C] arena */
  int32_t affine_mode; /* DCFP_AFFINE_SCALE_SHIFT | DCFP_AFFINE_INVSTD_MEAN */
  int32_t hints;       /* 0, or DCFP_HINT_* bits (was `reserved`, must-be-0, in ABI v2) */
} dcfp_layer_desc;

/* ---- K1: label-keyed segmented reduction over conv/BN feature maps -------------------------
 * north_star "class-conditional calibration statistics"; no reference counterpart for the
 * forward functor; the backward functor restates the dgamma reduction of
 * autograd's BN backward that feeds pruners/dcfp_pruner.py:18.                              */
int dcfp_class_stats(const dcfp_layer_desc* desc_host, void* stream);
/* Same, for up to DCFP_MAX_GROUP_LAYERS resident feature maps per call (all K, dtype, dy-nullness
 * must agree; layouts may mix).  The layers of one layout share ONE launch -- as many as the tensor
 * maps + layer table fit in kernel parameter space (80..160, more are split transparently); tiny or
 * unaligned maps take one small launch each.  Layers only share the grid; outputs stay per layer. */
int dcfp_class_stats_grouped(const dcfp_layer_desc* descs_host, int n_layers, void* stream);

/* ---- K2a: EIC update -- pruners/dcfp_pruner.py:15-20 ------------------------------------------
 *   flag = grad*gamma > 0;  g = flag ? |grad| : eic;  eic = eic*r + g*(1-r)      (fp32, two
 *   roundings then add, exactly the reference's association; first_step treats eic as 0).
 * One launch over all layers: grad_ptrs/gamma_ptrs are device arrays of n_layers device
 * pointers (fp32 [C_l]); offsets is a device array [n_layers+1] into the concatenated eic.   */
int dcfp_eic_update(const float* const* grad_ptrs, const float* const* gamma_ptrs, const int32_t* offsets,
                    int n_layers, float* eic, float r, float one_minus_r, int first_step, void* stream);
/* Same on already-concatenated grad/gamma vectors of n floats. */
int dcfp_eic_update_flat(const float* grad, const float* gamma, float* eic, int n, float r, float one_minus_r,
                         int first_step, void* stream);
/* dgamma[c] = sum_k S1[k*C + c]  (fp64 arena -> fp32), the bridge from K1-backward to K2a; C may be
 * the total channel count of a shared [K, sum C] arena (one launch for all layers).             */
int dcfp_reduce_classes(const double* S1, int K, int C, float* out, void* stream);
/* End-of-step fold of a per-step arena, ONE launch:  dgamma[c] = sum_k step[0][k][c];
 * total[m][k][c] += step[m][k][c] for both moments m in {0,1};  step[m][k][c] = 0.
 * step / total: fp64 [2][K][C] (S1 rows then S2 rows).  total may be NULL (no pass-wide stats). */
int dcfp_fold_step(double* step, double* total, int K, int C, float* dgamma, void* stream);
/* Same with a second, fp32 per-step arena of the same shape (may be NULL): the fused BN backward fills it with vector
 * atomics (dcfp_bn_desc.arena_f32); dgamma and total take step + step32 (summed in fp64), both are zeroed.       */
int dcfp_fold_step2(double* step, float* step32, double* total, int K, int C, float* dgamma, void* stream);

/* ---- K2b: global threshold + keep masks -- pruners/dcfp_pruner.py:43-92 ------------------------
 * score: concatenated fp32 scores of the n_layers scored layers; layer l owns
 * [layer_off[l], layer_off[l+1]) and belongs to group layer_group[l] in {0,1}
 * (get_bn_group, :36-37); the values 2/3 mean "masked with thresh[0/1] but excluded from the
 * threshold set" (a BN in except_layers whose conv is not, :45 vs :73).  thresh[g] = k_idx[g]-th smallest score of group g (ascending, 0-based;
 * get_thresh :59-64; k_idx[g] < 0 -> group empty, thresh 0).  mask = score > thresh (strict,
 * :77); a layer with fewer than min_keep[l] survivors additionally keeps its min_keep[l]
 * highest-scoring channels, ties broken lowest-index-first (:79-82; torch.sort's tie order is
 * implementation-defined there).  All arrays are device arrays except k_idx_host[2].
 * kept_out[l] (optional) receives the number of surviving channels of layer l.                 */
int dcfp_thresh_mask(const float* score, const int32_t* layer_off, const int32_t* layer_group, const int32_t* min_keep,
                     int n_layers, int n_total, const int64_t* k_idx_host, float* mask_out, float* thresh_out,
                     int32_t* kept_out, void* stream);

/* ---- K3: channel gather -- pruners/channel_pruner.py:907-948 (deploy_subnet) --------------------
 * dst[o', i', :] = src[out_idx[o'], in_idx[i'], :] for a weight [O, I, khw] of elt_size-byte
 * elements (khw = kh*kw, 1 for BN vectors/bias with I = 1).  in_idx NULL -> all I inputs kept. */
int dcfp_channel_gather(const void* src, void* dst, const int32_t* out_idx, int n_out, const int32_t* in_idx, int n_in,
                        int I, int khw, int elt_size, void* stream);
typedef struct dcfp_gather_desc {
  const void* src;
  void* dst;
  const int32_t* out_idx;
  const int32_t* in_idx; /* or NULL */
  int32_t n_out, n_in, I, khw;
} dcfp_gather_desc;
/* All weighted modules of a model in ONE launch.  desc_workspace: device scratch of at least
 * dcfp_channel_gather_workspace(n) bytes; the table is copied there with cudaMemcpyAsync. */
size_t dcfp_channel_gather_workspace(int n);
int dcfp_channel_gather_grouped(const dcfp_gather_desc* descs_host, int n, int elt_size, void* desc_workspace,
                                size_t workspace_bytes, void* stream);

/* ---- bias compensation -- pruners/channel_pruner.py:873-905 (resize_subnet_bias) -----------------
 * offset[o] = sum_i act[i] * sum_e W[o, i, e]   (fp32; act = relu((1 - in_mask) * beta_parent)) */
int dcfp_bias_comp(const float* W, int O, int I, int khw, const float* act, float* offset_out, void* stream);

/* ---- class-balance pixel weights -- datasets/Base.py:73-89 (get_label; finetune stage) ------------------
 * class_num[n][k] (k < K; bin K = ignore label) = per-image pixel counts; weight[n][p] (float64, as numpy) =
 * clip(w[label], 0, 1) with w = 1/(class_num+1) (mode 1) or the effective-number ratio
 * (1 + 1e-8 - beta^class_num[sample_class[n]]) / (1 + 1e-8 - beta^class_num[k]) (mode 2); ignored pixels get 0.
 * class_num: device [N][K+1] int64 (overwritten); sample_class: device [N] int32 (mode 2) or NULL; a sample_class outside
 * [0, K] is treated as a class without pixels (the reference raises IndexError there).                           */
int dcfp_class_balance_weights(const void* label, int label_dtype, int N, int H, int W, int K, int ignore_label,
                               const int32_t* sample_class, int mode, double beta, int64_t* class_num, double* weight,
                               void* stream);

/* ---- f1: training-mode BatchNorm2d (+ in-place ReLU) whose backward yields the class-keyed sums ---------------
 * Replaces nn.BatchNorm2d followed by nn.ReLU(inplace=True) as the reference nets use them
 * (networks/backbone/resnet.py:26-56, networks/tools/aspp.py:15-24) and, inside loss.backward()
 * (train.py:265), autograd's ReLU + batch-norm backward whose bn.weight.grad pruners/dcfp_pruner.py:18
 * reads.  channels_last maps only (dcfp_bn_supported tells; anything else stays with torch's BN and the
 * hook path of dcfp_class_stats).
 *   forward : mean/var over the N*h*w pixels (fp64 across CTAs), y = [relu](fma(x, gamma*invstd,
 *             beta - mean*gamma*invstd)); mean, invstd written for the backward; running statistics
 *             updated with `momentum` (the exponential_average_factor; unbiased variance), if given.
 *   backward: dz = dy where the forward output was > 0 (relu) else dy;  S1[k][c] += sum_{p in class k}
 *             dz*xhat, S2 += (dz*xhat)^2 -- the SAME class rows dcfp_class_stats' backward functor fills,
 *             so sum_k S1 == dgamma;  dgamma, dbeta (fp32 [C]);  dx = gamma*invstd*(dz - dbeta/M -
 *             xhat*dgamma/M) unless dx == NULL.  (x, dy) are read once for all sums, once more for dx.
 * `scratch`: caller-provided device buffer of dcfp_bn_scratch_bytes(C) bytes, 16-byte aligned, ZERO on entry
 * (fp64 partial sums striped against same-address atomic contention, which the element-wise pass turns into its
 * per-channel coefficients; the one-launch forward also keeps its coefficient vectors and grid-barrier flags there);
 * one scratch per call in flight.                                                                                  */
typedef struct dcfp_bn_desc {
  const void* x;         /* [N,h,w,C] (channels_last), dtype `dtype` */
  void* y;               /* forward out, same shape */
  const void* dy;        /* backward in */
  void* dx;              /* backward out, or NULL */
  const float* gamma;    /* [C] */
  const float* beta;     /* [C] */
  float* mean;           /* [C] forward: written; backward: read */
  float* invstd;         /* [C] */
  float* running_mean;   /* [C] or NULL (forward) */
  float* running_var;    /* [C] or NULL */
  void* scratch;         /* dcfp_bn_scratch_bytes(C) bytes, zero on entry */
  const uint8_t* keys;   /* backward: [N,h,w] class keys, or NULL (K == 1) */
  void* S1;              /* backward: [K, ld] class rows, fp64 -- or fp32 when arena_f32 is set */
  void* S2;
  float* dgamma;         /* backward out [C] */
  float* dbeta;          /* backward out [C] */
  int32_t N, C, h, w;
  int32_t dtype;         /* DCFP_F32 | DCFP_BF16 */
  int32_t relu;          /* 1: the BN output goes through ReLU (fused) */
  int32_t K, ld;         /* backward: class rows, row stride of S1/S2 (0 -> C) */
  float eps, momentum;
  int32_t phases;        /* 0: the whole call; 1: only the reduction pass (sums, statistics / gradients of gamma and
                            beta [+ S1/S2]); 2: only the element-wise pass (needs the scratch a phase-1 call left) --
                            lets a caller time the two apart */
  int32_t arena_f32;     /* backward: 1 = S1 / S2 are fp32 [K, ld] rows of a PER-STEP arena (zeroed every step by dcfp_fold_step2), 16-byte
                            aligned with ld % 4 == 0: the class rows then leave a CTA as 128-bit vector reductions
                            (red.global.v4.f32), a quarter of the atomic instructions of the fp64 form.  fp32 across the
                            <= #SMs CTAs of ONE launch only; every cross-step / cross-rank sum stays fp64.  0 = fp64 rows */
  void* workspace;       /* forward: dcfp_bn_workspace_bytes(C) bytes of device scratch, NOT zeroed, reusable by every
                            call on the same stream (per-CTA partial sums of the one-launch forward); NULL selects the
                            two-launch forward */
  int64_t workspace_bytes;
  const void* residual;  /* forward, or NULL: [N,h,w,C] like x, added to the normalised value BEFORE the ReLU -- the tail of a
                            bottleneck block, bn3 -> (+ shortcut) -> ReLU (networks/backbone/resnet.py:49-56), in one pass:
                            y = [relu](T(fma(x, scale, shift)) + residual).  The backward's gate is then y > 0: callers pass
                            dz = (y > 0) ? dy : 0 as `dy` with relu = 0, and dz is also the shortcut's gradient */
} dcfp_bn_desc;
int dcfp_bn_supported(int N, int C, int h, int w, int dtype);
size_t dcfp_bn_scratch_bytes(int C);
size_t dcfp_bn_workspace_bytes(int C);
int dcfp_bn_forward(const dcfp_bn_desc* desc_host, void* stream);
int dcfp_bn_backward(const dcfp_bn_desc* desc_host, void* stream);
/* ReLU backward behind a residual sum (the gate of dcfp_bn_desc.residual), with the gradient accumulation autograd would
 * run in front of it folded in:  dz[i] = y[i] > 0 ? dy[i] (+ dy2[i]) : 0  over n elements of `dtype` (any layout: the
 * four tensors share one).  dy2 may be NULL.  Bit-identical to torch's add followed by threshold_backward
 * (networks/backbone/resnet.py:55-56 in reverse: ReLU, then the fan-in of the block input's two gradients).          */
int dcfp_relu_grad(const void* y, const void* dy, const void* dy2, void* dz, int64_t n, int dtype, void* stream);

/* ---- misc ------------------------------------------------------------------------------------- */
const char* dcfp_last_error(void);
int dcfp_abi_version(void);
/* number of kernels launched by this library in the calling thread since the last reset */
int64_t dcfp_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* DCFP_B200_H_ */
