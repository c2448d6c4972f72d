// torch.ops.dcfp.* -- the thin PyTorch registration over the C ABI of include/dcfp_b200.h.
//
// Everything here is plumbing: validate device / dtype / contiguity, unwrap tensors to raw device
// pointers, fetch the CURRENT CUDA stream, call the extern "C" entry point, turn a non-zero return
// code into a RuntimeError carrying dcfp_last_error().  No arithmetic lives in this file and there
// is no CPU implementation registered: calling an op with CPU tensors raises.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include <vector>

#include "dcfp_b200.h"

namespace {

using at::Tensor;
using c10::optional;

void check_rc(int rc, const char* op) {
  TORCH_CHECK(rc == 0, "dcfp::", op, " failed (", rc, "): ", dcfp_last_error());
}

void* cur_stream() { return static_cast<void*>(at::cuda::getCurrentCUDAStream().stream()); }

void require_cuda(const Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), "dcfp: `", name, "` must be a CUDA tensor (there is no CPU fallback)");
}

int label_dtype_of(const Tensor& label) {
  switch (label.scalar_type()) {
    case at::kByte: return DCFP_LABEL_U8;
    case at::kInt: return DCFP_LABEL_I32;
    case at::kLong: return DCFP_LABEL_I64;
    default: TORCH_CHECK(false, "dcfp: label dtype must be uint8 / int32 / int64, got ", label.scalar_type());
  }
}

// Fills one descriptor; keeps nothing alive (the caller holds the tensors).
dcfp_layer_desc make_desc(const Tensor& x, const optional<Tensor>& dy, const optional<Tensor>& scale,
                          const optional<Tensor>& shift, const optional<Tensor>& keys, const Tensor& S1, const Tensor& S2,
                          int64_t K, int64_t affine_mode) {
  require_cuda(x, "x");
  TORCH_CHECK(x.dim() == 4, "dcfp: feature map must be 4-D [N,C,h,w], got ", x.dim(), "-D");
  TORCH_CHECK(x.scalar_type() == at::kFloat || x.scalar_type() == at::kBFloat16, "dcfp: feature map must be fp32 or bf16");
  dcfp_layer_desc d{};
  const bool nchw = x.is_contiguous();
  const bool nhwc = !nchw && x.is_contiguous(at::MemoryFormat::ChannelsLast);
  TORCH_CHECK(nchw || nhwc, "dcfp: feature map must be contiguous (NCHW) or channels_last");
  d.layout = nchw ? DCFP_NCHW : DCFP_NHWC;
  d.dtype = x.scalar_type() == at::kFloat ? DCFP_F32 : DCFP_BF16;
  d.x = x.data_ptr();
  d.N = static_cast<int32_t>(x.size(0));
  d.C = static_cast<int32_t>(x.size(1));
  d.h = static_cast<int32_t>(x.size(2));
  d.w = static_cast<int32_t>(x.size(3));
  d.K = static_cast<int32_t>(K);
  d.affine_mode = static_cast<int32_t>(affine_mode);
  if (dy.has_value()) {
    const Tensor& g = *dy;
    require_cuda(g, "dy");
    TORCH_CHECK(g.sizes() == x.sizes() && g.scalar_type() == x.scalar_type(), "dcfp: dy must match x in shape and dtype");
    TORCH_CHECK(nchw ? g.is_contiguous() : g.is_contiguous(at::MemoryFormat::ChannelsLast), "dcfp: dy must share x's layout");
    d.dy = g.data_ptr();
  }
  auto per_channel = [&](const optional<Tensor>& t, const char* name) -> const float* {
    if (!t.has_value()) return nullptr;
    require_cuda(*t, name);
    TORCH_CHECK(t->scalar_type() == at::kFloat && t->is_contiguous() && t->numel() == x.size(1), "dcfp: `", name,
                "` must be a contiguous fp32 [C] tensor");
    return t->data_ptr<float>();
  };
  d.scale = per_channel(scale, "scale");
  d.shift = per_channel(shift, "shift");
  if (keys.has_value()) {
    const Tensor& k = *keys;
    require_cuda(k, "keys");
    TORCH_CHECK(k.scalar_type() == at::kByte && k.is_contiguous() && k.dim() == 3 && k.size(0) == x.size(0) &&
                    k.size(1) == x.size(2) && k.size(2) == x.size(3),
                "dcfp: keys must be a contiguous uint8 [N,h,w] tensor at the feature map's resolution (dcfp::label_keys)");
    d.keys = k.data_ptr<uint8_t>();
  }
  // S1/S2: fp64 [K, C], either contiguous or a column slice of a shared [K, sum C] arena
  auto arena = [&](const Tensor& t, const char* name) -> double* {
    require_cuda(t, name);
    TORCH_CHECK(t.scalar_type() == at::kDouble && t.dim() == 2 && t.size(0) == K && t.size(1) == x.size(1) &&
                    (t.size(1) == 1 || t.stride(1) == 1) && (K == 1 || t.stride(0) >= t.size(1)),
                "dcfp: `", name, "` must be an fp64 [K, C] tensor with unit column stride");
    return t.data_ptr<double>();
  };
  d.S1 = arena(S1, "S1");
  d.S2 = arena(S2, "S2");
  const int64_t ld1 = K == 1 ? x.size(1) : S1.stride(0), ld2 = K == 1 ? x.size(1) : S2.stride(0);
  TORCH_CHECK(ld1 == ld2, "dcfp: S1 and S2 must share one row stride");
  d.ld = static_cast<int32_t>(ld1);
  return d;
}

// keys[N,h,w] (uint8) from label[N,H0,W0]; cnt[K] += pixels per class (optional)
Tensor label_keys(const Tensor& label, int64_t h, int64_t w, int64_t K, optional<Tensor> cnt) {
  require_cuda(label, "label");
  TORCH_CHECK(label.dim() == 3 && label.is_contiguous(), "dcfp::label_keys: label must be contiguous [N,H0,W0]");
  Tensor keys = at::empty({label.size(0), h, w}, label.options().dtype(at::kByte));
  double* cp = nullptr;
  if (cnt.has_value()) {
    require_cuda(*cnt, "cnt");
    TORCH_CHECK(cnt->scalar_type() == at::kDouble && cnt->is_contiguous() && cnt->numel() == K,
                "dcfp::label_keys: cnt must be a contiguous fp64 [K] tensor");
    cp = cnt->data_ptr<double>();
  }
  c10::cuda::CUDAGuard guard(label.device());
  check_rc(dcfp_label_keys(label.data_ptr(), label_dtype_of(label), static_cast<int>(label.size(0)),
                           static_cast<int>(label.size(1)), static_cast<int>(label.size(2)), static_cast<int>(h),
                           static_cast<int>(w), static_cast<int>(K), keys.data_ptr<uint8_t>(), cp, cur_stream()),
           "label_keys");
  return keys;
}

void class_stats(const Tensor& x, const optional<Tensor>& dy, const optional<Tensor>& scale, const optional<Tensor>& shift,
                 const optional<Tensor>& keys, Tensor S1, Tensor S2, int64_t K, int64_t affine_mode) {
  const dcfp_layer_desc d = make_desc(x, dy, scale, shift, keys, S1, S2, K, affine_mode);
  c10::cuda::CUDAGuard guard(x.device());
  check_rc(dcfp_class_stats(&d, cur_stream()), "class_stats");
}

// keys: one tensor per layer (layers of equal resolution pass the same tensor), or empty when K == 1
void class_stats_grouped(at::TensorList xs, at::TensorList dys, at::TensorList scales, at::TensorList shifts,
                         at::TensorList keys, at::TensorList S1s, at::TensorList S2s, int64_t K, int64_t affine_mode) {
  const size_t n = xs.size();
  TORCH_CHECK(n > 0, "dcfp::class_stats_grouped: empty layer list");
  TORCH_CHECK(S1s.size() == n && S2s.size() == n, "dcfp::class_stats_grouped: S1/S2 lists must match xs");
  TORCH_CHECK(dys.empty() || dys.size() == n, "dcfp::class_stats_grouped: dys must be empty or match xs");
  TORCH_CHECK(scales.size() == shifts.size() && (scales.empty() || scales.size() == n),
              "dcfp::class_stats_grouped: scales/shifts must be empty or match xs");
  TORCH_CHECK(keys.empty() || keys.size() == n, "dcfp::class_stats_grouped: keys must be empty or match xs");
  std::vector<dcfp_layer_desc> descs(n);
  for (size_t i = 0; i < n; ++i) {
    optional<Tensor> dy = dys.empty() ? optional<Tensor>() : optional<Tensor>(dys[i]);
    optional<Tensor> sc = scales.empty() ? optional<Tensor>() : optional<Tensor>(scales[i]);
    optional<Tensor> sf = shifts.empty() ? optional<Tensor>() : optional<Tensor>(shifts[i]);
    optional<Tensor> ky = keys.empty() ? optional<Tensor>() : optional<Tensor>(keys[i]);
    descs[i] = make_desc(xs[i], dy, sc, sf, ky, S1s[i], S2s[i], K, affine_mode);
  }
  c10::cuda::CUDAGuard guard(xs[0].device());
  for (size_t first = 0; first < n; first += DCFP_MAX_GROUP_LAYERS) {
    const int m = static_cast<int>(std::min<size_t>(DCFP_MAX_GROUP_LAYERS, n - first));
    check_rc(dcfp_class_stats_grouped(descs.data() + first, m, cur_stream()), "class_stats_grouped");
  }
}

Tensor reduce_classes(const Tensor& S1) {
  require_cuda(S1, "S1");
  TORCH_CHECK(S1.dim() == 2 && S1.scalar_type() == at::kDouble && S1.is_contiguous(), "dcfp::reduce_classes: S1 must be fp64 [K,C]");
  Tensor out = at::empty({S1.size(1)}, S1.options().dtype(at::kFloat));
  c10::cuda::CUDAGuard guard(S1.device());
  check_rc(dcfp_reduce_classes(S1.data_ptr<double>(), static_cast<int>(S1.size(0)), static_cast<int>(S1.size(1)),
                               out.data_ptr<float>(), cur_stream()),
           "reduce_classes");
  return out;
}

// step / total: fp64 [2, K, C]; step32: optional fp32 [2, K, C]; returns dgamma fp32 [C]
Tensor fold_step(Tensor step, const optional<Tensor>& total, const optional<Tensor>& step32) {
  require_cuda(step, "step");
  TORCH_CHECK(step.scalar_type() == at::kDouble && step.is_contiguous() && step.dim() == 3 && step.size(0) == 2,
              "dcfp::fold_step: step must be a contiguous fp64 [2,K,C] arena");
  double* tp = nullptr;
  if (total.has_value()) {
    require_cuda(*total, "total");
    TORCH_CHECK(total->scalar_type() == at::kDouble && total->is_contiguous() && total->sizes() == step.sizes(),
                "dcfp::fold_step: total must match step");
    tp = total->data_ptr<double>();
  }
  float* sp = nullptr;
  if (step32.has_value()) {
    require_cuda(*step32, "step32");
    TORCH_CHECK(step32->scalar_type() == at::kFloat && step32->is_contiguous() && step32->sizes() == step.sizes(),
                "dcfp::fold_step: step32 must be a contiguous fp32 tensor of step's shape");
    sp = step32->data_ptr<float>();
  }
  Tensor out = at::empty({step.size(2)}, step.options().dtype(at::kFloat));
  c10::cuda::CUDAGuard guard(step.device());
  check_rc(dcfp_fold_step2(step.data_ptr<double>(), sp, tp, static_cast<int>(step.size(1)), static_cast<int>(step.size(2)),
                           out.data_ptr<float>(), cur_stream()),
           "fold_step");
  return out;
}

// pointer table for the one-launch EIC update: built on the host, shipped with one async copy
void eic_update(at::TensorList grads, at::TensorList gammas, const Tensor& offsets, Tensor eic, double r, double one_minus_r,
                bool first_step) {
  const size_t n = grads.size();
  TORCH_CHECK(n > 0 && gammas.size() == n, "dcfp::eic_update: grads/gammas must be equally long, non-empty lists");
  require_cuda(eic, "eic");
  TORCH_CHECK(eic.scalar_type() == at::kFloat && eic.is_contiguous(), "dcfp::eic_update: eic must be contiguous fp32");
  require_cuda(offsets, "offsets");
  TORCH_CHECK(offsets.scalar_type() == at::kInt && offsets.is_contiguous() && offsets.numel() == static_cast<int64_t>(n) + 1,
              "dcfp::eic_update: offsets must be int32 [n_layers+1]");
  Tensor table = at::empty({static_cast<int64_t>(2 * n)}, at::TensorOptions().dtype(at::kLong).pinned_memory(true));
  int64_t* t = table.data_ptr<int64_t>();
  for (size_t i = 0; i < n; ++i) {
    require_cuda(grads[i], "grad");
    require_cuda(gammas[i], "gamma");
    TORCH_CHECK(grads[i].scalar_type() == at::kFloat && gammas[i].scalar_type() == at::kFloat && grads[i].is_contiguous() &&
                    gammas[i].is_contiguous() && grads[i].numel() == gammas[i].numel(),
                "dcfp::eic_update: layer ", i, ": grad/gamma must be contiguous fp32 of equal length");
    t[i] = reinterpret_cast<int64_t>(grads[i].data_ptr<float>());
    t[n + i] = reinterpret_cast<int64_t>(gammas[i].data_ptr<float>());
  }
  c10::cuda::CUDAGuard guard(eic.device());
  Tensor dev = table.to(eic.device(), /*non_blocking=*/true);
  const float* const* gp = reinterpret_cast<const float* const*>(dev.data_ptr<int64_t>());
  check_rc(dcfp_eic_update(gp, gp + n, offsets.data_ptr<int32_t>(), static_cast<int>(n), eic.data_ptr<float>(),
                           static_cast<float>(r), static_cast<float>(one_minus_r), first_step ? 1 : 0, cur_stream()),
           "eic_update");
}

void eic_update_flat(const Tensor& grad, const Tensor& gamma, Tensor eic, double r, double one_minus_r, bool first_step) {
  require_cuda(grad, "grad");
  require_cuda(gamma, "gamma");
  require_cuda(eic, "eic");
  TORCH_CHECK(grad.scalar_type() == at::kFloat && gamma.scalar_type() == at::kFloat && eic.scalar_type() == at::kFloat,
              "dcfp::eic_update_flat: fp32 tensors required");
  TORCH_CHECK(grad.is_contiguous() && gamma.is_contiguous() && eic.is_contiguous() && grad.numel() == eic.numel() &&
                  gamma.numel() == eic.numel(),
              "dcfp::eic_update_flat: contiguous tensors of equal length required");
  c10::cuda::CUDAGuard guard(eic.device());
  check_rc(dcfp_eic_update_flat(grad.data_ptr<float>(), gamma.data_ptr<float>(), eic.data_ptr<float>(),
                                static_cast<int>(eic.numel()), static_cast<float>(r), static_cast<float>(one_minus_r),
                                first_step ? 1 : 0, cur_stream()),
           "eic_update_flat");
}

std::tuple<Tensor, Tensor, Tensor> thresh_mask(const Tensor& score, const Tensor& layer_off, const Tensor& layer_group,
                                               const Tensor& min_keep, int64_t k0, int64_t k1) {
  require_cuda(score, "score");
  require_cuda(layer_off, "layer_off");
  require_cuda(layer_group, "layer_group");
  require_cuda(min_keep, "min_keep");
  TORCH_CHECK(score.scalar_type() == at::kFloat && score.is_contiguous() && score.dim() == 1, "dcfp::thresh_mask: score must be fp32 [n]");
  const int64_t n_layers = layer_group.numel();
  TORCH_CHECK(layer_off.scalar_type() == at::kInt && layer_group.scalar_type() == at::kInt && min_keep.scalar_type() == at::kInt,
              "dcfp::thresh_mask: layer tables must be int32");
  TORCH_CHECK(layer_off.numel() == n_layers + 1 && min_keep.numel() == n_layers, "dcfp::thresh_mask: layer table sizes disagree");
  Tensor mask = at::empty_like(score);
  Tensor thresh = at::empty({2}, score.options());
  Tensor kept = at::empty({n_layers}, layer_off.options());
  const int64_t k_idx[2] = {k0, k1};
  c10::cuda::CUDAGuard guard(score.device());
  check_rc(dcfp_thresh_mask(score.data_ptr<float>(), layer_off.data_ptr<int32_t>(), layer_group.data_ptr<int32_t>(),
                            min_keep.data_ptr<int32_t>(), static_cast<int>(n_layers), static_cast<int>(score.numel()), k_idx,
                            mask.data_ptr<float>(), thresh.data_ptr<float>(), kept.data_ptr<int32_t>(), cur_stream()),
           "thresh_mask");
  return {mask, thresh, kept};
}

struct GatherShape {
  int64_t O, I, khw;
};
GatherShape gather_shape(const Tensor& src) {
  TORCH_CHECK(src.dim() >= 1, "dcfp::channel_gather: scalar tensor");
  GatherShape g{src.size(0), src.dim() >= 2 ? src.size(1) : 1, 1};
  for (int64_t d = 2; d < src.dim(); ++d) g.khw *= src.size(d);
  return g;
}
const int32_t* idx_ptr(const optional<Tensor>& idx, const char* name) {
  if (!idx.has_value()) return nullptr;
  require_cuda(*idx, name);
  TORCH_CHECK(idx->scalar_type() == at::kInt && idx->is_contiguous() && idx->dim() == 1, "dcfp::channel_gather: `", name,
              "` must be int32 [n]");
  return idx->data_ptr<int32_t>();
}
std::vector<int64_t> gathered_sizes(const Tensor& src, const optional<Tensor>& out_idx, const optional<Tensor>& in_idx) {
  std::vector<int64_t> sizes = src.sizes().vec();
  if (out_idx.has_value()) sizes[0] = out_idx->numel();
  if (in_idx.has_value()) {
    TORCH_CHECK(src.dim() >= 2, "dcfp::channel_gather: in_idx given for a 1-D tensor");
    sizes[1] = in_idx->numel();
  }
  return sizes;
}

Tensor channel_gather(const Tensor& src, const optional<Tensor>& out_idx, const optional<Tensor>& in_idx) {
  require_cuda(src, "src");
  TORCH_CHECK(src.is_contiguous(), "dcfp::channel_gather: src must be contiguous");
  const int es = static_cast<int>(src.element_size());
  const GatherShape g = gather_shape(src);
  Tensor dst = at::empty(gathered_sizes(src, out_idx, in_idx), src.options());
  const int n_out = static_cast<int>(dst.size(0));
  const int n_in = src.dim() >= 2 ? static_cast<int>(dst.size(1)) : 1;
  if (dst.numel() == 0) return dst;  // empty selection: nothing to launch
  c10::cuda::CUDAGuard guard(src.device());
  check_rc(dcfp_channel_gather(src.data_ptr(), dst.data_ptr(), idx_ptr(out_idx, "out_idx"), n_out, idx_ptr(in_idx, "in_idx"), n_in,
                               static_cast<int>(g.I), static_cast<int>(g.khw), es, cur_stream()),
           "channel_gather");
  return dst;
}

// an index tensor with numel()==0 and dim()==0 ... cannot express "keep all"; use a parallel bool list instead
std::vector<Tensor> channel_gather_grouped(at::TensorList srcs, at::TensorList out_idx, at::TensorList in_idx,
                                           at::IntArrayRef has_out, at::IntArrayRef has_in) {
  const size_t n = srcs.size();
  TORCH_CHECK(n > 0 && out_idx.size() == n && in_idx.size() == n && has_out.size() == n && has_in.size() == n,
              "dcfp::channel_gather_grouped: list lengths disagree");
  std::vector<Tensor> dsts(n);
  std::vector<dcfp_gather_desc> descs(n);
  const int es = static_cast<int>(srcs[0].element_size());
  for (size_t i = 0; i < n; ++i) {
    const Tensor& src = srcs[i];
    require_cuda(src, "src");
    TORCH_CHECK(src.is_contiguous() && static_cast<int>(src.element_size()) == es,
                "dcfp::channel_gather_grouped: sources must be contiguous and share one element size");
    optional<Tensor> oi = has_out[i] ? optional<Tensor>(out_idx[i]) : optional<Tensor>();
    optional<Tensor> ii = has_in[i] ? optional<Tensor>(in_idx[i]) : optional<Tensor>();
    const GatherShape g = gather_shape(src);
    dsts[i] = at::empty(gathered_sizes(src, oi, ii), src.options());
    descs[i] = dcfp_gather_desc{src.data_ptr(),
                                dsts[i].data_ptr(),
                                idx_ptr(oi, "out_idx"),
                                idx_ptr(ii, "in_idx"),
                                static_cast<int32_t>(dsts[i].size(0)),
                                static_cast<int32_t>(src.dim() >= 2 ? dsts[i].size(1) : 1),
                                static_cast<int32_t>(g.I),
                                static_cast<int32_t>(g.khw)};
  }
  c10::cuda::CUDAGuard guard(srcs[0].device());
  const size_t ws_bytes = dcfp_channel_gather_workspace(static_cast<int>(n));
  Tensor ws = at::empty({static_cast<int64_t>((ws_bytes + 7) / 8)}, srcs[0].options().dtype(at::kLong));
  check_rc(dcfp_channel_gather_grouped(descs.data(), static_cast<int>(n), es, ws.data_ptr(), ws_bytes, cur_stream()),
           "channel_gather_grouped");
  return dsts;
}

Tensor bias_comp(const Tensor& W, const Tensor& act) {
  require_cuda(W, "W");
  require_cuda(act, "act");
  TORCH_CHECK(W.scalar_type() == at::kFloat && W.is_contiguous() && W.dim() >= 2, "dcfp::bias_comp: W must be contiguous fp32 [O,I,...]");
  TORCH_CHECK(act.scalar_type() == at::kFloat && act.is_contiguous() && act.numel() == W.size(1),
              "dcfp::bias_comp: act must be contiguous fp32 [I]");
  const GatherShape g = gather_shape(W);
  Tensor out = at::empty({W.size(0)}, W.options());
  c10::cuda::CUDAGuard guard(W.device());
  check_rc(dcfp_bias_comp(W.data_ptr<float>(), static_cast<int>(g.O), static_cast<int>(g.I), static_cast<int>(g.khw),
                          act.data_ptr<float>(), out.data_ptr<float>(), cur_stream()),
           "bias_comp");
  return out;
}

// -> (weight fp64 [N,H,W], class_num int64 [N,K]); datasets/Base.py:73-89
std::tuple<Tensor, Tensor> class_balance_weights(const Tensor& label, int64_t K, const optional<Tensor>& sample_class, int64_t mode,
                                                 double beta, int64_t ignore_label) {
  require_cuda(label, "label");
  TORCH_CHECK(label.dim() == 3 && label.is_contiguous(), "dcfp::class_balance_weights: label must be contiguous [N,H,W]");
  const int32_t* cls = nullptr;
  if (sample_class.has_value()) {
    require_cuda(*sample_class, "sample_class");
    TORCH_CHECK(sample_class->scalar_type() == at::kInt && sample_class->is_contiguous() && sample_class->numel() == label.size(0),
                "dcfp::class_balance_weights: sample_class must be int32 [N]");
    cls = sample_class->data_ptr<int32_t>();
  }
  Tensor counts = at::empty({label.size(0), K + 1}, label.options().dtype(at::kLong));
  Tensor weight = at::empty(label.sizes(), label.options().dtype(at::kDouble));
  c10::cuda::CUDAGuard guard(label.device());
  check_rc(dcfp_class_balance_weights(label.data_ptr(), label_dtype_of(label), static_cast<int>(label.size(0)),
                                      static_cast<int>(label.size(1)), static_cast<int>(label.size(2)), static_cast<int>(K),
                                      static_cast<int>(ignore_label), cls, static_cast<int>(mode), beta,
                                      counts.data_ptr<int64_t>(), weight.data_ptr<double>(), cur_stream()),
           "class_balance_weights");
  return {weight, counts.slice(1, 0, K)};
}

// ---- f1: fused training-mode BatchNorm2d (+ReLU) ------------------------------------------------------------------
const float* chan_f32(const Tensor& t, int64_t C, const char* name) {
  require_cuda(t, name);
  TORCH_CHECK(t.scalar_type() == at::kFloat && t.is_contiguous() && t.numel() == C, "dcfp::bn: `", name, "` must be a contiguous fp32 [C] tensor");
  return t.data_ptr<float>();
}
dcfp_bn_desc bn_desc(const Tensor& x, const Tensor& gamma, const Tensor& beta, const Tensor& mean, const Tensor& invstd,
                     const Tensor& sums, bool relu) {
  require_cuda(x, "x");
  TORCH_CHECK(x.dim() == 4 && x.is_contiguous(at::MemoryFormat::ChannelsLast), "dcfp::bn: x must be a channels_last [N,C,h,w] tensor");
  TORCH_CHECK(x.scalar_type() == at::kFloat || x.scalar_type() == at::kBFloat16, "dcfp::bn: x must be fp32 or bf16");
  const int64_t C = x.size(1);
  dcfp_bn_desc d{};
  d.x = x.data_ptr();
  d.gamma = chan_f32(gamma, C, "gamma");
  d.beta = chan_f32(beta, C, "beta");
  d.mean = const_cast<float*>(chan_f32(mean, C, "mean"));
  d.invstd = const_cast<float*>(chan_f32(invstd, C, "invstd"));
  require_cuda(sums, "scratch");
  TORCH_CHECK(sums.scalar_type() == at::kDouble && sums.is_contiguous() &&
                  static_cast<size_t>(sums.numel()) * sizeof(double) >= dcfp_bn_scratch_bytes(static_cast<int>(C)),
              "dcfp::bn: scratch must be a contiguous, zeroed fp64 tensor of at least bn_scratch_bytes(C) bytes");
  d.scratch = sums.data_ptr<double>();
  d.N = static_cast<int32_t>(x.size(0));
  d.C = static_cast<int32_t>(C);
  d.h = static_cast<int32_t>(x.size(2));
  d.w = static_cast<int32_t>(x.size(3));
  d.dtype = x.scalar_type() == at::kFloat ? DCFP_F32 : DCFP_BF16;
  d.relu = relu ? 1 : 0;
  return d;
}

int64_t bn_workspace_bytes(int64_t C) { return static_cast<int64_t>(dcfp_bn_workspace_bytes(static_cast<int>(C))); }
int64_t bn_scratch_bytes(int64_t C) { return static_cast<int64_t>(dcfp_bn_scratch_bytes(static_cast<int>(C))); }

bool bn_supported(int64_t N, int64_t C, int64_t h, int64_t w, bool bf16) {
  return dcfp_bn_supported(static_cast<int>(N), static_cast<int>(C), static_cast<int>(h), static_cast<int>(w), bf16 ? DCFP_BF16 : DCFP_F32) != 0;
}

// -> (y, mean, invstd); sums: the zeroed scratch (fp64 tensor of >= bn_scratch_bytes(C) bytes)
std::tuple<Tensor, Tensor, Tensor> bn_forward(const Tensor& x, const Tensor& gamma, const Tensor& beta,
                                              const optional<Tensor>& running_mean, const optional<Tensor>& running_var,
                                              Tensor sums, double momentum, double eps, bool relu, int64_t phases,
                                              const optional<Tensor>& workspace, const optional<Tensor>& residual) {
  Tensor y = phases == 1 ? at::empty({0}, x.options()) : at::empty_like(x, x.options(), at::MemoryFormat::ChannelsLast);
  Tensor mean = at::empty({x.size(1)}, x.options().dtype(at::kFloat));
  Tensor invstd = at::empty({x.size(1)}, x.options().dtype(at::kFloat));
  dcfp_bn_desc d = bn_desc(x, gamma, beta, mean, invstd, sums, relu);
  d.phases = static_cast<int32_t>(phases);
  d.y = phases == 1 ? const_cast<void*>(d.x) : y.data_ptr();  // phase 1 writes no y (validation wants a non-null pointer)
  TORCH_CHECK(running_mean.has_value() == running_var.has_value(), "dcfp::bn_forward: running_mean / running_var must come together");
  if (running_mean.has_value()) {
    d.running_mean = const_cast<float*>(chan_f32(*running_mean, x.size(1), "running_mean"));
    d.running_var = const_cast<float*>(chan_f32(*running_var, x.size(1), "running_var"));
  }
  d.momentum = static_cast<float>(momentum);
  d.eps = static_cast<float>(eps);
  if (workspace.has_value()) {
    require_cuda(*workspace, "workspace");
    TORCH_CHECK(workspace->is_contiguous(), "dcfp::bn_forward: workspace must be contiguous");
    d.workspace = workspace->data_ptr();
    d.workspace_bytes = static_cast<int64_t>(workspace->numel() * workspace->element_size());
  }
  if (residual.has_value()) {
    require_cuda(*residual, "residual");
    TORCH_CHECK(residual->sizes() == x.sizes() && residual->scalar_type() == x.scalar_type() &&
                    residual->is_contiguous(at::MemoryFormat::ChannelsLast),
                "dcfp::bn_forward: residual must match x (shape, dtype, channels_last)");
    d.residual = residual->data_ptr();
  }
  c10::cuda::CUDAGuard guard(x.device());
  check_rc(dcfp_bn_forward(&d, cur_stream()), "bn_forward");
  return {y, mean, invstd};
}

// -> (dx or an empty tensor, dgamma, dbeta); S1/S2: fp64 [K,C] class rows (+=); sums: the zeroed scratch
std::tuple<Tensor, Tensor, Tensor> bn_backward(const Tensor& x, const Tensor& dy, const Tensor& gamma, const Tensor& beta,
                                               const Tensor& mean, const Tensor& invstd, const optional<Tensor>& keys, Tensor S1,
                                               Tensor S2, int64_t K, Tensor sums, bool relu, bool need_dx, int64_t phases) {
  dcfp_bn_desc d = bn_desc(x, gamma, beta, mean, invstd, sums, relu);
  d.phases = static_cast<int32_t>(phases);
  if (phases == 1) need_dx = false;  // reduction pass only: no outputs besides sums / S1 / S2
  require_cuda(dy, "dy");
  TORCH_CHECK(dy.sizes() == x.sizes() && dy.scalar_type() == x.scalar_type() && dy.is_contiguous(at::MemoryFormat::ChannelsLast),
              "dcfp::bn_backward: dy must match x (shape, dtype, channels_last)");
  d.dy = dy.data_ptr();
  if (S1.scalar_type() == at::kFloat) {  // fp32 per-step arena: vector reductions (dcfp_bn_desc.arena_f32)
    require_cuda(S1, "S1");
    require_cuda(S2, "S2");
    TORCH_CHECK(S2.scalar_type() == at::kFloat && S1.dim() == 2 && S2.dim() == 2 && S1.size(0) == K && S2.size(0) == K &&
                    S1.size(1) == x.size(1) && S2.size(1) == x.size(1) && S1.stride(1) == 1 && S2.stride(1) == 1 &&
                    S1.stride(0) == S2.stride(0),
                "dcfp::bn_backward: fp32 S1 / S2 must be [K, C] views with unit column stride and one row stride");
    if (keys.has_value()) {
      require_cuda(*keys, "keys");
      TORCH_CHECK(keys->scalar_type() == at::kByte && keys->is_contiguous() && keys->dim() == 3 && keys->size(0) == x.size(0) &&
                      keys->size(1) == x.size(2) && keys->size(2) == x.size(3), "dcfp::bn_backward: keys must be uint8 [N,h,w]");
      d.keys = keys->data_ptr<uint8_t>();
    }
    d.S1 = S1.data_ptr<float>();
    d.S2 = S2.data_ptr<float>();
    d.ld = static_cast<int32_t>(K == 1 ? x.size(1) : S1.stride(0));
    d.arena_f32 = 1;
  } else {
    // reuse the K1 descriptor checks for keys / S1 / S2
    const dcfp_layer_desc L = make_desc(x, dy, invstd, mean, keys, S1, S2, K, DCFP_AFFINE_INVSTD_MEAN);
    d.keys = L.keys;
    d.S1 = L.S1;
    d.S2 = L.S2;
    d.ld = L.ld;
  }
  d.K = static_cast<int32_t>(K);
  Tensor dx = need_dx ? at::empty_like(x, x.options(), at::MemoryFormat::ChannelsLast) : at::empty({0}, x.options());
  Tensor dgamma = at::empty({x.size(1)}, x.options().dtype(at::kFloat));
  Tensor dbeta = at::empty({x.size(1)}, x.options().dtype(at::kFloat));
  d.dx = need_dx ? dx.data_ptr() : nullptr;
  d.dgamma = dgamma.data_ptr<float>();
  d.dbeta = dbeta.data_ptr<float>();
  c10::cuda::CUDAGuard guard(x.device());
  check_rc(dcfp_bn_backward(&d, cur_stream()), "bn_backward");
  return {dx, dgamma, dbeta};
}

// dz = (y > 0) ? dy (+ dy2) : 0; the tensors share shape, dtype and (dense) layout
Tensor relu_grad(const Tensor& y, const Tensor& dy, const optional<Tensor>& dy2) {
  require_cuda(y, "y");
  require_cuda(dy, "dy");
  TORCH_CHECK(y.scalar_type() == at::kFloat || y.scalar_type() == at::kBFloat16, "dcfp::relu_grad: fp32 or bf16 tensors");
  TORCH_CHECK(y.is_non_overlapping_and_dense() && dy.sizes() == y.sizes() && dy.strides() == y.strides() && dy.scalar_type() == y.scalar_type(),
              "dcfp::relu_grad: dy must match y (shape, strides, dtype), dense");
  if (dy2.has_value()) {
    require_cuda(*dy2, "dy2");
    TORCH_CHECK(dy2->sizes() == y.sizes() && dy2->strides() == y.strides() && dy2->scalar_type() == y.scalar_type(),
                "dcfp::relu_grad: dy2 must match y (shape, strides, dtype)");
  }
  Tensor dz = at::empty_like(y);  // preserves the (dense) strides
  TORCH_CHECK(dz.strides() == y.strides(), "dcfp::relu_grad: could not allocate an output with the input's layout");
  c10::cuda::CUDAGuard guard(y.device());
  check_rc(dcfp_relu_grad(y.data_ptr(), dy.data_ptr(), dy2.has_value() ? dy2->data_ptr() : nullptr, dz.data_ptr(), y.numel(),
                          y.scalar_type() == at::kFloat ? DCFP_F32 : DCFP_BF16, cur_stream()),
           "relu_grad");
  return dz;
}

int64_t launch_count(bool reset) { return dcfp_launch_count(reset ? 1 : 0); }
int64_t abi_version() { return dcfp_abi_version(); }

}  // namespace

TORCH_LIBRARY(dcfp, m) {
  m.def("label_keys(Tensor label, int h, int w, int K, Tensor(a!)? cnt) -> Tensor", &label_keys);
  m.def("class_stats(Tensor x, Tensor? dy, Tensor? scale, Tensor? shift, Tensor? keys, Tensor(a!) S1, Tensor(b!) S2, int K, int affine_mode=0) -> ()",
        &class_stats);
  m.def(
      "class_stats_grouped(Tensor[] xs, Tensor[] dys, Tensor[] scales, Tensor[] shifts, Tensor[] keys, Tensor(a!)[] S1s, "
      "Tensor(b!)[] S2s, int K, int affine_mode=0) -> ()",
      &class_stats_grouped);
  m.def("reduce_classes(Tensor S1) -> Tensor", &reduce_classes);
  m.def("fold_step(Tensor(a!) step, Tensor(b!)? total, Tensor(c!)? step32=None) -> Tensor", &fold_step);
  m.def("eic_update(Tensor[] grads, Tensor[] gammas, Tensor offsets, Tensor(a!) eic, float r, float one_minus_r, bool first_step) -> ()",
        &eic_update);
  m.def("eic_update_flat(Tensor grad, Tensor gamma, Tensor(a!) eic, float r, float one_minus_r, bool first_step) -> ()",
        &eic_update_flat);
  m.def("thresh_mask(Tensor score, Tensor layer_off, Tensor layer_group, Tensor min_keep, int k0, int k1) -> (Tensor, Tensor, Tensor)",
        &thresh_mask);
  m.def("channel_gather(Tensor src, Tensor? out_idx, Tensor? in_idx) -> Tensor", &channel_gather);
  m.def("channel_gather_grouped(Tensor[] srcs, Tensor[] out_idx, Tensor[] in_idx, int[] has_out, int[] has_in) -> Tensor[]",
        &channel_gather_grouped);
  m.def("bias_comp(Tensor W, Tensor act) -> Tensor", &bias_comp);
  m.def("class_balance_weights(Tensor label, int K, Tensor? sample_class, int mode, float beta, int ignore_label) -> (Tensor, Tensor)",
        &class_balance_weights);
  m.def("bn_supported(int N, int C, int h, int w, bool bf16) -> bool", &bn_supported);
  m.def("bn_scratch_bytes(int C) -> int", &bn_scratch_bytes);
  m.def("bn_workspace_bytes(int C) -> int", &bn_workspace_bytes);
  m.def("bn_forward(Tensor x, Tensor gamma, Tensor beta, Tensor(a!)? running_mean, Tensor(b!)? running_var, Tensor(c!) sums, float momentum, "
        "float eps, bool relu, int phases=0, Tensor? workspace=None, Tensor? residual=None) -> (Tensor, Tensor, Tensor)",
        &bn_forward);
  m.def("bn_backward(Tensor x, Tensor dy, Tensor gamma, Tensor beta, Tensor mean, Tensor invstd, Tensor? keys, Tensor(a!) S1, "
        "Tensor(b!) S2, int K, Tensor(c!) sums, bool relu, bool need_dx, int phases=0) -> (Tensor, Tensor, Tensor)",
        &bn_backward);
  m.def("relu_grad(Tensor y, Tensor dy, Tensor? dy2=None) -> Tensor", &relu_grad);
  m.def("launch_count(bool reset) -> int", &launch_count);
  m.def("abi_version() -> int", &abi_version);
}
