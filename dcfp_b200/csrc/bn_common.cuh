// Fused BN: scratch layout shared by the reduction kernels (bn_stats_kernel, the FUSED instantiation of K1) and the
// element-wise kernels (bn_apply_kernel, bn_dx_kernel), and the per-channel coefficient math between them.
//
// Same-address fp64 atomics serialise in the L2 slice at ~38 cycles each (measured: a per-layer K1 launch whose 1 184
// warps each added their rows to the same 2*C addresses spent 20-60 us in that tail).  So per-channel totals go to one of
// kBnStripes copies (stripe = blockIdx % kBnStripes), after a CTA-level combine.  The element-wise kernel that follows on
// the stream sums the stripes of ITS channels in its prologue (bn_coef_forward / bn_coef_backward: a few dozen L2 loads
// and ~10 fp64 operations per channel, in parallel in every block) -- a "last CTA finalises" step inside the reduction
// kernel was measured at 4-5 us of serial tail per launch (fence + ticket + one CTA doing all channels).
#pragma once
#include "common.cuh"

namespace dcfp {

constexpr int kBnStripes = 4;

constexpr int kBnMaxCtas = 512;  // grid-barrier flags of the cooperative kernels (one per CTA; grids are <= #SMs)

// scratch = double stripes[kBnStripes][2][C] | float coef[5][C] (cooperative forward: scale, shift) | unsigned counter,
// pad[3] | unsigned flags[kBnMaxCtas] (cooperative forward's grid barrier); ZERO on entry
__host__ __device__ inline size_t bn_scratch_bytes(int C) {
  return static_cast<size_t>(kBnStripes) * 2 * C * sizeof(double) + static_cast<size_t>(5) * C * sizeof(float) + 16 +
         kBnMaxCtas * sizeof(unsigned);
}
__host__ __device__ inline double* bn_stripes(void* scratch) { return static_cast<double*>(scratch); }
__host__ __device__ inline float* bn_coef(void* scratch, int C) {
  return reinterpret_cast<float*>(static_cast<char*>(scratch) + static_cast<size_t>(kBnStripes) * 2 * C * sizeof(double));
}
__host__ __device__ inline unsigned* bn_counter(void* scratch, int C) { return reinterpret_cast<unsigned*>(bn_coef(scratch, C) + 5 * C); }
__host__ __device__ inline unsigned* bn_flags(void* scratch, int C) { return bn_counter(scratch, C) + 4; }

struct BnFinal {
  void* scratch;
  const float* gamma;
  const float* beta;
  float* mean;          // forward: out;  backward: in
  float* invstd;
  float* running_mean;  // forward, optional
  float* running_var;
  float* dgamma;        // backward out
  float* dbeta;
  double inv_m;         // 1 / (N*h*w)
  double unbias;        // M / (M - 1)
  float eps, momentum;
  int C;
};

// stripe sums of 4 consecutive channels (c0 % 4 == 0): 128-bit read-only loads through L1 -- the stripes were written by the
// PREVIOUS kernel on the stream, and the blocks of an SM all read the same few KB
__device__ __forceinline__ void bn_sum_stripes4(const double* stripes, int C, int c0, double* s0, double* s1) {
#pragma unroll
  for (int j = 0; j < 4; ++j) s0[j] = s1[j] = 0.0;
#pragma unroll
  for (int s = 0; s < kBnStripes; ++s) {
    const double2* p0 = reinterpret_cast<const double2*>(stripes + static_cast<size_t>(s) * 2 * C + c0);
    const double2* p1 = reinterpret_cast<const double2*>(stripes + static_cast<size_t>(s) * 2 * C + C + c0);
    const double2 a = __ldg(p0), b = __ldg(p0 + 1), c = __ldg(p1), d = __ldg(p1 + 1);
    s0[0] += a.x, s0[1] += a.y, s0[2] += b.x, s0[3] += b.y;
    s1[0] += c.x, s1[1] += c.y, s1[2] += d.x, s1[3] += d.y;
  }
}

// Per-channel coefficients from the stripes, for the thread that owns channels c0 .. c0 + 3 (element-wise kernels' prologue).
// forward: stripes hold (sum x, sum x^2)
__device__ __forceinline__ void bn_coef_forward4(const BnFinal& F, int c0, float* mean_f, float* invstd_f, float* scale, float* shift,
                                                 double* mean, double* var) {
  double s[4], q[4];
  bn_sum_stripes4(bn_stripes(F.scratch), F.C, c0, s, q);
  const float4 g = __ldg(reinterpret_cast<const float4*>(F.gamma + c0)), b = __ldg(reinterpret_cast<const float4*>(F.beta + c0));
  const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    mean[j] = s[j] * F.inv_m;
    var[j] = fmax(q[j] * F.inv_m - mean[j] * mean[j], 0.0);
    mean_f[j] = static_cast<float>(mean[j]);
    invstd_f[j] = static_cast<float>(rsqrt(var[j] + static_cast<double>(F.eps)));
    // the SAME two fp32 expressions are re-evaluated by the backward kernels from the saved mean / invstd: bit-exact ReLU gate
    scale[j] = __fmul_rn(gg[j], invstd_f[j]);
    shift[j] = __fmaf_rn(-mean_f[j], scale[j], bb[j]);
  }
}
// backward: stripes hold (sum dz, sum dz * xhat);  dx = a * dz + b * x + d,  gate z = fma(x, zs, zt)
__device__ __forceinline__ void bn_coef_backward4(const BnFinal& F, int c0, float* a_f, float* b_f, float* d_f, float* zs, float* zt,
                                                  double* dgamma, double* dbeta) {
  bn_sum_stripes4(bn_stripes(F.scratch), F.C, c0, dbeta, dgamma);
  const float4 g = __ldg(reinterpret_cast<const float4*>(F.gamma + c0)), b = __ldg(reinterpret_cast<const float4*>(F.beta + c0));
  const float4 m = __ldg(reinterpret_cast<const float4*>(F.mean + c0)), is = __ldg(reinterpret_cast<const float4*>(F.invstd + c0));
  const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w}, mm[4] = {m.x, m.y, m.z, m.w}, ii[4] = {is.x, is.y, is.z, is.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double a = static_cast<double>(gg[j]) * ii[j];
    const double mg = dgamma[j] * F.inv_m, mb = dbeta[j] * F.inv_m;
    a_f[j] = static_cast<float>(a);
    b_f[j] = static_cast<float>(-a * ii[j] * mg);
    d_f[j] = static_cast<float>(a * (static_cast<double>(mm[j]) * ii[j] * mg - mb));
    zs[j] = __fmul_rn(gg[j], ii[j]);
    zt[j] = __fmaf_rn(-mm[j], zs[j], bb[j]);
  }
}

}  // namespace dcfp
