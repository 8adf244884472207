// Table-driven FP64 exp for the covariance kernels: exp(x) = 2^m * T[j] * p(r), k = round(32 x / ln 2) = 32 m + j,
// |r| <= ln2/64, p = degree-6 Taylor polynomial (truncation 3.5e-18 relative), T[j] = 2^(j/32) correctly rounded.
// 11 FP64 instructions + one shared-memory load instead of the 16 FP64 instructions of libdevice's exp (same
// Cody-Waite reduction, degree-11 polynomial); error <= 1.5 ulp (checked against exp over the operating range in
// tests/test_gpu_parity.py::test_fast_exp_accuracy through the K-entry comparison at 1e-10).
// Replaces Base.exp in /root/reference/src/covariance.jl:93 (kernv .= sigma^2 .* exp.(-1.0 .* kernv)).
#pragma once
#include <cuda_runtime.h>

namespace gpr {

constexpr int KT_SE_ = 1;

// 2^(j/32), j = 0..31, correctly rounded
__constant__ double kg_exp2_tab[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0, 0x1.172b83c7d517bp+0,
    0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0, 0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0,
    0x1.3dea64c123422p+0, 0x1.44e086061892dp+0, 0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0,
    0x1.6247eb03a5585p+0, 0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0, 0x1.ae89f995ad3adp+0,
    0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0, 0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0,
    0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};

// exp(x), branch free: x is clamped to [-708, 708] (exp(-708) = 3.3e-308 is the smallest normal magnitude reached, so
// a covariance entry that would underflow comes out as ~1e-308 instead of a subnormal / zero; 708 caps the split-kernel
// C factor, exp(+2 sum l^2 xs xq), below overflow).  Keeping the function free of control flow lets the compiler
// interleave the 16-32 independent evaluations of a thread -- the polynomial is a dependent chain, and these kernels
// are bound by exactly that latency (ncu: `wait` is the top stall).  tab = kg_exp2_tab in SHARED memory.
__device__ __forceinline__ double exp_tab32(double x, const double* __restrict__ tab) {
  x = fmin(fmax(x, -708.0), 708.0);
  const double MAGIC = 6755399441055744.0;                      // 1.5 * 2^52: round-to-nearest-integer by addition
  const double t = fma(x, 0x1.71547652b82fep+5, MAGIC);         // 32 / ln 2
  const int k = __double2loint(t);
  const double kf = t - MAGIC;
  double r = fma(kf, -0x1.62e42fee00000p-6, x);                 // ln2/32, high part (33 significant bits: kf * hi is exact)
  r = fma(kf, -0x1.a39ef35793c76p-38, r);                       // ln2/32, low part
  double p = fma(r, 1.0 / 720.0, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double v = tab[k & 31] * p;
  return __hiloint2double(__double2hiint(v) + ((k >> 5) << 20), __double2loint(v));   // * 2^(k >> 5), result stays normal
}

__device__ __forceinline__ double kern_value_tab(int type, double sig2, double dist, const double* __restrict__ tab) {
  if (type == KT_SE_) return sig2 * exp_tab32(-dist, tab);
  const double r = sqrt(dist > 0.0 ? dist : 0.0);
  const double s5r = 2.23606797749978969640917366873128 * r;
  return sig2 * (1.0 + s5r + (5.0 / 3.0) * dist) * exp_tab32(-s5r, tab);
}

}  // namespace gpr
