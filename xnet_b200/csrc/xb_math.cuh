// Elementary functions used by the rate evaluation.
//   FAST   : CUDA libm (exp/log/pow, <= 1-2 ulp).
//   STRICT : portable algorithms built from individually rounded IEEE operations, so that the
//            CPU oracle's restatement of the same algorithms produces identical bits (pmath, see
//            DESIGN.md "strict arithmetic").
#pragma once
#include "xb_device.cuh"
#include "xb_pmath.cuh"

namespace xb {

// One copy of each libm routine per kernel (fast arithmetic): the rate evaluation calls them from a dozen places, most
// of them once per Newton iteration, and inlined copies (45-120 instructions each) are instruction-cache footprint that
// three zones in different phases of the kernel pay for on every pass.  Same code, same results as inlined.
static __device__ __noinline__ double fast_exp(const double x) { return exp(x); }
static __device__ __noinline__ double fast_log(const double x) { return log(x); }
static __device__ __noinline__ double fast_log10(const double x) { return log10(x); }
static __device__ __noinline__ double fast_pow10(const double x) { return pow(10.0, x); }

template <bool STRICT>
struct Math {
  // safe_exp (xnet_util.F90:17-18,69): argument clamped to [-1021*ln2*0.9, 1024*ln2*0.9]
  static __device__ __forceinline__ double safe_exp(double x) {
    const double ln_2 = 0.6931471805599453;
    const double exp_max = 1024.0 * ln_2 * 0.9, exp_min = -1021.0 * ln_2 * 0.9;
    double v = x > exp_min ? x : exp_min;
    v = v < exp_max ? v : exp_max;
    return STRICT ? pm_exp(v) : fast_exp(v);
  }
  static __device__ __forceinline__ double log10(double x) { return STRICT ? pm_log10(x) : fast_log10(x); }
  static __device__ __forceinline__ double pow10(double x) { return STRICT ? pm_pow10(x) : fast_pow10(x); }
  // t09(3,4,6,7) = T9^(-1/3), T9^(1/3), T9^(5/3), ln T9   (xnet_integrate.F90:650-656); which = 0..3
  static __device__ __noinline__ double t9_power(double t9, int which) {
    if (STRICT) {
      const double l = pm_log(t9);
      if (which == 3) return l;
      const double e = (which == 0) ? -1.0 / 3.0 : (which == 1) ? 1.0 / 3.0 : 5.0 / 3.0;
      return pm_exp(__dmul_rn(e, l));
    }
    if (which == 3) return fast_log(t9);
    return pow(t9, (which == 0) ? -1.0 / 3.0 : (which == 1) ? 1.0 / 3.0 : 5.0 / 3.0);
  }
};

}  // namespace xb
