// Elementary functions used by the rate evaluation.
//   FAST   : CUDA libm (exp/log/pow, <= 1-2 ulp).
//   STRICT : portable algorithms built from individually rounded IEEE operations, so that the
//            CPU oracle's restatement of the same algorithms produces identical bits (pmath, see
//            DESIGN.md "strict arithmetic").
#pragma once
#include "xb_device.cuh"
#include "xb_pmath.cuh"

namespace xb {

template <bool STRICT>
struct Math {
  // safe_exp (xnet_util.F90:17-18,69): argument clamped to [-1021*ln2*0.9, 1024*ln2*0.9]
  static __device__ __forceinline__ double safe_exp(double x) {
    const double ln_2 = 0.6931471805599453;
    const double exp_max = 1024.0 * ln_2 * 0.9, exp_min = -1021.0 * ln_2 * 0.9;
    double v = x > exp_min ? x : exp_min;
    v = v < exp_max ? v : exp_max;
    return STRICT ? pm_exp(v) : exp(v);
  }
  static __device__ __forceinline__ double log10(double x) { return STRICT ? pm_log10(x) : ::log10(x); }
  static __device__ __forceinline__ double pow10(double x) { return STRICT ? pm_pow10(x) : pow(10.0, x); }
  // t09(3,4,6,7) = T9^(-1/3), T9^(1/3), T9^(5/3), ln T9   (xnet_integrate.F90:650-656)
  static __device__ __forceinline__ void t9_powers(double t9, double *t09) {
    if (STRICT) {
      const double l = pm_log(t9);
      t09[2] = pm_exp(__dmul_rn(-1.0 / 3.0, l));
      t09[3] = pm_exp(__dmul_rn(1.0 / 3.0, l));
      t09[5] = pm_exp(__dmul_rn(5.0 / 3.0, l));
      t09[6] = l;
    } else {
      t09[2] = pow(t9, -1.0 / 3.0);
      t09[3] = pow(t9, 1.0 / 3.0);
      t09[5] = pow(t9, 5.0 / 3.0);
      t09[6] = log(t9);
    }
  }
};

}  // namespace xb
