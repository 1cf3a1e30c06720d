/* Oracle (TEST INFRASTRUCTURE ONLY): portable exp/log/log10/10^x built only from individually
 * rounded IEEE binary64 operations (compile with -ffp-contract=off).  Restates the "pmath"
 * specification of DESIGN.md (Cody-Waite reduction by ln2 in two parts, degree-5 rational
 * correction for exp, s=f/(2+f) series with 7 coefficients for log) so that a device evaluation of
 * the same sequence is bit-identical.  Used when the oracle is built with -DOX_PMATH
 * (liboracle_pm.so); the default oracle uses libm like the reference. */
#ifndef OX_PMATH_H
#define OX_PMATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline double oxp_from_bits(int64_t b) { double d; memcpy(&d, &b, 8); return d; }
static inline int64_t oxp_bits(double d) { int64_t b; memcpy(&b, &d, 8); return b; }
static inline double oxp_scale2(double x, int k) { return oxp_from_bits(oxp_bits(x) + ((int64_t)k << 52)); }

static inline double oxp_exp(double x) {
  static const double ln2hi = 6.93147180369123816490e-01, ln2lo = 1.90821492927058770002e-10,
                      invln2 = 1.44269504088896338700e+00,
                      P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
                      P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
                      P5 = 4.13813679705723846039e-08;
  double kf = rint(x * invln2);
  int k = (int)kf;
  double hi = x - kf * ln2hi;
  double lo = kf * ln2lo;
  double r = hi - lo;
  double t = r * r;
  double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
  if (k == 0) return y;
  if (k < -1021) return oxp_scale2(y, k + 1000) * 9.33263618503218878990e-302;
  return oxp_scale2(y, k);
}

static inline double oxp_log(double x) {
  static const double ln2hi = 6.93147180369123816490e-01, ln2lo = 1.90821492927058770002e-10,
                      Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                      Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                      Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                      Lg7 = 1.479819860511658591e-01;
  int64_t bits = oxp_bits(x);
  int e = (int)((bits >> 52) & 0x7ff) - 1023;
  int64_t mant = bits & 0x000fffffffffffffLL;
  double m, f, s, z, w, t1, t2, R, hfsq, dk;
  if (mant >= 0x6a09e667f3bcdLL) { e += 1; bits = mant | 0x3fe0000000000000LL; }
  else bits = mant | 0x3ff0000000000000LL;
  m = oxp_from_bits(bits);
  f = m - 1.0;
  s = f / (2.0 + f);
  z = s * s;
  w = z * z;
  t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  R = t2 + t1;
  hfsq = 0.5 * (f * f);
  dk = (double)e;
  return dk * ln2hi - ((hfsq - (s * (hfsq + R) + dk * ln2lo)) - f);
}

static inline double oxp_log10(double x) { return oxp_log(x) * 4.34294481903251816668e-01; }

static inline double oxp_pow10(double x) {
  static const double ln10 = 2.30258509299404590109e+00;
  static const double p10[5] = {1e1, 1e2, 1e4, 1e8, 1e16};
  double nf = floor(x);
  int n = (int)nf, an, b, q, i;
  double f = x - nf;
  double r = oxp_exp(f * ln10);
  double scale = 1.0, big = 1.0;
  an = n < 0 ? -n : n;
  for (b = 0; b < 5; ++b)
    if (an & (1 << b)) scale = scale * p10[b];
  q = an >> 5;
  for (i = 0; i < q; ++i) big = big * 1e32;
  if (n >= 0) return (r * scale) * big;
  return (r / scale) / big;
}
#endif
