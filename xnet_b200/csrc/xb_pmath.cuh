// pmath: portable exp / log / log10 / 10^x from individually rounded IEEE-754 binary64 operations
// (no FMA, no extended precision), so the same sequence evaluated on the host gives the same bits.
// Algorithms are the classic argument-reduction + minimax-polynomial schemes (Cody-Waite reduction
// by ln2 split in two parts; exp via the rational form r*c/(2-c); log via s=f/(2+f) series).
// Accuracy: about 1 ulp (exp, log), 2-3 ulp (10^x); checked against libm in tests.
#pragma once

namespace xb {

#ifdef __CUDACC__
#define PM_HD __host__ __device__ __forceinline__
#define PM_MUL(a, b) __dmul_rn((a), (b))
#define PM_ADD(a, b) __dadd_rn((a), (b))
#define PM_SUB(a, b) __dadd_rn((a), -(b))
#define PM_DIV(a, b) __ddiv_rn((a), (b))
#else
#define PM_HD inline
#define PM_MUL(a, b) ((a) * (b))
#define PM_ADD(a, b) ((a) + (b))
#define PM_SUB(a, b) ((a) - (b))
#define PM_DIV(a, b) ((a) / (b))
#endif

__device__ __forceinline__ double pm_scale2(double x, int k) {
  // x * 2^k for results in the normal range (|k| <= 1022 after clamping by safe_exp)
  long long bits = __double_as_longlong(x);
  bits += (long long)k << 52;
  return __longlong_as_double(bits);
}

__device__ __forceinline__ double pm_exp(double x) {
  const double ln2hi = 6.93147180369123816490e-01, ln2lo = 1.90821492927058770002e-10;
  const double invln2 = 1.44269504088896338700e+00;
  const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
               P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
               P5 = 4.13813679705723846039e-08;
  // k = nearest integer to x/ln2
  const double kf = rint(PM_MUL(x, invln2));
  const int k = (int)kf;
  const double hi = PM_SUB(x, PM_MUL(kf, ln2hi));
  const double lo = PM_MUL(kf, ln2lo);
  const double r = PM_SUB(hi, lo);
  const double t = PM_MUL(r, r);
  const double c = PM_SUB(r, PM_MUL(t, PM_ADD(P1, PM_MUL(t, PM_ADD(P2, PM_MUL(t, PM_ADD(P3, PM_MUL(t, PM_ADD(P4, PM_MUL(t, P5))))))))));
  const double y = PM_SUB(1.0, PM_SUB(PM_SUB(lo, PM_DIV(PM_MUL(r, c), PM_SUB(2.0, c))), hi));
  if (k == 0) return y;
  if (k < -1021) {  // gradual underflow region: scale in two exact steps
    return PM_MUL(pm_scale2(y, k + 1000), 9.33263618503218878990e-302);
  }
  return pm_scale2(y, k);
}

__device__ __forceinline__ double pm_log(double x) {
  const double ln2hi = 6.93147180369123816490e-01, ln2lo = 1.90821492927058770002e-10;
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
               Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  // x = m * 2^e, m in [sqrt(2)/2, sqrt(2))   (positive normal x only; rate arguments are)
  long long bits = __double_as_longlong(x);
  int e = (int)((bits >> 52) & 0x7ff) - 1023;
  long long mant = bits & 0x000fffffffffffffLL;
  if (mant >= 0x6a09e667f3bcdLL) { e += 1; bits = mant | 0x3fe0000000000000LL; }  // m in [sqrt2/2,1)
  else bits = mant | 0x3ff0000000000000LL;                                         // m in [1,sqrt2)
  const double m = __longlong_as_double(bits);
  const double f = PM_SUB(m, 1.0);
  const double s = PM_DIV(f, PM_ADD(2.0, f));
  const double z = PM_MUL(s, s);
  const double w = PM_MUL(z, z);
  const double t1 = PM_MUL(w, PM_ADD(Lg2, PM_MUL(w, PM_ADD(Lg4, PM_MUL(w, Lg6)))));
  const double t2 = PM_MUL(z, PM_ADD(Lg1, PM_MUL(w, PM_ADD(Lg3, PM_MUL(w, PM_ADD(Lg5, PM_MUL(w, Lg7)))))));
  const double R = PM_ADD(t2, t1);
  const double hfsq = PM_MUL(0.5, PM_MUL(f, f));
  const double dk = (double)e;
  // log(x) = k*ln2hi - ((hfsq - (s*(hfsq+R) + k*ln2lo)) - f)
  return PM_SUB(PM_MUL(dk, ln2hi),
                PM_SUB(PM_SUB(hfsq, PM_ADD(PM_MUL(s, PM_ADD(hfsq, R)), PM_MUL(dk, ln2lo))), f));
}

__device__ __forceinline__ double pm_log10(double x) {
  return PM_MUL(pm_log(x), 4.34294481903251816668e-01);  // log(x) * (1/ln 10)
}

__device__ __forceinline__ double pm_pow10(double x) {
  // 10^x = 10^n * exp(f*ln10), n = floor(x), f in [0,1); 10^n by exact repeated products for |n|<=22
  // and correctly rounded table values composed beyond that.
  const double ln10 = 2.30258509299404590109e+00;
  const double nf = floor(x);
  int n = (int)nf;
  const double f = PM_SUB(x, nf);
  double r = pm_exp(PM_MUL(f, ln10));
  // 10^n via binary powers of ten (all factors <= 1e16 are exact doubles)
  const double p10[5] = {1e1, 1e2, 1e4, 1e8, 1e16};
  int an = n < 0 ? -n : n;
  double scale = 1.0;
  double big = 1.0;  // accumulates 1e32-and-above factors separately to keep exact parts exact
#pragma unroll
  for (int b = 0; b < 5; ++b)
    if (an & (1 << b)) scale = PM_MUL(scale, p10[b]);
  int q = an >> 5;  // multiples of 1e32
  for (int i = 0; i < q; ++i) big = PM_MUL(big, 1e32);
  if (n >= 0) return PM_MUL(PM_MUL(r, scale), big);
  return PM_DIV(PM_DIV(r, scale), big);
}

}  // namespace xb
