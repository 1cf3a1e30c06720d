// Device-side data model and per-zone cooperative (one CTA = one zone) routines of xnet_b200.
//
// Each __device__ routine below is executed by ALL threads of a CTA on ONE zone whose working set
// lives in shared memory (Jacobian tile included when N <= ~161); zone-invariant network tables
// are read through L2.  Citations give the XNet routine each one replaces
// (paths relative to /root/reference/source).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "xb_host.h"

namespace xb {


struct DevNet {
  int ny, nrtot, nffn, msize;
  int nr[4], roff[4], nan[4];
  const double *rcT;         // [7][nrtot]  REACLIB coefficients, coefficient-major (coalesced)
  const int *ni;             // [nrtot][6]  nuclei of each reaction, 1-based, 0 = blank
  const int *iwk, *irev;     // [nrtot]
  const int *iffn;           // [nrtot]     1-based FFN table row (class 0 only)
  const int *la, *le;        // [4][ny]     1-based inclusive extents (Fortran values)
  const int *mu[4];          // [nan]       0-based index into the concatenated csect vector
  const double *a[4];        // [nan]
  const int *nn[4][4];       // [nan]       0-based reactant species of each term
  const double *aa, *zz, *mex, *be;  // [ny]
  const double *g, *lng;     // [ny][24]    partition functions and their host-evaluated logs
  const double *dlng;        // [ny][24]    log(g(i)/g(i-1)) for i>=1 (self-heating only)
  const double *t9i;         // [24]
  const double *ffnsum, *ffnenu;  // [nffn][143]
  double mex_n, mex_p;
};

struct DevCtl {
  int iweak0, iscrn, iheat, isolv, kstmx, kitmx, ijac, iconvc;
  double changemx, changemxt, yacc, tolm, tolc, tolt9, ymin, tdel_maxmult;
};

struct DevZones {
  int nfix, nzone;
  const int *nh, *off, *nstart, *zfix;
  const double *th, *t9h, *rhoh;
  const double *tstart, *tstop, *tdelstart, *t9start, *rhostart, *yestart, *ystart, *xext, *aext, *zext;
  const double *rho_scale;
  const double *rs_t, *rs_tdel, *rs_t9, *rs_rho, *rs_t9o, *rs_y, *rs_yo;
  const int *rs_kstep0;
};

struct DevOut {
  double *y, *t, *t9, *rho, *tdel;
  int *kstep, *ktot, *status;
  int hist_cap;
  int *hist_n, *hist_kstep, *hist_kmon;
  double *hist_t, *hist_t9, *hist_rho, *hist_tdel, *hist_edot, *hist_sqnu, *hist_y;
};

// per-CTA scratch in global memory (L2 resident)
struct DevScratch {
  double *base;            // [ncta][stride]
  size_t stride;           // doubles per CTA
  size_t off_csect, off_rffn, off_qkffn, off_jac;  // offsets (doubles)
  unsigned long long *phase;  // [8] cycle counters (rates,yderiv,jac,lu,bksub,other,total,zones)
};

enum Phase { PH_RATES = 0, PH_YDERIV, PH_JAC, PH_LU, PH_BKSUB, PH_OTHER, PH_TOTAL, PH_ZONES };

// ---------------------------------------------------------------------------------------------
// Arithmetic policy.  STRICT: every product and sum individually rounded (no FMA contraction), so
// sums reproduce the sequential CPU evaluation bit for bit.  FAST: FMA wherever it applies.
// ---------------------------------------------------------------------------------------------
template <bool STRICT>
struct Ar {
  static __device__ __forceinline__ double mul(double a, double b) { return STRICT ? __dmul_rn(a, b) : a * b; }
  static __device__ __forceinline__ double add(double a, double b) { return STRICT ? __dadd_rn(a, b) : a + b; }
  static __device__ __forceinline__ double sub(double a, double b) { return STRICT ? __dadd_rn(a, -b) : a - b; }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  // a*b + c
  static __device__ __forceinline__ double mad(double a, double b, double c) {
    return STRICT ? __dadd_rn(__dmul_rn(a, b), c) : fma(a, b, c);
  }
  // c - a*b
  static __device__ __forceinline__ double nmad(double a, double b, double c) {
    return STRICT ? __dadd_rn(c, -__dmul_rn(a, b)) : fma(-a, b, c);
  }
};

}  // namespace xb
