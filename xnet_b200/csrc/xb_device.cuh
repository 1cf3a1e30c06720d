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


// Term plan of a segmented sum (ydot over species, Jacobian over non-zero elements).  Records are
// sorted by (segment, reference summation order); thread t of an NT-thread CTA owns the contiguous
// chunk [t*L, (t+1)*L) and record (t*L+k) is stored at k*NT+t, so every load is coalesced.
//   product = (((a * csect[mu]) * y[n0]) * y[n1]) * y[n2]) * y[n3]     (missing factors: y[ny] = 1)
//   ctl: bits 0-1 action after adding the product (0 none, 1 store to destination, 2 store partial),
//        bit 2 = diagonal element (Jacobian), bits 3.. = destination / partial slot
struct DevPlan {
  int nrec, L, nseg, nmulti, npart;
  const double *a;           // [L*NT]
  const uint2 *nn;           // [L*NT]  4 x u16 species indices
  const uint2 *mc;           // [L*NT]  x = csect index, y = ctl
  const int *seg_start;      // [nseg+1] sorted-order record ranges (strict path)
  const int *seg_dest;       // [nseg]   ydot: 4*species+class;  Jacobian: (offset<<1)|diag
  const int *multi_start;    // [nmulti+1] partial-slot ranges of segments split over chunks
  const int *multi_dest;     // [nmulti]  destination as in ctl>>2 (bit 0 = diag)
};

// byte offsets of the zone working set inside the CTA's dynamic shared memory
struct SmemOff { int scal, jac, y, yt, ydot, rhs, rdiag, part, prow, indx, orig, pos, fhs, lu_s, lu_u, lds2, cs, total; };

// structured LU (xb_slu.cuh): sizes, the index tables staged into every CTA's shared memory, species -> position
struct DevSlu {
  double tau, itau;                          // threshold of the column test: tau*|a(i,k)| <= |a(k,k)| (multipliers <= 1/tau); 1/tau
  int ntest;                                 // rows 0..ntest-1 are subject to the threshold test (all but the T9 row)
  int on, n, nstore, nbord, nwl, force_fb;   // nwl: warps that factor; force_fb > 0: every force_fb-th LU takes the dense fallback (tests)
  int off_tab, tab_rel;                      // staged table: byte offset in shared memory; 64-bit entries relative to the store S
  int bo, bs, rhs_off;                       // border rows: base and column stride (doubles); rhs = S + rhs_off
  int ph_rel;                                // ZoneScalars::ph as a 64-bit word offset from S (profiling build)
  int fail_rel;                              // ZoneScalars::slu_fail as a 32-bit word offset from S (negative: the scalars precede the store)
  const uint2 *tab;                          // [n+1] x = cb | cf << 16 (entry n: the right-hand side), y = ncb | nrb << 8
  const int *pos;                            // [n] equation (species, then T9) -> position in the reordered system
  int nround;                                // rounds of the pivot schedule
  const uint2 *sch_warp;                     // [nround+2][nwl]     lp | pv << 16,  nrow | ntest << 4 | k << 8   (SluSchedule, xb_host.h)
  const unsigned *sch_lane;                  // [nround+2][nwl*32]  ap | up << 16
  int pair;                                  // 1: pair-aligned tables, two pivots per round (sch2_*); 0: one pivot per round (sch_*)
  const uint4 *sch2_warp;                    // [nround+3][nwl]     SluSchedule2 (xb_host.h)
  const uint2 *sch2_lane;                    // [nround+3][nwl*32]
  int sol_ngroup;                            // groups of four back-substitution steps (even)
  const uint2 *sol_lane;                     // [sol_ngroup+2][32] four 16-bit entries: address of U(i,k) / of the pivot's reciprocal / 0xffff (slu_solve, phase C)
  const int *sol_out;                        // steps whose column reaches above the 32-row window, descending, then -1000000
};

struct DevNet {
  int ny, nrtot, nffn, msize, ld, nffr, tile;
  int cs_lo;                 // reactions [cs_lo, nrtot) have their rate cached in shared memory
  SmemOff lay;
  DevPlan py, pj;            // ydot and Jacobian plans
  const int *ffr_n1, *ffr_iffn;  // [nffr] FFN reactions in reaction order: first reactant (0-based), table row (1-based)
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
  const double *ffn_ec, *ffn_beta;   // [nffn][143] logft tables (xnet_ffn.F90:100-137)
  const double *ffn_qval;    // [nffn]
  const int *has_logft;      // [nffn] 0 = tabulated rate, 1/2 = logft (ft-/ft+)
  int any_logft;             // some FFN reaction uses the logft form (needs the electron chemical potential)
  double mex_n, mex_p;
  // equation of state and screening (iscrn > 0 or iheat > 0), see xb_eos.cuh
  const double *helm;        // [13][201][541] Helmholtz table planes
  const double *helm_t, *helm_d;  // [201], [541] grid temperatures and densities
  const double *zzi;         // [ny] zz**1.58
  const double *zseq53;      // [nzs] j**(5/3)
  const double *sc_zw, *sc_zi;    // [nrtot] reaction charge parameters zeta_w, zeta_i (0 for one-reactant reactions)
  const int2 *sc_iz;         // [nrtot] x = iz1 | iz2<<8 | iz3<<16 | iz4<<24, y = composite charge
  int nzs;                   // entries of the per-charge strong-screening table
  DevSlu slu;                // structured-LU variant only (xb_slu.cu)
  // matched-pair fluxes (xnet_flux.F90:67-150): set m sums entries [fl_start[m], fl_start[m+1]) in reaction order
  int mflx;
  const int *fl_start;       // [mflx+1]
  const int *fl_k;           // [nrtot] reaction index (concatenated csect order)
  const double *fl_dc;       // [nrtot] signed double-counting factor: +dcflx(1,m) forward, -dcflx(2,m) reverse
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
  const double *zexti;       // [nfix] zext**1.58 (host)
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
  double *hist_flx;          // [nzone][cap][mflx] or null
};

// per-CTA scratch in global memory (L2 resident)
struct DevScratch {
  double *base;            // [ncta][stride]
  size_t stride;           // doubles per CTA
  size_t off_csect, off_rffn, off_qkffn, off_jac, off_prod, off_yo, off_dcsect, off_dlnrffn, off_bdf;  // offsets (doubles)
  unsigned long long *phase;  // [PH_COUNT] cycle counters, see enum Phase
};

// Zone-invariant state lives in constant memory (uploaded before every launch of a context): device
// code reads it through the constant bank instead of chasing kernel-parameter pointers.
// one copy per translation unit (the dense kernels and the structured-LU kernel bind their own)
static __constant__ DevNet c_net;
static __constant__ DevCtl c_ctl;
extern __shared__ __align__(16) unsigned char xb_smem[];

enum Phase { PH_RATES = 0, PH_YDERIV, PH_JAC, PH_LU, PH_BKSUB, PH_OTHER, PH_TOTAL, PH_ZONES,
             PH_LU_PANEL, PH_LU_SWAP, PH_LU_GEMM, PH_SPARE,
             PH_DBG0, PH_DBG_END = PH_DBG0 + 12, PH_COUNT = PH_DBG_END };

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
