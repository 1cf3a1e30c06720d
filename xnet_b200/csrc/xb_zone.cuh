// One-CTA-per-zone implicit step: rate evaluation, ydot, Jacobian assembly, LU, back-substitution,
// Newton loop, retry ladder, timestep control and the evolve loop, all resident in one kernel.
// Citations: XNet tree, /root/reference/source.
#pragma once
#include "xb_device.cuh"
#include "xb_math.cuh"
#include "xb_lu.cuh"
#include "xb_eos.cuh"
#ifdef XB_SLU
#include "xb_slu.cuh"
#endif

namespace xb {

// Per-zone scalars (shared memory).  Written by thread 0 between barriers, read by everyone.
// per-zone scalars of the BDF integrator (xb_bdf.cuh; module variables of xnet_integrate_bdf.F90 :66-127)
struct BdfScalars {
  int q, deltaq, q_age, j_age, p_age, nit_nrslv, nit_nr, nit_ts, nerr_ts, ncf_ts, ierr_nr, ierr_ts, retry_ts, started;
  int refactor, rebuild, iterate, restore, rescale, reload, conv;
  double hvec[6], lvec[6], dt_scale, etaq[3], eta, eta_next, gam, gamhat, gamratio, t9dot0, acnrm, crate, tq[4], tq2save;
  double del, delp, a1;
};

struct ZoneScalars {
  double t, tt, to, tdel, tdel_next, tdel_old, tdelstart, tstart, tstop;
  double t9, t9t, t9o, t9dot, rho, rhot, rhoo, ye, yet, yeo, cv, etae, detaedt9;
  double xext, aext, zext, zexti, rho_scale;
  // composition moments and plasma parameters of the current rate evaluation (iscrn > 0 or iheat > 0)
  double mom[2], ytot, abar, zbar, z2bar, zibar, ztilde, zinter, lambda0, gammae, dztildedt9;
  double hw0, hi0, dlnhw0dt9, dlnhi0dt9;
  double rdt, mult, xtot_init, enm, enold, edot, sqnu;
  double t9rhs, dt9, relt9;
  double red[4];             // results of block-wide reductions
  double t09[4];             // T9^(-1/3), T9^(1/3), T9^(5/3), ln T9 of the current rate evaluation
  const double *th, *t9h, *rhoh;
  double *g_csect, *g_rffn, *g_qkffn, *g_prod, *g_jac, *g_yo;  // per-CTA global scratch (set in the kernel prologue)
  double *g_dcsect, *g_dlnrffn;   // self-heating: d csect / d T9 [nrtot], d ln rffn / d T9 [nffn]
  double *g_bdf;                  // BDF (isolv = 3): BdfScalars, then the Nordsieck arrays and work vectors (kBdfVectors x msize doubles)
  int nh, has_scale, nt, ntt, nto, ints, iweak;
  int kmon[5], ktot[5];
  int inr, its, flag, piv, info, tile_ctr;
  int rhs_in_jac;            // global-memory Jacobian: the right-hand side rides as column N of the per-CTA scratch (fused path)
  int slu_fail, slu_fail_k;  // structured LU: threshold test failed, first violated column.  slu_fail, slu_fail_k, slu_ipb stay adjacent (slu_factor)
  int slu_ipb[8];            // interchanges of the border block
  int lu_dense, nlu_zone;   // factors are the dense fallback's / factorisations so far
  long long ph[PH_COUNT];    // see enum Phase
  long long tmark, dmark;
};

// Views of the zone working set
struct ZoneMem {
  ZoneScalars *s;
  double *jac;      // [N columns][ld] column-major (shared tile, or global scratch for large N)
  int ld, tile;     // leading dimension (odd); tile = 1 when the Jacobian is the shared-memory tile
  double *y, *yo, *yt, *ydot, *rhs;  // [ny]; yt has ny+1 entries (yt[ny] = 1); rhs [N]
  double *part;     // partial sums of the term plans / scratch of block reductions
  double *prow;     // [kPbufDoubles] panel buffer of the LU (xb_lu.cuh); [2][32] block values of the solve
  double *rdiag;    // [N] reciprocal diagonal of U (fast arithmetic only)
  double *prod;     // per-CTA global scratch for the strict term products
  double *gg;       // [ny+1] partition functions (aliases part)
  double *cs_sm;    // shared-memory cache of csect[cs_lo..nrtot)
  double *fhs;      // [2][nzs] strong-screening energy per charge and its T9 derivative (iscrn > 0)
  double *lu_s, *lu_u;   // two-level LU of global-memory Jacobians: panel [32][lds2] and U block [ncol][32] (else null)
  int lds2;
  double *csect;    // [nrtot] (global scratch)
  double *rffn;     // [nffn]
  double *dcsect, *dlnrffn;   // self-heating only (global scratch)
  double *qkffn;    // [nffn+1]
  int *indx;        // [N] 0-based pivot rows
  int *orig, *pos;  // [N] slot permutation of the Jacobian rows: slot s holds row orig[s]; row i sits in slot pos[i]
};

// The zone working set is addressed from the dynamic shared-memory symbol with offsets that the host
// placed in constant memory, so every access compiles to LDS/STS with 32-bit addressing.
__device__ __forceinline__ ZoneMem zone_mem() {
  const SmemOff &L = c_net.lay;
  ZoneMem zm;
  zm.s = reinterpret_cast<ZoneScalars *>(xb_smem + L.scal);
  zm.ld = c_net.ld; zm.tile = c_net.tile;
  zm.jac = c_net.tile ? reinterpret_cast<double *>(xb_smem + L.jac) : zm.s->g_jac;
  zm.y = reinterpret_cast<double *>(xb_smem + L.y);
  zm.yo = zm.s->g_yo;
  zm.yt = reinterpret_cast<double *>(xb_smem + L.yt);
  zm.ydot = reinterpret_cast<double *>(xb_smem + L.ydot);
  zm.rhs = reinterpret_cast<double *>(xb_smem + L.rhs);     // tile mode: column N of the tile
  if (!c_net.tile && zm.s->rhs_in_jac) zm.rhs = zm.s->g_jac + (size_t)c_net.msize * c_net.ld;
  zm.gg = reinterpret_cast<double *>(xb_smem + L.part);     // lives only inside cross_sect
  zm.cs_sm = reinterpret_cast<double *>(xb_smem + L.cs);
  zm.fhs = reinterpret_cast<double *>(xb_smem + L.fhs);
  zm.lds2 = L.lds2;
  zm.lu_s = L.lds2 ? reinterpret_cast<double *>(xb_smem + L.lu_s) : nullptr;
  zm.lu_u = L.lds2 ? reinterpret_cast<double *>(xb_smem + L.lu_u) : nullptr;
  zm.rdiag = reinterpret_cast<double *>(xb_smem + L.rdiag);
  zm.part = reinterpret_cast<double *>(xb_smem + L.part);
  zm.prow = reinterpret_cast<double *>(xb_smem + L.prow);
  zm.indx = reinterpret_cast<int *>(xb_smem + L.indx);
  zm.orig = reinterpret_cast<int *>(xb_smem + L.orig);
  zm.pos = reinterpret_cast<int *>(xb_smem + L.pos);
  zm.csect = zm.s->g_csect; zm.rffn = zm.s->g_rffn; zm.qkffn = zm.s->g_qkffn; zm.prod = zm.s->g_prod;
  zm.dcsect = zm.s->g_dcsect; zm.dlnrffn = zm.s->g_dlnrffn;
  return zm;
}
#define XB_ZONE_PROLOGUE                \
  const DevNet &net = c_net;            \
  const DevCtl &ctl = c_ctl;            \
  ZoneMem zm = zone_mem();              \
  (void)net; (void)ctl;

// fine-grained debug timers (thread 0): XB_DBG(zm, i) adds the cycles since the previous mark to slot i
#define XB_DBG(zm, i)                                             \
  if (threadIdx.x == 0) {                                         \
    long long now_ = clock64();                                   \
    (zm).s->ph[PH_DBG0 + (i)] += now_ - (zm).s->dmark;            \
    (zm).s->dmark = now_;                                         \
  }
#define XB_DBG_MARK(zm) \
  if (threadIdx.x == 0) (zm).s->dmark = clock64();

#define XB_PHASE_BEGIN(zm) \
  if (threadIdx.x == 0) (zm).s->tmark = clock64();
#define XB_PHASE_END(zm, which)                                   \
  if (threadIdx.x == 0) {                                         \
    long long now_ = clock64();                                   \
    (zm).s->ph[which] += now_ - (zm).s->tmark;                    \
    (zm).s->tmark = now_;                                         \
  }

// --- trajectory interpolation: t9rhofind (xnet_conditions.F90:50-102) --------------------------
// First n with tf <= th(n) (trajectory times ascend; same n as the reference's linear scan from 1).
// nlo: a known lower bound of the answer -- the callers inside a step pass the index of the current time (tf >= t
// implies n(tf) >= n(t)).  The next three points are probed with independent loads (one L2 round trip; a step
// rarely crosses more), the rest by bisection: ~2 dependent memory latencies per call instead of ~11.
template <bool STRICT>
__device__ void t9rhofind(const ZoneScalars *s, double tf, int *nf, double *t9f, double *rhof, const int nlo = 1) {
  using A = Ar<STRICT>;
  const int ns = s->nh;
  const double sc = s->rho_scale;
  const bool hs = s->has_scale != 0;
  auto rhoat = [&](int n) { return hs ? A::mul(s->rhoh[n - 1], sc) : s->rhoh[n - 1]; };
  if (ns == 1) { *t9f = s->t9h[0]; *rhof = rhoat(1); *nf = 1; return; }
  int lo = (nlo < 1) ? 1 : (nlo > ns + 1 ? ns + 1 : nlo), hi = ns + 1;  // answer in [lo, ns+1]
  if (lo > 1) {
    double a[3];
    const double below = s->th[lo - 2];     // the bound is checked, not trusted: it holds iff th(lo-1) < tf
#pragma unroll
    for (int p = 0; p < 3; ++p) a[p] = (lo + p <= ns) ? s->th[lo + p - 1] : 1.7976931348623157e308;
    if (!(below < tf)) lo = 1;
    else {
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        if (lo < hi) {
          if (tf <= a[p]) hi = lo; else ++lo;
        }
      }
    }
  }
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (tf <= s->th[mid - 1]) hi = mid; else lo = mid + 1;
  }
  const int n = lo;
  *nf = n;
  if (n > 1 && n <= ns) {
    double rdt = A::div(1.0, A::sub(s->th[n - 1], s->th[n - 2]));
    double dt = A::sub(tf, s->th[n - 2]);
    double dt9 = A::sub(s->t9h[n - 1], s->t9h[n - 2]);
    double drho = A::sub(rhoat(n), rhoat(n - 1));
    *t9f = A::add(A::mul(A::mul(dt, rdt), dt9), s->t9h[n - 2]);
    *rhof = A::add(A::mul(A::mul(dt, rdt), drho), rhoat(n - 1));
  } else if (n == 1) {
    *t9f = s->t9h[0]; *rhof = rhoat(1);
  } else {
    *t9f = s->t9h[ns - 1]; *rhof = rhoat(ns);
  }
}

// --- block-wide helpers ----------------------------------------------------------------------------
// (The thread-strided loops of this file carry `#pragma unroll 1`: they make one or two trips, and the four-fold unrolled
// bodies the compiler would emit are instruction-cache footprint for nothing -- DESIGN 4c.)
// Two sums over k < n of terms produced by f(k, ta, tb).  STRICT: terms are formed in parallel and
// added by one thread in ascending k from 0.0 (the reference's sequential accumulation, bit for bit);
// FAST: shuffle tree.  Results land in out2[0..1] (shared) and are visible to every thread on return.
template <bool STRICT, class F>
__device__ __forceinline__ void block_sum2(const int n, F f, double *buf, double *out2) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (STRICT) {
#pragma unroll 1
    for (int k = tid; k < n; k += nt) { double ta, tb; f(k, ta, tb); buf[k] = ta; buf[n + k] = tb; }
    __syncthreads();
    if (tid < 2) {
      const double *b = buf + tid * n;
      double s = 0.0;
      for (int k = 0; k < n; ++k) s = __dadd_rn(s, b[k]);
      out2[tid] = s;
    }
    __syncthreads();
  } else {
    double sa = 0.0, sb = 0.0;
#pragma unroll 1
    for (int k = tid; k < n; k += nt) { double ta, tb; f(k, ta, tb); sa += ta; sb += tb; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      sa += __shfl_xor_sync(kFull, sa, off);
      sb += __shfl_xor_sync(kFull, sb, off);
    }
    const int w = tid >> 5, nw = nt >> 5;
    if ((tid & 31) == 0) { buf[w] = sa; buf[32 + w] = sb; }
    __syncthreads();
    if (tid < 32) {
      double va = (tid < nw) ? buf[tid] : 0.0, vb = (tid < nw) ? buf[32 + tid] : 0.0;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        va += __shfl_xor_sync(kFull, va, off);
        vb += __shfl_xor_sync(kFull, vb, off);
      }
      if (tid == 0) { out2[0] = va; out2[1] = vb; }
    }
    __syncthreads();
  }
}

// First index of the maximum of f(k), k < n (Fortran maxloc).  Result in out[0] (value) and *iout.
template <class F>
__device__ __forceinline__ void block_argmax_first(const int n, F f, double *buf, double *out, int *iout) {
  const int tid = threadIdx.x, nt = blockDim.x;
  double bv = -1.0; int bi = 0x7fffffff;
#pragma unroll 1
  for (int k = tid; k < n; k += nt) { const double v = f(k); if (v > bv) { bv = v; bi = k; } }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_xor_sync(kFull, bv, off);
    const int oi = __shfl_xor_sync(kFull, bi, off);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  const int w = tid >> 5, nw = nt >> 5;
  if ((tid & 31) == 0) { buf[w] = bv; reinterpret_cast<int *>(buf + 32)[2 * w] = bi; }
  __syncthreads();
  if (tid == 0) {
    double v = buf[0]; int i = reinterpret_cast<int *>(buf + 32)[0];
    for (int q = 1; q < nw; ++q) {
      const double ov = buf[q]; const int oi = reinterpret_cast<int *>(buf + 32)[2 * q];
      if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    out[0] = v; *iout = i;
  }
  __syncthreads();
}

// --- cross_sect (xnet_integrate.F90:569-927) incl. update_eos' Ye (xnet_abundances.F90:137-166),
//     update_iweak (:300-311), partf (xnet_data.F90:217-245), ffn_rate (xnet_ffn.F90:199-299) ------
template <bool STRICT>
__device__ __noinline__ void cross_sect_cta() {
  using A = Ar<STRICT>;
  using Mx = Math<STRICT>;
  XB_ZONE_PROLOGUE
  ZoneScalars *s = zm.s;
  const int tid = threadIdx.x, nt = blockDim.x, ny = net.ny;
  const double t9t = s->t9t, rhot = s->rhot;
  XB_DBG_MARK(zm);
  {
    // Ye from the composition moments (xnet_abundances.F90:137-166); sums in the reference order
    const double *aa = net.aa, *zz = net.zz, *yt = zm.yt;
    block_sum2<STRICT>(ny, [&](int k, double &ta, double &tb) { ta = A::mul(yt[k], aa[k]); tb = A::mul(yt[k], zz[k]); },
                       zm.part, s->red);
  }
  const bool eos_on = ctl.iscrn > 0 || ctl.iheat > 0;
  if (eos_on) {
    // remaining composition moments: sum Y, sum Y Z^2, sum Y Z^1.58 (xnet_abundances.F90:137-166)
    const double *zz = net.zz, *zzi = net.zzi, *yt = zm.yt;
    block_sum2<STRICT>(ny, [&](int k, double &ta, double &tb) { ta = yt[k]; tb = A::mul(yt[k], A::mul(zz[k], zz[k])); },
                       zm.part, s->red + 2);
    block_sum2<STRICT>(ny, [&](int k, double &ta, double &tb) { ta = A::mul(yt[k], zzi[k]); tb = 0.0; }, zm.part, s->mom);
  }
  // With screening / self-heating the scalar EOS chain (cold code, ~25k cycles) is moved to the first lane of the
  // last warp and overlapped with everything that does not depend on it: the FFN table rates and the
  // one-reactant rates of the other warps.  logft rates need the electron chemical potential first: no split.
  const bool split = eos_on && nt >= 64 && !net.any_logft;
  // EOS / plasma / screening scalars of this rate evaluation: one thread (scalar chain)
  auto eos_chain = [&]() {
    const double yext = A::div(s->xext, s->aext);
    const double atot = A::add(s->red[0], s->xext);
    const double ztot = A::mad(yext, s->zext, s->red[1]);
      // update_eos -> eos_interface -> eosx (xnet_eos_starkiller.F90:64-135) and eos_screen (:244-323)
      const double ntot = A::add(s->red[2], yext);
      const double z2tot = A::add(s->red[3], A::mul(A::mul(yext, s->zext), s->zext));
      const double zitot = A::add(s->mom[0], A::mul(yext, s->zexti));
      s->ytot = ntot; s->abar = A::div(atot, ntot); s->zbar = A::div(ztot, ntot);
      s->z2bar = A::div(z2tot, ntot); s->zibar = A::div(zitot, ntot);
      const double temp = A::mul(t9t, 1e9);
      double e, de;
      helm_eta<STRICT>(net, temp, rhot, s->yet, &e, &de);
      s->etae = e; s->detaedt9 = A::mul(de, 1e9);
      if (ctl.iheat > 0) {   // cv only enters t9dot; without self-heating the reference computes and ignores it
        const double amu = 1.0 / (6.022140857e+23 * 1.6021766208e-06);
        double c;
        helm_cv<STRICT>(net, temp, rhot, s->abar, s->zbar, s->yet, &c);
        s->cv = A::mul(A::mul(c, amu), 1e9);
      }
      if (ctl.iscrn > 0) {
        using R = Num<STRICT>;
        const double bi = 0.86, kbi = 0.38, thbim2 = 0.58;
        double sratio, dz;
        salpeter_ratio<STRICT>(s->etae, ctl.iheat > 0, &sratio, &dz);
        const R zbar(s->zbar);
        const R ztilde = EMath<STRICT>::sqrt(R(s->z2bar) + R(sratio) * zbar);
        const R dzt = 0.5 * zbar / ztilde * dz * s->detaedt9;
        s->ztilde = ztilde.v; s->dztildedt9 = dzt.v;
        plasma<STRICT>(t9t, rhot, s->ytot, s->yet, s->zbar, s->zibar, ztilde.v, &s->zinter, &s->lambda0, &s->gammae);
        s->hw0 = (ztilde * s->lambda0).v;
        s->hi0 = (kbi * R(s->zinter) * EMath<STRICT>::pow(R(s->lambda0), bi)).v;
        s->dlnhw0dt9 = (dzt / ztilde - 1.5 / R(t9t)).v;
        s->dlnhi0dt9 = (-thbim2 * dzt / ztilde - 1.5 * bi / R(t9t)).v;
      }
  };
  if (tid == 0) {
    const double yext = A::div(s->xext, s->aext);
    const double atot = A::add(s->red[0], s->xext);
    const double ztot = A::mad(yext, s->zext, s->red[1]);
    s->yet = A::div(ztot, atot);
    s->cv = 0.0; s->etae = 0.0; s->detaedt9 = 0.0;
    s->iweak = (t9t < 0.01 && ctl.iweak0 != 0) ? -abs(ctl.iweak0) : ctl.iweak0;
    s->ktot[4] += 1;
    if (eos_on && !split) eos_chain();
  }
  // temperature powers: four threads of different warps evaluate one transcendental each
  // t09(3,4,6,7) = T9^(-1/3), T9^(1/3), T9^(5/3), ln T9   (xnet_integrate.F90:650-656)
  if ((tid & 31) == 0)
    for (int which = tid >> 5; which < 4; which += nt >> 5) s->t09[which] = Mx::t9_power(t9t, which);
  // partition functions
  {
    int ii = 1;
    for (; ii <= kNG; ++ii) if (t9t <= net.t9i[ii - 1]) break;
    if (tid == 0) zm.gg[0] = 1.0;
    if (ii == 1) {
#pragma unroll 1
      for (int k = tid; k < ny; k += nt) zm.gg[k + 1] = net.g[k * kNG];
    } else if (ii == kNG + 1) {
#pragma unroll 1
      for (int k = tid; k < ny; k += nt) zm.gg[k + 1] = net.g[k * kNG + kNG - 1];
    } else {
      double rdt9 = A::div(A::sub(t9t, net.t9i[ii - 2]), A::sub(net.t9i[ii - 1], net.t9i[ii - 2]));
      double omr = A::sub(1.0, rdt9);
#pragma unroll 1
      for (int k = tid; k < ny; k += nt) {
        double v = A::add(A::mul(rdt9, net.lng[k * kNG + ii - 1]), A::mul(omr, net.lng[k * kNG + ii - 2]));
        zm.gg[k + 1] = Mx::safe_exp(v);
      }
    }
    if (ctl.iheat > 0) {
      // d ln g / d T9 (xnet_data.F90:247-266): slope of ln g on the bracketing interval (end intervals clamped)
      const int idx = (ii == 1) ? 1 : (ii == kNG + 1) ? kNG - 1 : ii - 1;
      const double dt = A::sub(net.t9i[idx], net.t9i[idx - 1]);
      double *dl = zm.gg + (ny + 1);
      if (tid == 0) dl[0] = 0.0;
#pragma unroll 1
      for (int k = tid; k < ny; k += nt) dl[k + 1] = A::div(net.dlng[k * kNG + idx], dt);
    }
  }
  __syncthreads();
  XB_DBG(zm, 4);
  const double ene = A::mul(s->yet, rhot);
  const int iweak = s->iweak;
  const bool worker = !split || tid < nt - 32;       // threads that run the FFN block and the first rate pass
  const int wnt = split ? nt - 32 : nt;
  if (split && tid == nt - 32) {
#ifdef XB_EOS_TIMER
    const long long t0_ = clock64();
    eos_chain();
    s->ph[PH_DBG0 + 8] += clock64() - t0_;
#else
    eos_chain();
#endif
  }
  auto fhs_table = [&]() {
    // strong-screening energy per charge, DeWitt & Slattery (2003) fit (xnet_screening.F90:225-235)
    using R = Num<STRICT>;
    const double cds1 = -0.899172, cds2 = 0.602249, cds3 = -0.274823, cds4 = -1.401915, cds5 = 0.3230064;
    double *fhs = zm.fhs, *dfhs = zm.fhs + net.nzs;
    const double gammae = s->gammae;
#pragma unroll 1
    for (int j = tid; j < net.nzs; j += nt) {
      if (j == 0) { fhs[0] = 0.0; dfhs[0] = 0.0; continue; }
      const R gammaz = R(gammae) * net.zseq53[j];
      const R gammaz5 = EMath<STRICT>::pow(gammaz, cds5);
      const R lgammaz = EMath<STRICT>::log(gammaz);
      fhs[j] = (cds1 * gammaz + cds2 / cds5 * gammaz5 + cds3 * lgammaz + cds4).v;
      dfhs[j] = (-(cds1 * gammaz + cds2 * gammaz5 + cds3) / R(t9t)).v;
    }
  };
  if (ctl.iscrn > 0 && !split) fhs_table();

  // FFN tabulated weak rates
  if (net.nffn > 0 && worker) {
    const double t9grid[13] = {0.01, 0.1, 0.2, 0.4, 0.7, 1.0, 1.5, 2.0, 3.0, 5.0, 10.0, 30.0, 100.0};
    int i = 1;
    for (; i <= 13; ++i) if (t9t <= t9grid[i - 1]) break;
    const int lt1 = (i - 1 > 1) ? i - 1 : 1;
    const double enel = Mx::log10(ene);
    int le1 = (int)enel; if (le1 < 1) le1 = 1; if (le1 > 10) le1 = 10;
    const double dt9 = A::sub(t9grid[lt1], t9grid[lt1 - 1]);
    const double rdt9 = A::div(A::sub(t9t, t9grid[lt1 - 1]), dt9);
    const double dene = A::sub((double)(le1 + 1), (double)le1);
    const double rdene = A::div(A::sub(enel, (double)le1), dene);
    const int i1 = 13 * (le1 - 1) + lt1 - 1, i2 = i1 + 1, i3 = 13 * le1 + lt1 - 1, i4 = i3 + 1;  // 0-based
#pragma unroll 1
    for (int k = tid; k < net.nffn; k += wnt) {
      const double *enu = net.ffnenu + (size_t)k * kNGrid, *sum = net.ffnsum + (size_t)k * kNGrid;
      double qdr1 = A::sub(enu[i2], enu[i1]), qdr2 = A::sub(enu[i4], enu[i3]);
      double qr1 = A::mad(rdt9, qdr1, enu[i1]), qr2 = A::mad(rdt9, qdr2, enu[i3]);
      double qk = A::mad(rdene, A::sub(qr2, qr1), qr1);
      zm.qkffn[k + 1] = (qk > -90.0) ? Mx::pow10(qk) : 0.0;
      if (net.has_logft[k] > 0) {
        // logft rates (xnet_ffn.F90:222-289): rate = 10^beta + ln2 * phase-space integral / 10^logft
        using R = Num<STRICT>;
        const double ln_2 = 0.6931471805599453, ln_10 = 2.302585092994046, rfmin = 1.0e-30, lrfmin = -30.0;
        const double bok = 8.6173303e-02, m_e = 0.5109989461;
        const double cheme = -fabs((R(s->etae) * bok * t9t + m_e).v);   // sign(cheme, -has_logft-1.5)
        double phasei, dphaseidt9;
        effphase<STRICT>(t9t, cheme, net.ffn_qval[k], &phasei, &dphaseidt9);
        const double *ec = net.ffn_ec + (size_t)k * kNGrid, *be = net.ffn_beta + (size_t)k * kNGrid;
        const R rd(rdt9), re(rdene);
        const R dr1e = R(ec[i2]) - ec[i1], dr2e = R(ec[i4]) - ec[i3];
        R r1 = R(ec[i1]) + rd * dr1e, r2 = R(ec[i3]) + rd * dr2e;
        R rf_ec = r1 + re * (r2 - r1);
        if (phasei > rfmin && rf_ec.v > lrfmin) rf_ec = ln_2 * R(phasei) / R(Mx::pow10(rf_ec.v));
        else rf_ec = R(0.0);
        const R dr1b = R(be[i2]) - be[i1], dr2b = R(be[i4]) - be[i3];
        r1 = R(be[i1]) + rd * dr1b; r2 = R(be[i3]) + rd * dr2b;
        R rf_beta = r1 + re * (r2 - r1);
        rf_beta = (rf_beta.v > lrfmin) ? R(Mx::pow10(rf_beta.v)) : R(0.0);
        const R rft = rf_beta + rf_ec;
        double dln = 0.0;
        if (rf_beta.v < rfmin && rf_ec.v < rfmin) zm.rffn[k] = 0.0;
        else {
          zm.rffn[k] = rft.v;
          if (ctl.iheat > 0) {
            if (rf_ec.v < rfmin) dln = (ln_10 * (re * dr2b + (1.0 - re) * dr1b) / dt9).v;
            else {
              const R drbeta_dt = rf_beta * ln_10 * (re * dr2b + (1.0 - re) * dr1b) / dt9;
              const R drec_dt = rf_ec * (-ln_10 * (re * dr2e + (1.0 - re) * dr1e) / dt9 + R(dphaseidt9) / phasei);
              dln = ((drbeta_dt + drec_dt) / rft).v;
            }
          }
        }
        if (ctl.iheat > 0) zm.dlnrffn[k] = dln;
        continue;
      }
      double dr1 = A::sub(sum[i2], sum[i1]), dr2 = A::sub(sum[i4], sum[i3]);
      double r1 = A::mad(rdt9, dr1, sum[i1]), r2 = A::mad(rdt9, dr2, sum[i3]);
      double rf = A::mad(rdene, A::sub(r2, r1), r1);
      zm.rffn[k] = (rf < -30.0) ? 0.0 : Mx::pow10(rf);
      if (ctl.iheat > 0) {   // d ln rffn / d T9 (xnet_ffn.F90:290-297)
        const double ln_10 = 2.302585092994046;
        zm.dlnrffn[k] = (rf < -30.0) ? 0.0
            : A::div(A::mul(ln_10, A::add(A::mul(rdene, dr2), A::mul(A::sub(1.0, rdene), dr1))), dt9);
      }
    }
    if (tid == 0) zm.qkffn[0] = 0.0;
  }
  if (!split) __syncthreads();
  else if (worker) bar_named(2, nt - 32);            // the FFN rates are visible to the workers' rate pass

  XB_DBG(zm, 5);
  // REACLIB rates: one thread per reaction, coalesced coefficient loads
  double t09[7];
  t09[0] = 1.0;
  t09[1] = A::div(1.0, t9t);
  t09[2] = s->t09[0]; t09[3] = s->t09[1]; t09[5] = s->t09[2]; t09[6] = s->t09[3];
  t09[4] = t9t;
  const double rho2 = A::mul(rhot, rhot), rho3 = A::mul(rho2, rhot);
  // d t09 / d T9 (xnet_integrate.F90:658-665); the fractional powers follow from the published ones:
  // T9^(-4/3) etc. are evaluated as the reference does, by pow
  double dt09[7] = {0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0};
  const bool heat = ctl.iheat > 0;
  if (heat) {
    using R = Num<STRICT>;
    const R t(t9t);
    dt09[1] = (-(1.0 / (t * t))).v;
    dt09[2] = (-EMath<STRICT>::pow(t, -4.0 / 3.0) / 3.0).v;
    dt09[3] = (EMath<STRICT>::pow(t, -2.0 / 3.0) / 3.0).v;
    dt09[5] = (EMath<STRICT>::pow(t, 2.0 / 3.0) * 5.0 / 3.0).v;
    dt09[6] = (1.0 / t).v;
  }
  const double *dlg = zm.gg + (ny + 1);
  const int nrtot = net.nrtot;
  const double *__restrict__ rcT = net.rcT;
  const int2 *__restrict__ ni2 = reinterpret_cast<const int2 *>(net.ni);
  const int *__restrict__ iwkp = net.iwk, *__restrict__ irevp = net.irev;
  const double *gg = zm.gg;
#ifndef XB_RATE_UNROLL
#define XB_RATE_UNROLL 1
#endif
  constexpr int kRateUnroll = XB_RATE_UNROLL;
  // split: pass 0 = one-reactant rates by the worker warps while the EOS chain runs, pass 1 = the screened
  // classes by everybody once the plasma parameters and the per-charge table exist (one copy of the loop body)
  for (int pass = 0; pass < (split ? 2 : 1); ++pass) {
  int kbeg = 0, kend = nrtot, kstride = nt;
  if (split) {
    if (pass == 0) { kend = net.roff[1]; kstride = nt - 32; if (!worker) kbeg = kend; }
    else {
      kbeg = net.roff[1];
#ifdef XB_EOS_TIMER
      long long t1_ = 0;
      if (tid == 0) t1_ = clock64();
#endif
      __syncthreads();                                 // EOS chain and pass 0 done
#ifdef XB_EOS_TIMER
      if (tid == 0) { const long long n_ = clock64(); s->ph[PH_DBG0 + 9] += n_ - t1_; t1_ = n_; }
#endif
      if (ctl.iscrn > 0) { fhs_table(); __syncthreads(); }
#ifdef XB_EOS_TIMER
      if (tid == 0) s->ph[PH_DBG0 + 10] += clock64() - t1_;
#endif
    }
  }
#pragma unroll kRateUnroll
  for (int k = kbeg + tid; k < kend; k += kstride) {
    // every load of the reaction is issued up front (one L2 round trip per reaction, not three)
    const int iwk = __ldg(iwkp + k), irev = __ldg(irevp + k);
    const int2 n01 = __ldg(ni2 + 3 * k), n23 = __ldg(ni2 + 3 * k + 1), n45 = __ldg(ni2 + 3 * k + 2);
    double rc[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) rc[j] = __ldg(rcT + (size_t)j * nrtot + k);
    const int c = (k >= net.roff[3]) ? 3 : (k >= net.roff[2]) ? 2 : (k >= net.roff[1]) ? 1 : 0;
    // the screening data of the reaction travels with the coefficients (same L2 round trip, not three more behind zw > 0)
    const bool scr = ctl.iscrn > 0 && c > 0;
    const double zw = scr ? __ldg(net.sc_zw + k) : 0.0, zi = scr ? __ldg(net.sc_zi + k) : 0.0;
    const int2 izp = scr ? __ldg(net.sc_iz + k) : make_int2(0, 0);
    double cs, dcs = 0.0;
    if ((iweak < 0 && iwk == 0) || (iweak == 0 && iwk != 0)) {
      cs = 0.0;
    } else {
      double lam = 0.0, dlam = 0.0, dhscr = 0.0;
#pragma unroll
      for (int j = 0; j < 7; ++j) lam = A::mad(rc[j], t09[j], lam);
      if (heat) {
#pragma unroll
        for (int j = 0; j < 7; ++j) dlam = A::mad(rc[j], dt09[j], dlam);
      }
      // h = ascrn*h + lambda (xnet_integrate.F90:711-762); h = screening factor when iscrn > 0
      double hscr = 0.0;
      if (scr) {
        if (zw > 0.0) {
          const double *fhs = zm.fhs, *dfhs = zm.fhs + net.nzs;
          const int i1 = izp.x & 255, i2 = (izp.x >> 8) & 255, i3 = (izp.x >> 16) & 255, i4 = (izp.x >> 24) & 255, ic = izp.y;
          const double lambda = A::mul(s->hw0, zw);
          const double hi = A::mul(s->hi0, zi);
          double hs, dhs;
          if (c == 1) {
            hs = A::sub(A::add(fhs[i1], fhs[i2]), fhs[ic]);
            dhs = A::sub(A::add(dfhs[i1], dfhs[i2]), dfhs[ic]);
          } else if (c == 2) {   // the reference omits fhs(iz32) here (xnet_screening.F90:262); reproduced
            hs = A::sub(A::add(fhs[i1], fhs[i3]), fhs[ic]);
            dhs = A::sub(A::add(A::add(dfhs[i1], dfhs[i2]), dfhs[i3]), dfhs[ic]);
          } else {
            hs = A::sub(A::add(A::add(A::add(fhs[i1], fhs[i2]), fhs[i3]), fhs[i4]), fhs[ic]);
            dhs = A::sub(A::add(A::add(A::add(dfhs[i1], dfhs[i2]), dfhs[i3]), dfhs[i4]), dfhs[ic]);
          }
          screen_blend<STRICT>(lambda, lambda, hi, hs, A::mul(lambda, s->dlnhw0dt9), A::mul(hi, s->dlnhi0dt9), dhs, &hscr, &dhscr);
        }
      }
      const double ascrn = (ctl.iscrn > 0) ? 1.0 : 0.0;
      const double h = A::add(A::mul(ascrn, hscr), lam);
      double rpf = 1.0, dlnrpf = 0.0;
      if (irev == 1) {
        const int ni[6] = {n01.x, n01.y, n23.x, n23.y, n45.x, n45.y};
        const int nreact = c + 1, nslot = (c == 0) ? 5 : 6;
        double p;
        if (c == 0) p = gg[ni[0]];
        else { p = 1.0; for (int j = 0; j < nreact; ++j) p = A::mul(p, gg[ni[j]]); }
        for (int j = nreact; j < nslot; ++j) rpf = A::mul(rpf, gg[ni[j]]);
        rpf = A::div(rpf, p);
        if (heat) {   // d ln rpf / d T9 = sum over products - sum over reactants of d ln g / d T9
          double sr;
          if (c == 0) sr = dlg[ni[0]];
          else { sr = 0.0; for (int j = 0; j < nreact; ++j) sr = A::add(sr, dlg[ni[j]]); }
          for (int j = nreact; j < nslot; ++j) dlnrpf = A::add(dlnrpf, dlg[ni[j]]);
          dlnrpf = A::sub(dlnrpf, sr);
        }
      }
      const double eh = Mx::safe_exp(h);                // one copy of exp in the loop body (the branches below share it)
      if (c == 0) {
        if (iwk == 1) cs = A::mul(A::mul(ene, rpf), eh);
        else if (iwk == 2 || iwk == 3) cs = zm.rffn[net.iffn[k] - 1];
        else if (iwk == 7 || iwk == 8) cs = A::mul(rpf, 0.0);
        else cs = A::mul(rpf, eh);
      } else {
        const double rf = (c == 1) ? rhot : (c == 2) ? rho2 : rho3;
        if (iwk == 1) cs = A::mul(A::mul(A::mul(rf, ene), rpf), eh);
        else cs = A::mul(A::mul(rf, rpf), eh);
        if (c == 2 && cs < 1.0e-20) cs = 0.0;
      }
      if (heat) {   // xnet_integrate.F90:807-815 and the corresponding blocks of the other classes
        const double dh = A::add(A::mul(ascrn, dhscr), dlam);
        if (c == 0 && (iwk == 2 || iwk == 3)) dcs = A::mul(cs, zm.dlnrffn[net.iffn[k] - 1]);
        else if (c == 0 && (iwk == 7 || iwk == 8)) dcs = A::mul(cs, dlnrpf);
        else dcs = A::mul(cs, A::add(dh, dlnrpf));
      }
    }
    if (heat) zm.dcsect[k] = dcs;
    zm.csect[k] = cs;
    if (k >= net.cs_lo) zm.cs_sm[k - net.cs_lo] = cs;
  }
  }   // pass
  XB_DBG(zm, 6);
  __syncthreads();
  XB_DBG(zm, 7);
}

// --- term plans -------------------------------------------------------------------------------------
// product of one record:  (((a*csect[mu])*y[n0])*y[n1])*y[n2])*y[n3]   (absent factors are y[ny]=1)
#ifndef XB_PLAN_U
#define XB_PLAN_U 2
#endif
constexpr int kPlanU = XB_PLAN_U;   // records in flight per thread (host pads L to a multiple).  2 is the measured optimum of the
                                    // structured kernel (8: 2.63e6, 4: 2.68e6, 2: 2.71e6 zone-it/s on SN160): the smaller loop body wins
                                    // more through the instruction cache than the deeper batches win in memory-level parallelism

template <bool STRICT>
__device__ __forceinline__ double plan_mul(const double a, const double c, const uint2 nn, const double *y) {
  using A = Ar<STRICT>;
  double v = A::mul(a, c);
  v = A::mul(v, y[nn.x & 0xffffu]);
  v = A::mul(v, y[nn.x >> 16]);
  v = A::mul(v, y[nn.y & 0xffffu]);
  v = A::mul(v, y[nn.y >> 16]);
  return v;
}

// FAST: every thread reduces its chunk; `flush(ctl, acc)` is called at the end of each piece.  Loads of
// kPlanU records are issued together (two L2 round trips per batch instead of per record).
// rate of reaction mu: from the shared-memory cache when it is there, else from the L2-resident copy
__device__ __forceinline__ double cs_fetch(const double *cs, const double *cs_sm, const int cs_lo, const unsigned mu) {
  return ((int)mu >= cs_lo) ? cs_sm[(int)mu - cs_lo] : __ldcg(cs + mu);
}

template <class Flush>
__device__ __forceinline__ void plan_run_fast(const DevPlan &P, const double *cs, const double *cs_sm, const int cs_lo,
                                              const double *y, Flush flush) {
  const double *__restrict__ pa = P.a;
  const uint2 *__restrict__ pn = P.nn;
  const uint2 *__restrict__ pm = P.mc;
  const int L = P.L, nt = blockDim.x;
  int idx = threadIdx.x;
  double acc = 0.0;
  // software pipeline over batches of kPlanU records: the records of the next batch (coalesced loads from L2) are
  // in flight while the current batch gathers its rates and abundances from shared memory and multiplies
  double va[kPlanU]; uint2 na[kPlanU], ma[kPlanU];
#pragma unroll
  for (int u = 0; u < kPlanU; ++u) { va[u] = __ldg(pa + idx + u * nt); na[u] = __ldg(pn + idx + u * nt); ma[u] = __ldg(pm + idx + u * nt); }
  for (int k0 = 0; k0 < L; k0 += kPlanU) {
    double v[kPlanU], c[kPlanU], f0[kPlanU], f1[kPlanU], f2[kPlanU], f3[kPlanU];
    uint2 nn[kPlanU], mc[kPlanU];
#pragma unroll
    for (int u = 0; u < kPlanU; ++u) { v[u] = va[u]; nn[u] = na[u]; mc[u] = ma[u]; }
    idx += kPlanU * nt;
    if (k0 + kPlanU < L) {
#pragma unroll
      for (int u = 0; u < kPlanU; ++u) { va[u] = __ldg(pa + idx + u * nt); na[u] = __ldg(pn + idx + u * nt); ma[u] = __ldg(pm + idx + u * nt); }
    }
    // rate and abundance gathers, all in flight together; independent products; then the only sequential
    // part: accumulate and flush
#pragma unroll
    for (int u = 0; u < kPlanU; ++u) {
      c[u] = cs_fetch(cs, cs_sm, cs_lo, mc[u].x);
      f0[u] = y[nn[u].x & 0xffffu]; f1[u] = y[nn[u].x >> 16]; f2[u] = y[nn[u].y & 0xffffu]; f3[u] = y[nn[u].y >> 16];
    }
#pragma unroll
    for (int u = 0; u < kPlanU; ++u) v[u] = (((v[u] * c[u]) * f0[u]) * f1[u]) * f2[u] * f3[u];
#pragma unroll
    for (int u = 0; u < kPlanU; ++u) {
      acc += v[u];
      if (mc[u].y & 3u) { flush(mc[u].y, acc); acc = 0.0; }
    }
  }
}

// STRICT: products of all records to the per-CTA scratch in sorted order (thread t owns [t*L, (t+1)*L))
__device__ __forceinline__ void plan_products_strict(const DevPlan &P, const double *cs, const double *y, double *prod) {
  const double *__restrict__ pa = P.a;
  const uint2 *__restrict__ pn = P.nn;
  const uint2 *__restrict__ pm = P.mc;
  const int L = P.L, nt = blockDim.x, tid = threadIdx.x;
  double *out = prod + (size_t)tid * L;
  for (int k0 = 0; k0 < L; k0 += kPlanU) {
    double a[kPlanU], c[kPlanU];
    uint2 nn[kPlanU], mc[kPlanU];
#pragma unroll
    for (int u = 0; u < kPlanU; ++u) { const int idx = (k0 + u) * nt + tid; a[u] = __ldg(pa + idx); nn[u] = __ldg(pn + idx); mc[u] = __ldg(pm + idx); }
#pragma unroll
    for (int u = 0; u < kPlanU; ++u) c[u] = __ldcg(cs + mc[u].x);
#pragma unroll
    for (int u = 0; u < kPlanU; ++u) out[k0 + u] = plan_mul<true>(a[u], c[u], nn[u], y);
  }
}

// --- yderiv (xnet_integrate.F90:409-469): ydot_i = sum over the species' terms -----------------------
// FAST: chunked segmented reduction; segments cut by a chunk boundary are finished from per-chunk
// partials in chunk order (deterministic).  STRICT: products in parallel, then one thread per species
// adds them in the reference order, per reaction class, ((s1+s2)+s3)+s4.
template <bool STRICT>
__device__ __noinline__ void yderiv_cta() {
  XB_ZONE_PROLOGUE
  const DevPlan &P = net.py;
  const int tid = threadIdx.x, nt = blockDim.x, ny = net.ny;
  double *ydot = zm.ydot, *part = zm.part;
  if (STRICT) {
    plan_products_strict(P, zm.csect, zm.yt, zm.prod);
    __syncthreads();
    const double *prod = zm.prod;
#pragma unroll 1
    for (int i = tid; i < ny; i += nt) {
      double s[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double sc = 0.0;
        for (int q = P.seg_start[4 * i + c]; q < P.seg_start[4 * i + c + 1]; ++q) sc = __dadd_rn(sc, __ldcg(prod + q));
        s[c] = sc;
      }
      ydot[i] = __dadd_rn(__dadd_rn(__dadd_rn(s[0], s[1]), s[2]), s[3]);
    }
  } else {
    plan_run_fast(P, zm.csect, zm.cs_sm, net.cs_lo, zm.yt, [&](unsigned ctl, double acc) {
      if ((ctl & 3u) == 1u) ydot[ctl >> 3] = acc; else part[ctl >> 3] = acc;
    });
    __syncthreads();
#pragma unroll 1
    for (int m = tid; m < P.nmulti; m += nt) {
      double sm = 0.0;
      for (int q = P.multi_start[m]; q < P.multi_start[m + 1]; ++q) sm += part[q];
      ydot[P.multi_dest[m] >> 1] = sm;
    }
  }
  if (tid == 0) zm.s->ktot[3] += 1;
  __syncthreads();
  if (ctl.iheat > 0) {
    // t9dot = - sum_i mex_i ydot_i / cv (xnet_integrate.F90:471-487); (0 - t1) - t2 ... == -(t1 + t2 + ...) bit for bit
    using A = Ar<STRICT>;
    const double *mex = net.mex;
    const double cv = zm.s->cv;
    block_sum2<STRICT>(ny, [&](int k, double &ta, double &tb) { ta = A::div(A::mul(mex[k], ydot[k]), cv); tb = 0.0; },
                       part, zm.s->red);
    if (tid == 0) zm.s->t9dot = -zm.s->red[0];
    __syncthreads();
  }
}

// --- jacobian_build + jacobian_scale (xnet_jacobian_dense.F90:183-200, 303-341) ----------------------
// jac(i,j) = mult * sum_of_contributions(d ydot_i / d y_j) [+ diag on the diagonal], written once per
// non-zero element straight into the tile; no atomics, contributions added in the reference order.
template <bool STRICT, bool TILE>
__device__ __forceinline__ void jacobian_build_impl(const DevNet &net, const ZoneMem &zm, double *J) {
  if (TILE) XB_ASSUME_SHARED(J);
  const DevPlan &P = net.pj;
  const int tid = threadIdx.x, nt = blockDim.x, N = net.msize, ld = zm.ld;
  // self-heating: the unscaled dydotdy is assembled first (the T9 row is a linear functional of it), scaled last
  const bool heat = c_ctl.iheat > 0;
  const double diag = heat ? 0.0 : zm.s->rdt, mult = heat ? 1.0 : zm.s->mult;
  double *part = zm.part;
  int *pos = zm.pos;
  XB_ASSUME_SHARED(pos);
  // plan destinations are (column << 10 | row); row i of the matrix lives in slot pos[i]
  auto slot = [&](const unsigned d) { return (int)(d >> 10) * ld + pos[d & 1023u]; };
  if (STRICT) {   // strict arithmetic reproduces DGETRF's interchanges: no pivot prediction
#pragma unroll 1
    for (int k = tid; k < N; k += nt) { pos[k] = k; zm.orig[k] = k; }
  }
  XB_DBG_MARK(zm);
  {
    const int n2 = (N * ld) >> 1;
    double2 *J2 = reinterpret_cast<double2 *>(J);
    for (int e = tid; e < n2; e += nt) J2[e] = make_double2(0.0, 0.0);
    if (tid == 0 && ((N * ld) & 1)) J[N * ld - 1] = 0.0;
  }
  if (STRICT) {
    plan_products_strict(P, zm.csect, zm.yt, zm.prod);
    __syncthreads();
    const double *prod = zm.prod;
#pragma unroll 1
    for (int e = tid; e < P.nseg; e += nt) {
      double sc = 0.0;
      for (int q = P.seg_start[e]; q < P.seg_start[e + 1]; ++q) sc = __dadd_rn(sc, __ldcg(prod + q));
      const int d = P.seg_dest[e];
      double v = __dmul_rn(mult, sc);
      if (d & 1) v = __dadd_rn(v, diag);
      J[slot((unsigned)d >> 1)] = v;
    }
  } else {
    __syncthreads();  // zero fill complete before the scattered stores
    XB_DBG(zm, 0);
    plan_run_fast(P, zm.csect, zm.cs_sm, net.cs_lo, zm.yt, [&](unsigned ctl, double acc) {
      if ((ctl & 3u) == 1u) J[slot(ctl >> 3)] = fma(mult, acc, (ctl & 4u) ? diag : 0.0);
      else part[ctl >> 3] = acc;
    });
    XB_DBG(zm, 1);
    __syncthreads();
    XB_DBG(zm, 2);
#pragma unroll 1
    for (int m = tid; m < P.nmulti; m += nt) {
      double sm = 0.0;
      for (int q = P.multi_start[m]; q < P.multi_start[m + 1]; ++q) sm += part[q];
      const int d = P.multi_dest[m];
      J[slot((unsigned)d >> 1)] = fma(mult, sm, (d & 1) ? diag : 0.0);
    }
  }
  if (tid == 0) zm.s->ktot[2] += 1;
  __syncthreads();
  XB_DBG(zm, 3);
  if (heat) {
    using A = Ar<STRICT>;
    const int ny = net.ny;
    const DevPlan &Py = net.py;
    double *colT = J + (size_t)ny * ld;   // column ny: d ydot_i / d T9 = the ydot sums with dcsect/dT9 (xnet_jacobian_dense.F90:343-399)
    if (STRICT) {
      plan_products_strict(Py, zm.dcsect, zm.yt, zm.prod);
      __syncthreads();
      const double *prod = zm.prod;
#pragma unroll 1
      for (int i = tid; i < ny; i += nt) {
        double s4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double sc = 0.0;
          for (int q = Py.seg_start[4 * i + c]; q < Py.seg_start[4 * i + c + 1]; ++q) sc = __dadd_rn(sc, __ldcg(prod + q));
          s4[c] = sc;
        }
        colT[pos[i]] = __dadd_rn(__dadd_rn(__dadd_rn(s4[0], s4[1]), s4[2]), s4[3]);
      }
    } else {
      plan_run_fast(Py, zm.dcsect, zm.cs_sm, net.nrtot, zm.yt, [&](unsigned ctl, double acc) {
        if ((ctl & 3u) == 1u) colT[pos[ctl >> 3]] = acc; else part[ctl >> 3] = acc;
      });
      __syncthreads();
#pragma unroll 1
      for (int m = tid; m < Py.nmulti; m += nt) {
        double sm = 0.0;
        for (int q = Py.multi_start[m]; q < Py.multi_start[m + 1]; ++q) sm += part[q];
        colT[pos[Py.multi_dest[m] >> 1]] = sm;
      }
    }
    __syncthreads();
    // row ny: d t9dot / d(y_j, T9) = - sum_i mex_i dydotdy(j,i) / cv, summed over i in order
    {
      const double cv = zm.s->cv;
      const int rT = pos[ny];
      double *mexs = part;   // mass excesses staged in shared memory, in slot order: mexs[pos[i]] = mex[i]
      if (!STRICT) {
#pragma unroll 1
        for (int i = tid; i < N; i += nt) mexs[i] = 0.0;
        __syncthreads();
#pragma unroll 1
        for (int i = tid; i < ny; i += nt) mexs[pos[i]] = net.mex[i];
        __syncthreads();
      }
#pragma unroll 1
      for (int j = tid; j < N; j += nt) {
        const double *col = J + (size_t)j * ld;
        double sdot = 0.0;
        if (STRICT) {
          const double *mex = net.mex;
          for (int i = 0; i < ny; ++i) sdot = A::sub(sdot, A::div(A::mul(mex[i], col[pos[i]]), cv));
        } else {   // one pass over the slots (the T9 slot still holds 0), one division
          double s0 = 0.0, s1 = 0.0;
          int r = 0;
          for (; r + 1 < N; r += 2) { s0 = fma(mexs[r], col[r], s0); s1 = fma(mexs[r + 1], col[r + 1], s1); }
          if (r < N) s0 = fma(mexs[r], col[r], s0);
          sdot = -(s0 + s1) / cv;
        }
        J[(size_t)j * ld + rT] = sdot;
      }
    }
    __syncthreads();
    // jacobian_scale: jac = mult * dydotdy^T + diag on the diagonal (a warp per column, no index division)
    {
      const double rdt = zm.s->rdt, m = zm.s->mult;
      const int lane = tid & 31, wq = tid >> 5, nw = nt >> 5;
      for (int j = wq; j < N; j += nw) {
        double *col = J + (size_t)j * ld;
        const int dslot = pos[j];
        for (int r = lane; r < N; r += 32) {
          double v = A::mul(m, col[r]);
          if (r == dslot) v = A::add(v, rdt);
          col[r] = v;
        }
      }
    }
    __syncthreads();
  }
}
#ifdef XB_SLU
// Structured variant: the same term plan, destinations = addresses in the skyline store (static reordering, no slots).
template <bool STRICT>
__device__ __noinline__ void jacobian_build_cta() {
  static_assert(!STRICT, "the structured LU is a fast-arithmetic path");
  XB_ZONE_PROLOGUE
  const DevSlu &D = net.slu;
  const DevPlan &P = net.pj;
  const int tid = threadIdx.x, nt = blockDim.x;
  double *S = zm.jac, *part = zm.part;
  XB_ASSUME_SHARED(S);
  const bool heat = ctl.iheat > 0;
  const double diag = heat ? 0.0 : zm.s->rdt, mult = heat ? 1.0 : zm.s->mult;
  {
    const int n2 = D.nstore >> 1;
    double2 *S2 = reinterpret_cast<double2 *>(S);
    for (int e = tid; e < n2; e += nt) S2[e] = make_double2(0.0, 0.0);
    if (tid == 0 && (D.nstore & 1)) S[D.nstore - 1] = 0.0;
  }
  __syncthreads();
  plan_run_fast(P, zm.csect, zm.cs_sm, net.cs_lo, zm.yt, [&](unsigned ctl_, double acc) {
    if ((ctl_ & 3u) == 1u) S[ctl_ >> 3] = fma(mult, acc, (ctl_ & 4u) ? diag : 0.0);
    else part[ctl_ >> 3] = acc;
  });
  __syncthreads();
#pragma unroll 1
  for (int m = tid; m < P.nmulti; m += nt) {
    double sm = 0.0;
    for (int q = P.multi_start[m]; q < P.multi_start[m + 1]; ++q) sm += part[q];
    const int d = P.multi_dest[m];
    S[d >> 1] = fma(mult, sm, (d & 1) ? diag : 0.0);
  }
  if (tid == 0) zm.s->ktot[2] += 1;
  __syncthreads();
  if (heat) {
    const SluIdx I = slu_idx();
    const int n = I.n, ny = net.ny, nh = I.nh;
    const DevPlan &Py = net.py;
    const int cbT = slu_cb(I, n - 1);       // column n-1 (a border column: every heavy row stored): d ydot_i / d T9 (xnet_jacobian_dense.F90:343-399)
    const int *__restrict__ spos = D.pos;
    auto t9col = [&](const int eq) { const int pi = spos[eq]; return (pi >= nh) ? I.bo + (n - 1) * I.bs + (pi - nh) : cbT + pi; };
    plan_run_fast(Py, zm.dcsect, zm.cs_sm, net.nrtot, zm.yt, [&](unsigned ctl_, double acc) {
      if ((ctl_ & 3u) == 1u) S[t9col(ctl_ >> 3)] = acc; else part[ctl_ >> 3] = acc;
    });
    __syncthreads();
#pragma unroll 1
    for (int m = tid; m < Py.nmulti; m += nt) {
      double sm = 0.0;
      for (int q = Py.multi_start[m]; q < Py.multi_start[m + 1]; ++q) sm += part[q];
      S[t9col(Py.multi_dest[m] >> 1)] = sm;
    }
    __syncthreads();
    // row n-1: d t9dot / d(y_j, T9) = - sum_i mex_i dydotdy(i, j) / cv over the stored rows of column j
    double *mexp = part;                    // mass excesses in position order, 0 for the T9 row
#pragma unroll 1
    for (int i = tid; i < n; i += nt) mexp[i] = 0.0;
    __syncthreads();
#pragma unroll 1
    for (int i = tid; i < ny; i += nt) mexp[spos[i]] = net.mex[i];
    __syncthreads();
    const double cv = zm.s->cv;
    const int nbq = I.nb;
#pragma unroll 1
    for (int pj = tid; pj < n; pj += nt) {
      double s0 = 0.0;
      const int cbj = slu_cb(I, pj);
      const int r1 = slu_cbot(I, pj);
      for (int r = slu_cf(I, pj); r <= r1; ++r) s0 = fma(mexp[r], S[cbj + r], s0);
      for (int r = 0; r < nbq - 1; ++r) s0 = fma(mexp[nh + r], S[I.bo + pj * I.bs + r], s0);
      S[I.bo + pj * I.bs + nbq - 1] = -s0 / cv;
    }
    __syncthreads();
    // jacobian_scale: jac = mult * dydotdy^T + diag on the diagonal
    const double rdt = zm.s->rdt, m = zm.s->mult;
#pragma unroll 1
    for (int e = tid; e < D.nstore; e += nt) S[e] *= m;
    __syncthreads();
#pragma unroll 1
    for (int p = tid; p < n; p += nt) S[slu_addr(I, p, p)] += rdt;
    __syncthreads();
  }
}
#else
template <bool STRICT>
__device__ __noinline__ void jacobian_build_cta() {
  XB_ZONE_PROLOGUE
  if (net.tile) jacobian_build_impl<STRICT, true>(net, zm, zm.jac);
  else jacobian_build_impl<STRICT, false>(net, zm, zm.jac);
}
#endif

// with_rhs: zm.rhs rides through the factorisation as column N of the tile (forward solve for free)
template <bool STRICT>
__device__ __forceinline__ void lu_cta(const ZoneMem &zm, const int N, const bool with_rhs, const bool invert) {
  if (zm.tile) {
    // zm.rhs IS column N of the tile
    lu_blocked<STRICT, true, 1>(zm.jac, zm.ld, N, with_rhs ? 1 : 0, zm.indx, &zm.s->info, zm.prow, zm.rdiag,
                                zm.s->ph + PH_LU_PANEL, zm.orig, zm.pos, invert);
  } else if (zm.lu_s) {
    lu_two_level<STRICT>(zm.jac, zm.ld, N, (with_rhs && zm.s->rhs_in_jac) ? 1 : 0, zm.indx, &zm.s->info, zm.prow, zm.rdiag,
                         zm.s->ph + PH_LU_PANEL, zm.orig, zm.pos, invert, zm.lu_s, zm.lds2, zm.lu_u);
  } else {
    lu_blocked<STRICT, false, 2>(zm.jac, zm.ld, N, (with_rhs && zm.s->rhs_in_jac) ? 1 : 0, zm.indx, &zm.s->info, zm.prow,
                                 zm.rdiag, zm.s->ph + PH_LU_PANEL, zm.orig, zm.pos, invert);
  }
}
template <bool STRICT>
// the right-hand side is already in slot order (rhs[pos[i]] = b_i), so no interchanges are applied here
__device__ __forceinline__ void bksub_cta(const ZoneMem &zm, const int N, const bool fwd_done) {
  if (zm.tile) bksub_block<STRICT, true, 1>(zm.jac, zm.ld, N, zm.indx, zm.rhs, zm.rdiag, zm.prow, fwd_done, false, !STRICT);
  else bksub_block<STRICT, false, 2>(zm.jac, zm.ld, N, zm.indx, zm.rhs, zm.rdiag, zm.prow, fwd_done, false, !STRICT);
}

// --- step_be: Newton-Raphson loop (xnet_integrate_be.F90:233-482) -------------------------------
template <bool STRICT>
__device__ __noinline__ void step_be_cta() {
  using A = Ar<STRICT>;
  XB_ZONE_PROLOGUE
  ZoneScalars *s = zm.s;
  const int tid = threadIdx.x, nt = blockDim.x, ny = net.ny, N = net.msize;
  if (s->inr != 0) return;
  {
    const double *aa = net.aa, *y = zm.y;
    block_sum2<STRICT>(ny, [&](int k, double &ta, double &tb) { ta = A::mul(aa[k], y[k]); tb = 0.0; }, zm.part, s->red);
    if (tid == 0) {
      s->xtot_init = A::sub(A::add(s->red[0], s->xext), 1.0);
      s->rdt = A::div(1.0, s->tdel);
      s->mult = -1.0;
    }
    __syncthreads();
  }
  for (int kit = 1; kit <= ctl.kitmx; ++kit) {
    const bool rebuild = ((kit - 1) % ctl.ijac) == 0;
    const bool eval_rates = (ctl.iheat > 0) ? true : (kit == 1 && s->nh > 1);
    XB_PHASE_BEGIN(zm);
    if (eval_rates) cross_sect_cta<STRICT>();
    XB_PHASE_END(zm, PH_RATES);
    yderiv_cta<STRICT>();
    XB_PHASE_END(zm, PH_YDERIV);
    // right-hand side (xnet_integrate_be.F90:327-349); it does not depend on the Jacobian
    const double rdt = s->rdt;
    // in slot order: row k of the system sits in slot pos[k] of the Jacobian.  The slots are fixed by the
    // previous factorisation (or reset by the strict build), so the Jacobian is assembled first.
    if (rebuild) {
      jacobian_build_cta<STRICT>();
      XB_PHASE_END(zm, PH_JAC);
    }
#ifdef XB_SLU
    // structured LU: equation k sits at position spos[k] of the reordered system (T9 last); the solution comes back in
    // the same order
    const int *__restrict__ spos = net.slu.pos;
    auto fill_rhs = [&]() {
      if (kit > 1) {
#pragma unroll 1
        for (int k = tid; k < ny; k += nt) zm.rhs[spos[k]] = A::mad(A::sub(zm.y[k], zm.yt[k]), rdt, zm.ydot[k]);
        if (ctl.iheat > 0 && tid == 0) zm.rhs[N - 1] = A::mad(A::sub(s->t9, s->t9t), rdt, s->t9dot);
      } else {
#pragma unroll 1
        for (int k = tid; k < ny; k += nt) zm.rhs[spos[k]] = zm.ydot[k];
        if (ctl.iheat > 0 && tid == 0) zm.rhs[N - 1] = s->t9dot;
      }
    };
    fill_rhs();
    if (tid == 0 && rebuild) { s->slu_fail = 0; s->slu_fail_k = 0x7fffffff; s->nlu_zone += 1; }
    __syncthreads();
    if (rebuild) {
      slu_factor(zm.jac);
      __syncthreads();
      const bool forced = net.slu.force_fb > 0 && (s->nlu_zone % net.slu.force_fb) == 0;
      if (s->slu_fail != 0 || forced) {
        // a column failed the threshold test: this factorisation is redone by the dense partial-pivoting LU
        jacobian_build_cta<STRICT>();
        fill_rhs();
        if (tid == 0) {
          s->ktot[2] -= 1; s->lu_dense = 1; s->ph[PH_SPARE] += 1;
          // where the fallbacks come from: sums of the first violated column, the zone's factorisation count and the
          // Newton iteration (divide by the fallback count)
          s->ph[PH_DBG0] += (s->slu_fail_k < N) ? s->slu_fail_k : N; s->ph[PH_DBG0 + 1] += s->nlu_zone; s->ph[PH_DBG0 + 2] += kit;
          const long long now_ = clock64(); s->ph[PH_LU_PANEL] += now_ - s->tmark;   // the failed structured attempt
        }
        __syncthreads();
        slu_dense_factor(zm.jac, zm.rhs, s->g_jac, reinterpret_cast<int *>(s->g_jac + (size_t)(N + 1) * N), zm.part, &s->info);
        if (tid == 0) { const long long now_ = clock64(); s->ph[PH_LU_SWAP] += now_ - s->tmark; }   // fallback cycles (incl. the failed attempt)
      } else if (tid == 0) s->lu_dense = 0;
      __syncthreads();
      XB_PHASE_END(zm, PH_LU);
    }
    if (s->lu_dense) slu_dense_solve(s->g_jac, reinterpret_cast<const int *>(s->g_jac + (size_t)(N + 1) * N), zm.rhs, rebuild);
    else slu_solve(zm.jac, zm.rhs, s->slu_ipb, rebuild);
    __syncthreads();
    XB_PHASE_END(zm, PH_BKSUB);
#else
    if (kit > 1) {
#pragma unroll 1
      for (int k = tid; k < ny; k += nt) zm.rhs[zm.pos[k]] = A::mad(A::sub(zm.y[k], zm.yt[k]), rdt, zm.ydot[k]);
      if (ctl.iheat > 0 && tid == 0) zm.rhs[zm.pos[ny]] = A::mad(A::sub(s->t9, s->t9t), rdt, s->t9dot);
    } else {
#pragma unroll 1
      for (int k = tid; k < ny; k += nt) zm.rhs[zm.pos[k]] = zm.ydot[k];
      if (ctl.iheat > 0 && tid == 0) zm.rhs[zm.pos[ny]] = s->t9dot;
    }
    __syncthreads();
    if (rebuild) {
      lu_cta<STRICT>(zm, N, true, true);
      XB_PHASE_END(zm, PH_LU);
    }
    bksub_cta<STRICT>(zm, N, rebuild && (zm.tile || s->rhs_in_jac));
    XB_PHASE_END(zm, PH_BKSUB);
#endif
    {
      const double *aa = net.aa;
      const double ymin = ctl.ymin;
      double *yt = zm.yt;
      const double *rhs = zm.rhs;
      block_sum2<STRICT>(ny, [&](int k, double &ta, double &tb) {
#ifdef XB_SLU
        const double dy = rhs[spos[k]];
#else
        const double dy = rhs[k];
#endif
        const double ytmp = A::add(yt[k], dy);
        double rel, ynew;
        if (ytmp < ymin) { ynew = 0.0; rel = 0.0; }
        else { ynew = ytmp; rel = fabs(A::div(dy, ytmp)); }
        yt[k] = ynew;
        ta = rel;
        tb = A::mul(aa[k], ynew);
      }, zm.part, s->red);
    }
    if (tid == 0) {
      const double xtot = A::sub(A::add(s->red[1], s->xext), 1.0);
      const double testm = A::sub(xtot, s->xtot_init);
      const double testn = (ctl.iconvc != 0) ? s->red[0] : testm;
      const double toln = (ctl.iconvc != 0) ? ctl.tolc : ctl.tolm;
      s->relt9 = 0.0;
      if (ctl.iheat > 0) {   // temperature update and its convergence measure (xnet_integrate_be.F90:402-416)
        s->dt9 = zm.rhs[ny];   // T9 is equation ny in both orderings (the structured LU keeps it last)
        s->t9t = A::add(s->t9t, s->dt9);
        s->relt9 = (fabs(s->t9t) > 2.2250738585072014e-308) ? fabs(A::div(s->dt9, s->t9t)) : 0.0;
      }
      s->flag = (fabs(testn) <= toln && s->relt9 < ctl.tolt9) ? 1 : 0;
      if (s->flag) s->inr = kit;
    }
    __syncthreads();
    XB_PHASE_END(zm, PH_OTHER);
    if (s->flag) break;
  }
}

// --- timestep (xnet_integrate.F90:13-268), executed by thread 0 after the optional rate pass ---
template <bool STRICT>
__device__ __noinline__ void timestep_cta() {
  using A = Ar<STRICT>;
  XB_ZONE_PROLOGUE
  ZoneScalars *s = zm.s;
  const int ny = net.ny;
  if (threadIdx.x == 0) {
    s->tdel_old = s->tdel;
    s->flag = (fabs(s->tdel_old) < 2.2250738585072014e-308) ? 1 : 0;
  }
  __syncthreads();
  const bool mask_init = s->flag != 0;
  if (mask_init) {
    XB_PHASE_BEGIN(zm);
    cross_sect_cta<STRICT>();
    XB_PHASE_END(zm, PH_RATES);
    yderiv_cta<STRICT>();
    XB_PHASE_END(zm, PH_YDERIV);
  }
  XB_PHASE_BEGIN(zm);
  const bool fwd = s->tdel_old >= 0.0;
  if (fwd) {
    // controlling species: first maximum of |dY/Y| per unit time (xnet_integrate.F90:118-139)
    const double *y = zm.y, *yo = zm.yo, *ydot = zm.ydot;
    const double ymin = ctl.ymin, yacc = ctl.yacc, tdo = s->tdel_old;
    block_argmax_first(ny, [&](int k) {
      double r = 0.0;
      const double yk = y[k];
      if (yk > ymin) {
        const double yfloor = fmax(yk, yacc);
        if (mask_init) r = fabs(A::div(ydot[k], yfloor));
        else r = A::div(fabs(A::div(A::sub(yk, yo[k]), yfloor)), tdo);
      }
      return r;
    }, zm.part, s->red, &s->ints);
  }
  if (threadIdx.x == 0) {
    const double changeth = 0.1;
    double changey, changet9, tdel_init, tdel_dy, tdel_dt9;
    const double tdel_stop = A::sub(s->tstop, s->t);
    if (mask_init) {
      const double changest = (s->tdelstart < 0.0) ? 0.5 : 0.01;
      changey = A::mul(changest, fmin(0.1, ctl.changemx));
      changet9 = A::mul(changest, fmin(0.01, ctl.changemxt));
      tdel_init = fmin(A::mul(changest, fabs(s->tdelstart)), tdel_stop);
      s->tdel_next = (s->nh > 2) ? A::mul(1.0e-4, tdel_stop) : tdel_stop;
    } else { changey = ctl.changemx; changet9 = ctl.changemxt; tdel_init = 0.0; }
    if (fwd) {
      const double vmax = s->red[0];
      s->ints += 1;
      tdel_dy = (fabs(vmax) > 0.0) ? A::div(changey, vmax) : s->tdel_next;
      double rtau_t9 = 0.0;   // temperature time scale (xnet_integrate.F90:141-160); zero without self-heating
      if (ctl.iheat > 0)
        rtau_t9 = mask_init ? fabs(A::div(s->t9dot, s->t9)) : A::div(fabs(A::div(A::sub(s->t9, s->t9o), s->t9)), s->tdel_old);
      tdel_dt9 = (rtau_t9 > 0.0) ? A::div(changet9, rtau_t9) : s->tdel_next;
      s->tdel = fmax(tdel_init, fmin(fmin(tdel_stop, s->tdel_next), fmin(tdel_dy, tdel_dt9)));
    } else s->tdel = -s->tdel_old;
    s->tt = fmin(A::add(s->t, s->tdel), s->tstop);
    if (ctl.iheat == 0 && s->nh > 1) {
      t9rhofind<STRICT>(s, s->tt, &s->ntt, &s->t9t, &s->rhot, s->nt);
      if (s->ntt - 1 > s->nt) {
        const double sc = s->rho_scale; const bool hs = s->has_scale != 0;
        for (int j = s->nt; j <= s->ntt - 1; ++j) {
          const double t9j = s->t9h[j - 1];
          const double rhoj = hs ? A::mul(s->rhoh[j - 1], sc) : s->rhoh[j - 1];
          if ((t9j > s->t9 && t9j > s->t9t) || (t9j < s->t9 && t9j < s->t9t) ||
              (rhoj > s->rho && rhoj > s->rhot) || (rhoj < s->rho && rhoj < s->rhot)) {
            s->tdel = A::sub(s->th[j - 1], s->t);
            break;
          }
        }
        s->tt = fmin(A::add(s->t, s->tdel), s->tstop);
      }
      double dtherm = 0.0;
      for (int i = 1; i <= 10; ++i) {
        t9rhofind<STRICT>(s, s->tt, &s->ntt, &s->t9t, &s->rhot, s->nt);
        if (s->t9 > 0.0)
          dtherm = A::add(A::div(A::mul(10.0, fabs(A::sub(s->t9t, s->t9))), s->t9),
                          A::div(fabs(A::sub(s->rhot, s->rho)), s->rho));
        if (dtherm < changeth) break;
        s->tdel = A::mul(0.5, s->tdel);
        s->tt = fmin(A::add(s->t, s->tdel), s->tstop);
      }
    }
  }
  __syncthreads();
  XB_PHASE_END(zm, PH_OTHER);
}

// --- solve_be: retry ladder + accept (xnet_integrate_be.F90:32-231) -----------------------------
template <bool STRICT>
__device__ __noinline__ void solve_be_cta() {
  using A = Ar<STRICT>;
  XB_ZONE_PROLOGUE
  ZoneScalars *s = zm.s;
  const int tid = threadIdx.x, nt = blockDim.x, ny = net.ny;
  const int ktsmx = 10;
  if (s->its != 0) return;
  if (tid == 0) s->inr = 0;
  __syncthreads();
  int mykts = 0;
  for (int kts = 1; kts <= ktsmx; ++kts) {
    step_be_cta<STRICT>();
    const int inr = s->inr;
    __syncthreads();
    if (inr == 0) {
      mykts = kts + 1;
#pragma unroll 1
      for (int k = tid; k < ny; k += nt) zm.yt[k] = zm.y[k];
      if (tid == 0) {
        s->tdel = A::div(s->tdel, ctl.tdel_maxmult);
        s->tt = A::add(s->t, s->tdel);
        s->yet = s->ye;
        s->kmon[1] = ctl.kitmx + 1; s->ktot[1] += ctl.kitmx + 1;
        t9rhofind<STRICT>(s, s->tt, &s->ntt, &s->t9t, &s->rhot, s->nt);
        if (ctl.iheat > 0) s->t9t = s->t9;
        if (kts == ktsmx - 1) { s->tdel = 0.0; s->tdelstart = 0.0; }
      }
      __syncthreads();
      if (kts == ktsmx - 1) timestep_cta<STRICT>();
    } else {
      mykts = kts;
      if (tid == 0) { s->kmon[1] = inr; s->ktot[1] += inr; }
      break;
    }
  }
  __syncthreads();
  const int inr = s->inr;
  if (inr > 0) {
#pragma unroll 1
    for (int k = tid; k < ny; k += nt) { zm.yo[k] = zm.y[k]; zm.y[k] = zm.yt[k]; }
  }
  if (tid == 0) {
    s->kmon[0] = mykts; s->ktot[0] += mykts;
    if (inr > 0) {
      s->tdel_next = A::mul(s->tdel, ctl.tdel_maxmult);
      s->nto = s->nt; s->nt = s->ntt;
      s->to = s->t; s->t = s->tt;
      s->t9o = s->t9; s->t9 = s->t9t;
      s->rhoo = s->rho; s->rho = s->rhot;
      s->yeo = s->ye; s->ye = s->yet;
    } else s->its = 1;
  }
  __syncthreads();
}

// benuc (xnet_data.F90:157-175) and enudot (:990-1020); all threads call
template <bool STRICT>
__device__ __forceinline__ void benuc_cta(const ZoneMem &zm, const double *y, double *enm) {
  using A = Ar<STRICT>;
  const DevNet &net = c_net;
  const double *zz = net.zz, *be = net.be;
  block_sum2<STRICT>(net.ny, [&](int k, double &ta, double &tb) { ta = A::mul(y[k], zz[k]); tb = A::mul(y[k], be[k]); },
                     zm.part, zm.s->red);
  if (threadIdx.x == 0) {
    const double ztot = zm.s->red[0], btot = zm.s->red[1];
    const double e = A::sub(A::add(A::mul(net.mex_p, ztot), A::mul(net.mex_n, A::sub(1.0, ztot))), btot);
    *enm = A::mul(A::mul(1.6021766208e-06, 6.022140857e+23), e);
  }
  __syncthreads();
}
// sum over the FFN reactions in reaction order (all other reactions contribute +0 in the reference loop)
template <bool STRICT>
__device__ __forceinline__ void enudot_cta(const ZoneMem &zm, const double *y, double *sqnu) {
  using A = Ar<STRICT>;
  const DevNet &net = c_net;
  if (net.nffn > 0) {
    const int *n1 = net.ffr_n1, *ff = net.ffr_iffn;
    const double *qk = zm.qkffn;
    block_sum2<STRICT>(net.nffr, [&](int k, double &ta, double &tb) { ta = A::mul(y[n1[k]], qk[ff[k]]); tb = 0.0; },
                       zm.part, zm.s->red);
    if (threadIdx.x == 0) *sqnu = A::mul(A::mul(1.6021766208e-06, 6.022140857e+23), zm.s->red[0]);
    __syncthreads();
  }
}

// flux (xnet_flux.F90:104-131): one thread per matched reaction pair; the pair's reactions are summed in reaction
// order (classes 1..4, index ascending), each term dcflx * sign(csect * prod yt, ifl) with the reference's products.
// Rates are those of the last cross_sect call, abundances zm.yt (= the accepted y when called from ts_output).
template <bool STRICT>
__device__ __forceinline__ void flux_cta(const ZoneMem &zm, double *__restrict__ flx) {
  using A = Ar<STRICT>;
  const DevNet &net = c_net;
  const double *yt = zm.yt;
  for (int m = threadIdx.x; m < net.mflx; m += blockDim.x) {
    double acc = 0.0;
    for (int e = net.fl_start[m]; e < net.fl_start[m + 1]; ++e) {
      const int k = net.fl_k[e];
      const int *ni = net.ni + (size_t)k * 6;
      const double cs = cs_fetch(zm.csect, zm.cs_sm, net.cs_lo, (unsigned)k);
      double flt;
      if (k < net.roff[1]) flt = A::mul(cs, yt[ni[0] - 1]);
      else if (k < net.roff[2]) flt = A::mul(A::mul(cs, yt[ni[0] - 1]), yt[ni[1] - 1]);
      else if (k < net.roff[3]) flt = A::mul(cs, A::mul(A::mul(yt[ni[0] - 1], yt[ni[1] - 1]), yt[ni[2] - 1]));
      else flt = A::mul(cs, A::mul(A::mul(A::mul(yt[ni[0] - 1], yt[ni[1] - 1]), yt[ni[2] - 1]), yt[ni[3] - 1]));
      const double dc = net.fl_dc[e];
      // dcflx * sign(flt, ifl): the magnitude of the term, signed by the direction of the reaction
      acc = A::add(acc, A::mul(fabs(dc), dc < 0.0 ? -fabs(flt) : fabs(flt)));
    }
    flx[m] = acc;
  }
}

template <bool STRICT>
__device__ inline void hist_put(const DevOut &out, int iz, int ny, int kstep, const ZoneMem &zm) {
  // all threads call; record = one ts_output record (xnet_output.F90:78-79), flux block when requested
  if (out.hist_cap <= 0) return;
  const int n = out.hist_n[iz];
  if (n < out.hist_cap) {
    const size_t o = (size_t)iz * out.hist_cap + n;
    for (int k = threadIdx.x; k < ny; k += blockDim.x) out.hist_y[o * ny + k] = zm.y[k];
    if (out.hist_flx && c_net.mflx > 0) {   // ts_output: flux() for kstep > 0, zeros in the initial record
      double *f = out.hist_flx + o * (size_t)c_net.mflx;
      if (kstep > 0) flux_cta<STRICT>(zm, f);
      else for (int m = threadIdx.x; m < c_net.mflx; m += blockDim.x) f[m] = 0.0;
    }
    if (threadIdx.x == 0) {
      const ZoneScalars *s = zm.s;
      out.hist_kstep[o] = kstep; out.hist_kmon[2 * o] = s->kmon[0]; out.hist_kmon[2 * o + 1] = s->kmon[1];
      out.hist_t[o] = s->t; out.hist_t9[o] = s->t9; out.hist_rho[o] = s->rho; out.hist_tdel[o] = s->tdel;
      out.hist_edot[o] = s->edot; out.hist_sqnu[o] = s->sqnu;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out.hist_n[iz] = n + 1;
}

}  // namespace xb
#ifndef XB_SLU
#include "xb_bdf.cuh"
#endif
namespace xb {

// --- full_net for one zone (xnet_evolve.F90:19-269; zone load net.F90:246-272) ------------------
template <bool STRICT>
__device__ __forceinline__ void full_net_cta(const DevZones &zn, const DevOut &out, int iz) {
  using A = Ar<STRICT>;
  XB_ZONE_PROLOGUE
  ZoneScalars *s = zm.s;
  const int tid = threadIdx.x, nt = blockDim.x, ny = net.ny;
  const int fx = zn.zfix[iz];
  const bool restart = zn.rs_t != nullptr;
  int kstep0 = 0;
  if (tid == 0) {
    s->nh = zn.nh[fx]; s->th = zn.th + zn.off[fx]; s->t9h = zn.t9h + zn.off[fx]; s->rhoh = zn.rhoh + zn.off[fx];
    s->has_scale = zn.rho_scale != nullptr; s->rho_scale = s->has_scale ? zn.rho_scale[iz] : 1.0;
    s->tstart = zn.tstart[fx]; s->tstop = zn.tstop[fx]; s->tdelstart = zn.tdelstart[fx];
    s->xext = zn.xext[fx]; s->aext = zn.aext[fx]; s->zext = zn.zext[fx]; s->zexti = zn.zexti[fx];
    s->ye = zn.yestart[fx];
    s->cv = s->etae = s->detaedt9 = s->t9dot = 0.0;
    for (int k = 0; k < 5; ++k) { s->kmon[k] = 0; s->ktot[k] = 0; }
    for (int k = 0; k < PH_COUNT; ++k) s->ph[k] = 0;
    s->ints = 0; s->its = 0; s->edot = 0.0; s->sqnu = 0.0;
    if (ctl.isolv == 3) reinterpret_cast<BdfScalars *>(s->g_bdf)->started = 0;
    if (out.hist_cap > 0) out.hist_n[iz] = 0;
    if (restart) {
      s->t = zn.rs_t[iz]; s->tdel = zn.rs_tdel[iz]; s->tdel_next = A::mul(s->tdel, ctl.tdel_maxmult);
      s->t9 = zn.rs_t9[iz]; s->rho = zn.rs_rho[iz]; s->t9o = zn.rs_t9o[iz];
    } else {
      s->tdel = 0.0; s->nt = zn.nstart[fx]; s->t = s->tstart; s->t9 = zn.t9start[fx];
      s->rho = s->has_scale ? A::mul(zn.rhostart[fx], s->rho_scale) : zn.rhostart[fx];
      s->tdel_next = s->tdel; s->t9o = s->t9;
    }
  }
  if (restart) {
    kstep0 = zn.rs_kstep0[iz];
#pragma unroll 1
    for (int k = tid; k < ny; k += nt) {
      zm.y[k] = zn.rs_y[(size_t)iz * ny + k]; zm.yo[k] = zn.rs_yo[(size_t)iz * ny + k]; zm.yt[k] = zm.y[k];
      zm.ydot[k] = 0.0;
    }
  } else {
#pragma unroll 1
    for (int k = tid; k < ny; k += nt) {
      const double v = zn.ystart[(size_t)fx * ny + k];
      zm.y[k] = v; zm.yo[k] = v; zm.yt[k] = v; zm.ydot[k] = 0.0;
    }
  }
  for (int k = tid; k <= net.nffn; k += nt) zm.qkffn[k] = 0.0;
#ifndef XB_SLU
#pragma unroll 1
  for (int k = tid; k < net.msize; k += nt) { zm.orig[k] = k; zm.pos[k] = k; }   // no pivot history yet
#else
  if (tid == 0) { s->lu_dense = 0; s->nlu_zone = 0; s->slu_fail = 0; }
#endif
  __syncthreads();
  if (tid == 0) {
    if (restart) { double t9x, rhox; t9rhofind<STRICT>(s, s->t, &s->nt, &t9x, &rhox); }
    s->tdel_old = s->tdel; s->nto = s->nt; s->ntt = s->nt; s->to = s->t; s->tt = s->t;
    s->t9t = s->t9; s->rhoo = s->rho; s->rhot = s->rho; s->yeo = s->ye; s->yet = s->ye;
    s->tmark = clock64();
    s->ph[PH_TOTAL] = -s->tmark;
  }
  __syncthreads();
  benuc_cta<STRICT>(zm, zm.y, &s->enm);
  if (tid == 0) s->enold = restart ? s->enm : 0.0;
  __syncthreads();
  // constant conditions without self-heating: the rates are evaluated once, in the first timestep call (tdel = 0),
  // and never by step_be; a restarted zone has tdel != 0, so the rates of its state are evaluated here
  if (restart && s->nh == 1 && ctl.iheat == 0) cross_sect_cta<STRICT>();
  if (!restart) hist_put<STRICT>(out, iz, ny, 0, zm);
  int mykstep = 0;
  bool lzsolve = true;
  for (int kstep = kstep0 + 1; kstep <= kstep0 + ctl.kstmx; ++kstep) {
    timestep_cta<STRICT>();
#ifndef XB_SLU
    if (ctl.isolv == 3) solve_bdf_cta<STRICT>(); else solve_be_cta<STRICT>();
#else
    solve_be_cta<STRICT>();
#endif
    bool lzoutput = true;
    const int its = s->its;
    __syncthreads();
    if (its == 0) {
      if (tid == 0) s->enold = s->enm;
      if (s->t >= s->tstop) { mykstep = kstep; lzsolve = false; if (tid == 0) s->its = -1; }
    } else { lzsolve = false; lzoutput = false; if (tid == 0 && its == 1) s->its = 2; }
    __syncthreads();
    if (lzoutput) {
      benuc_cta<STRICT>(zm, zm.yt, &s->enm);
      enudot_cta<STRICT>(zm, zm.yt, &s->sqnu);
      if (tid == 0) s->edot = -A::div(A::sub(s->enm, s->enold), s->tdel);
      __syncthreads();
      hist_put<STRICT>(out, iz, ny, kstep, zm);
    }
    if (!lzsolve) break;
  }
  __syncthreads();
#pragma unroll 1
  for (int k = tid; k < ny; k += nt) out.y[(size_t)iz * ny + k] = zm.y[k];
  if (tid == 0) {
    out.t[iz] = s->t; out.t9[iz] = s->t9; out.rho[iz] = s->rho; out.tdel[iz] = s->tdel;
    out.kstep[iz] = mykstep; out.status[iz] = s->its;
    for (int k = 0; k < 5; ++k) out.ktot[(size_t)iz * 5 + k] = s->ktot[k];
    s->ph[PH_TOTAL] += clock64();
  }
  __syncthreads();
}

}  // namespace xb
