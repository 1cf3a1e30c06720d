// Variable-order BDF integrator on the device: solve_bdf of the XNet tree,
// source/xnet_integrate_bdf.F90:155-310, with bdf_reset :490, bdf_adjust :595, bdf_predict :631, bdf_update :688,
// bdf_step :833, bdf_check :1122, bdf_correct :1326, bdf_eta :1397, bdf_adjust_order :1844, bdf_rescale :1939,
// bdf_restore :1980, error_weights :2048, normw :2087.  (check_stability = .false. in the reference: bdf_stab and
// bdf_stab_detect are dead code there and are not built here.)
//
// One CTA = one zone, as everywhere in the fused path: the zone masks of the reference (bdf_active, refactor, rebuild,
// iterate, restore, rescale, retry_ts) are per-zone conditions, so a zone that needs a retry, a refactorisation or an
// order change does exactly that while the other zones of the launch do something else -- no lock-step over the batch.
// The Newton matrix I - gam*J is assembled (jacobian_build with diag = 1, mult = -gam), factored and solved by the same
// device routines backward Euler uses; with the reference's max_p_age = max_j_age = -1 every refactorisation is a
// rebuild, so no separate copy of dydotdy is kept (the module-level xb_jacobian_scale serves a host-driven solve_bdf).
// The Nordsieck arrays z, z0, zt0 [6][neq], the work vectors and the scalars (BdfScalars) live in the CTA's global scratch
// (L2 resident: the shared-memory tile of the dense LU leaves no room).  Scalar control runs on thread 0 between barriers, vector updates on all threads,
// weighted norms through the block reductions (strict arithmetic: sequential sums, as the reference forms them).
#pragma once

namespace xb {

constexpr int kBdfMaxOrder = 5;
constexpr int kBdfSkip = -1, kBdfSuccess = 0, kBdfFail = 1;
constexpr int kBdfMaxItTs = 20, kBdfMaxItNr = 3, kBdfMaxPAge = -1, kBdfMaxJAge = -1, kBdfMaxNerrTs = 10, kBdfMaxNcfTs = 10,
              kBdfNerrTsSmall = 2, kBdfNerrTsQreduce = 3, kBdfNstepQwait = 10;
constexpr double kBdfEtaMin = 0.1, kBdfEtaSmall = 0.2, kBdfEtaReduce = 0.25, kBdfEtaMax = 10.0, kBdfEtaThresh = 1.5,
                 kBdfDgmaxp = 0.3, kBdfDgmaxj = 0.2, kBdfTolNr = 0.1, kBdfCrDown = 0.3, kBdfCrDiverge = 2.0;
constexpr int kBdfVectors = 3 * (kBdfMaxOrder + 1) + 6;     // z, z0, zt0 [6] + ewt, acor, acorp, acorhat, dvec, ydot0 (doubles per equation)
constexpr int kBdfScalDoubles = (sizeof(BdfScalars) + 63) / 64 * 8;   // the scalars sit in front of the vectors (global scratch:
                                                                    // the shared-memory tile of the dense LU leaves no room)

// views of the zone's BDF vectors
struct BdfMem {
  double *z, *z0, *zt0, *ewt, *acor, *acorp, *acorhat, *dvec, *ydot0;
  int neq;
  __device__ __forceinline__ double &Z(double *a, int j, int i) const { return a[(size_t)j * neq + i]; }
};
__device__ __forceinline__ BdfMem bdf_mem(const ZoneScalars *s, const int neq) {
  BdfMem m; double *g = s->g_bdf + kBdfScalDoubles;
  m.neq = neq;
  m.z = g; m.z0 = g + (size_t)6 * neq; m.zt0 = g + (size_t)12 * neq;
  m.ewt = g + (size_t)18 * neq; m.acor = m.ewt + neq; m.acorp = m.acor + neq; m.acorhat = m.acorp + neq;
  m.dvec = m.acorhat + neq; m.ydot0 = m.dvec + neq;
  return m;
}

template <bool STRICT>
__device__ __forceinline__ double bdf_powi(const double a, const int n) {    // libgfortran pow_r8_i4: square and multiply
  using A = Ar<STRICT>;
  if (n == 0) return 1.0;
  double pw = 1.0, x = a; unsigned u = (unsigned)(n < 0 ? -n : n);
  for (;;) { if (u & 1u) pw = A::mul(pw, x); u >>= 1; if (u) x = A::mul(x, x); else break; }
  return n < 0 ? A::div(1.0, pw) : pw;
}
template <bool STRICT>
__device__ __forceinline__ double bdf_pow(const double x, const double y) {  // x ** y, real exponent
  return STRICT ? pm_exp(__dmul_rn(y, pm_log(x))) : pow(x, y);
}
__device__ __forceinline__ double bdf_spacing(const double x) {              // Fortran SPACING
  if (x == 0.0) return 2.2250738585072014e-308;
  const int e = (int)((__double_as_longlong(fabs(x)) >> 52) & 0x7ff);        // biased exponent of a normal number
  return __longlong_as_double((long long)max(e - 52, 1) << 52);
}
__device__ __forceinline__ int bdf_pascal(const int i, const int j) {        // binomial(i, j)
  constexpr int A[6][6] = {{1, 0, 0, 0, 0, 0}, {1, 1, 0, 0, 0, 0}, {1, 2, 1, 0, 0, 0}, {1, 3, 3, 1, 0, 0}, {1, 4, 6, 4, 1, 0}, {1, 5, 10, 10, 5, 1}};
  return A[i][j];
}
__device__ __forceinline__ double bdf_rtol(const int i, const int ny) { return (i < ny) ? c_ctl.tolc : c_ctl.tolt9; }
__device__ __forceinline__ double bdf_atol(const int i, const int ny) { return (i < ny) ? c_ctl.yacc : 1.0e-99; }

// weighted norm of x (normw): result in s->red[0], visible to all threads on return
template <bool STRICT>
__device__ __forceinline__ double bdf_normw(const ZoneMem &zm, const double *x, const double *wt, const int n) {
  using A = Ar<STRICT>;
  ZoneScalars *s = zm.s;
  const int ic = c_ctl.iconvc;
  if (ic == 0) {
    int idx;
    block_argmax_first(n, [&](int k) { return fabs(A::mul(x[k], wt[k])); }, zm.part, s->red, &idx);
    return fmax(s->red[0], 0.0);
  }
  block_sum2<STRICT>(n, [&](int k, double &ta, double &tb) {
    const double v = A::mul(x[k], wt[k]);
    ta = (ic == 1) ? fabs(v) : A::mul(v, v); tb = 0.0;
  }, zm.part, s->red);
  const double r = s->red[0];
  if (ic == 1) return r;
  if (ic == 2) return sqrt(r);
  return sqrt(A::div(r, (double)n));
}

template <bool STRICT>
__device__ __forceinline__ void bdf_error_weights(const double *y, double *wt, const int n, const int ny) {
  using A = Ar<STRICT>;
  const int ic = c_ctl.iconvc; const double ymin = c_ctl.ymin;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double rt = bdf_rtol(i, ny), at = bdf_atol(i, ny), v = y[i];
    double w;
    if (ic == 0 || ic == 1) {
      if (v < ymin) w = 0.0;
      else if (v < at) w = A::div(1.0, A::mul(rt, at));
      else w = A::div(1.0, A::mul(rt, fabs(v)));
    } else {
      const double a = fmax(fabs(v), ymin);
      w = A::div(1.0, A::add(A::mul(rt, a), at));
    }
    wt[i] = w;
  }
}

// all threads; thread 0 does the scalar part
template <bool STRICT>
__device__ __forceinline__ void bdf_reset(const ZoneMem &zm, const BdfMem &B) {
  ZoneScalars *s = zm.s; BdfScalars &b = *reinterpret_cast<BdfScalars *>(s->g_bdf);
  using A = Ar<STRICT>;
  const int tid = threadIdx.x, nt = blockDim.x, ny = c_net.ny, neq = B.neq;
  const bool heat = c_ctl.iheat > 0;
  const double tdel = s->tdel;
  for (int i = tid; i < neq; i += nt) {
    for (int j = 0; j <= kBdfMaxOrder; ++j) { B.Z(B.z, j, i) = 0.0; B.Z(B.z0, j, i) = 0.0; B.Z(B.zt0, j, i) = 0.0; }
    if (i < ny) { B.Z(B.z, 0, i) = zm.yt[i]; B.Z(B.z, 1, i) = A::mul(zm.ydot[i], tdel); B.ydot0[i] = zm.ydot[i]; }
    else if (heat) { B.Z(B.z, 0, i) = s->t9t; B.Z(B.z, 1, i) = A::mul(s->t9dot, tdel); }
    B.ewt[i] = 0.0; B.acor[i] = 0.0; B.acorp[i] = 0.0;
  }
  if (tid == 0) {
    b.retry_ts = 0; b.nit_nrslv = 0; b.nit_nr = 0; b.nit_ts = 0; b.ierr_nr = 0; b.ierr_ts = 0; b.nerr_ts = 0; b.ncf_ts = 0;
    b.q = 1; b.deltaq = 0; b.j_age = kBdfMaxJAge + 1; b.p_age = kBdfMaxPAge + 1; b.q_age = 0;
    for (int i = 0; i <= kBdfMaxOrder; ++i) { b.hvec[i] = tdel; b.lvec[i] = 0.0; }
    b.dt_scale = tdel; b.etaq[0] = b.etaq[1] = b.etaq[2] = 0.0; b.eta = 0.0; b.eta_next = kBdfEtaMax;
    b.gam = tdel; b.gamhat = tdel; b.gamratio = 1.0;
    b.acnrm = 0.0; b.crate = 0.0; b.tq[0] = b.tq[1] = b.tq[2] = b.tq[3] = 0.0; b.tq2save = 0.0;
    b.t9dot0 = s->t9dot;
  }
  __syncthreads();
}

// bdf_adjust_order: lvec by thread 0, then the Nordsieck update by all threads
template <bool STRICT>
__device__ __forceinline__ void bdf_adjust_order(const ZoneMem &zm, const BdfMem &B) {
  using A = Ar<STRICT>;
  ZoneScalars *s = zm.s; BdfScalars &b = *reinterpret_cast<BdfScalars *>(s->g_bdf);
  const int tid = threadIdx.x, nt = blockDim.x, neq = B.neq;
  const int dq = b.deltaq, q = b.q;
  __syncthreads();
  if (dq == -1) {
    if (tid == 0) {
      for (int i = 0; i <= kBdfMaxOrder; ++i) b.lvec[i] = 0.0;
      b.lvec[2] = 1.0;
      double hsum = 0.0;
      for (int j = 1; j <= q - 2; ++j) {
        hsum = A::add(hsum, b.hvec[j]);
        const double xi = A::div(hsum, s->tdel);
        for (int i = j + 2; i >= 2; --i) b.lvec[i] = A::add(A::mul(b.lvec[i], xi), b.lvec[i - 1]);
      }
    }
    __syncthreads();
    for (int i = tid; i < neq; i += nt) {
      const double zq = B.Z(B.z, q, i);
      for (int j = 2; j <= q - 1; ++j) B.Z(B.z, j, i) = A::sub(B.Z(B.z, j, i), A::mul(zq, b.lvec[j]));
      B.Z(B.z, q, i) = 0.0;
    }
    __syncthreads();
    if (tid == 0) { b.q = q - 1; b.q_age = 0; }
  } else if (dq == 1) {
    if (tid == 0) {
      for (int i = 0; i <= kBdfMaxOrder; ++i) b.lvec[i] = 0.0;
      b.lvec[2] = 1.0;
      double alpha0 = -1.0, alpha1 = 1.0, prod = 1.0, xiold = 1.0, hsum = s->tdel;
      for (int j = 1; j <= q - 1; ++j) {
        hsum = A::add(hsum, b.hvec[j + 1]);
        const double xi = A::div(hsum, s->tdel);
        prod = A::mul(prod, xi);
        alpha0 = A::sub(alpha0, A::div(1.0, (double)(j + 1)));
        alpha1 = A::add(alpha1, A::div(1.0, xi));
        for (int i = j + 2; i >= 2; --i) b.lvec[i] = A::add(A::mul(b.lvec[i], xiold), b.lvec[i - 1]);
        xiold = xi;
      }
      b.a1 = A::div(-A::add(alpha0, alpha1), prod);
    }
    __syncthreads();
    const double a1 = b.a1;
    for (int i = tid; i < neq; i += nt) {
      const double ac = B.acor[i];
      for (int j = 2; j <= q; ++j) B.Z(B.z, j, i) = A::add(B.Z(B.z, j, i), A::mul(A::mul(a1, ac), b.lvec[j]));
      B.Z(B.z, q + 1, i) = A::mul(a1, ac);
    }
    __syncthreads();
    if (tid == 0) { b.q = q + 1; b.q_age = 0; }
  }
  if (tid == 0) b.deltaq = 0;
  __syncthreads();
}

template <bool STRICT>
__device__ __forceinline__ void bdf_rescale(const ZoneMem &zm, const BdfMem &B) {     // mask true
  using A = Ar<STRICT>;
  ZoneScalars *s = zm.s; BdfScalars &b = *reinterpret_cast<BdfScalars *>(s->g_bdf);
  const int tid = threadIdx.x, nt = blockDim.x, neq = B.neq;
  __syncthreads();
  const double eta = b.eta; const int q = b.q;
  if (tid == 0) {
    s->tdel = A::mul(eta, b.dt_scale); b.hvec[0] = s->tdel; s->tt = A::add(s->t, s->tdel);
    b.dt_scale = s->tdel;
  }
  double f[kBdfMaxOrder + 1];
#pragma unroll
  for (int j = 1; j <= kBdfMaxOrder; ++j) f[j] = bdf_powi<STRICT>(eta, j);
  for (int i = tid; i < neq; i += nt)
    for (int j = 1; j <= q; ++j) B.Z(B.z, j, i) = A::mul(B.Z(B.z, j, i), f[j]);
  __syncthreads();
}

template <bool STRICT>
__device__ __forceinline__ void bdf_update(const ZoneMem &zm, const BdfMem &B, const bool first) {    // zone active
  using A = Ar<STRICT>;
  ZoneScalars *s = zm.s; BdfScalars &b = *reinterpret_cast<BdfScalars *>(s->g_bdf);
  if (threadIdx.x == 0) {
    const int q = b.q; const double tdel = s->tdel;
    double *l = b.lvec;
    l[0] = 1.0; l[1] = 1.0;
    for (int i = 2; i <= kBdfMaxOrder; ++i) l[i] = 0.0;
    double xi_inv = 1.0, xistar_inv = 1.0, alpha0 = -1.0, alpha0_hat = -1.0, hsum = tdel;
    if (q > 1) {
      for (int j = 2; j <= q - 1; ++j) {
        hsum = A::add(hsum, b.hvec[j - 1]); xi_inv = A::div(tdel, hsum); alpha0 = A::sub(alpha0, A::div(1.0, (double)j));
        for (int i = j; i >= 1; --i) l[i] = A::add(l[i], A::mul(l[i - 1], xi_inv));
      }
      alpha0 = A::sub(alpha0, A::div(1.0, (double)q));
      xistar_inv = A::sub(-l[1], alpha0);
      hsum = A::add(hsum, b.hvec[q - 1]); xi_inv = A::div(tdel, hsum);
      alpha0_hat = A::sub(-l[1], xi_inv);
      for (int i = q; i >= 1; --i) l[i] = A::add(l[i], A::mul(l[i - 1], xistar_inv));
    }
    const double a1 = A::add(A::sub(1.0, alpha0_hat), alpha0), a2 = A::add(1.0, A::mul((double)q, a1));
    b.tq[1] = fabs(A::div(a1, A::mul(alpha0, a2)));
    b.tq[3] = fabs(A::div(A::mul(a2, xistar_inv), A::mul(l[q], xi_inv)));
    if (q > 1) {
      const double c = A::div(xistar_inv, l[q]), a3 = A::add(alpha0, A::div(1.0, (double)q)), a4 = A::add(alpha0_hat, xi_inv);
      b.tq[0] = fabs(A::div(A::mul(c, A::add(A::sub(1.0, a4), a3)), a3));
    } else b.tq[0] = 1.0;
    hsum = A::add(hsum, b.hvec[q]); xi_inv = A::div(tdel, hsum);
    const double a5 = A::sub(alpha0, A::div(1.0, (double)(q + 1))), a6 = A::sub(alpha0_hat, xi_inv);
    const double cinv = A::div(A::add(A::sub(1.0, a6), a5), a2);
    b.tq[2] = fabs(A::div(cinv, A::mul(A::mul(xi_inv, (double)(q + 2)), a5)));
    b.gam = A::div(tdel, l[1]);
    if (!first) b.gamratio = A::div(b.gam, b.gamhat); else { b.gamhat = b.gam; b.gamratio = 1.0; }
  }
  bdf_error_weights<STRICT>(B.z, B.ewt, B.neq, c_net.ny);
  __syncthreads();
}

template <bool STRICT>
__device__ __forceinline__ void bdf_predict(const ZoneMem &zm, const BdfMem &B) {
  using A = Ar<STRICT>;
  const int q = reinterpret_cast<const BdfScalars *>(zm.s->g_bdf)->q, neq = B.neq;
  for (int k = threadIdx.x; k < neq; k += blockDim.x) {
    double zz[kBdfMaxOrder + 1], zt[kBdfMaxOrder + 1];
#pragma unroll
    for (int j = 0; j <= kBdfMaxOrder; ++j) { zz[j] = B.Z(B.z, j, k); zt[j] = zz[j]; B.Z(B.z0, j, k) = zz[j]; }
#pragma unroll
    for (int i = 0; i <= kBdfMaxOrder; ++i)
#pragma unroll
      for (int j = i + 1; j <= kBdfMaxOrder; ++j)
        if (i <= q && j <= q) zt[i] = A::add(zt[i], A::mul(zz[j], (double)bdf_pascal(j, i)));
#pragma unroll
    for (int j = 0; j <= kBdfMaxOrder; ++j) B.Z(B.zt0, j, k) = zt[j];
  }
  __syncthreads();
}

// bdf_step: the corrector iteration of one attempt (zone active)
template <bool STRICT>
__device__ __noinline__ void bdf_step(const bool active, const int max_it_nrslv) {
  using A = Ar<STRICT>;
  XB_ZONE_PROLOGUE
  ZoneScalars *s = zm.s; BdfScalars &b = *reinterpret_cast<BdfScalars *>(s->g_bdf);
  const int tid = threadIdx.x, nt = blockDim.x, ny = net.ny, N = net.msize, neq = N;
  const bool heat = ctl.iheat > 0;
  const BdfMem B = bdf_mem(s, neq);
  if (active) {
    for (int k = tid; k < neq; k += nt) {
      B.acor[k] = 0.0;
      if (k < ny) zm.yt[k] = B.Z(B.zt0, 0, k);
    }
    if (tid == 0) {
      if (heat) s->t9t = B.Z(B.zt0, 0, neq - 1);
      b.nit_nr = 0; b.nit_nrslv = 0;
    }
  }
  if (tid == 0) {
    b.refactor = active && (b.retry_ts || b.p_age > kBdfMaxPAge || fabs(A::sub(b.gamratio, 1.0)) > kBdfDgmaxp);
    b.rebuild = b.refactor && b.j_age > kBdfMaxJAge;
    b.iterate = active; b.crate = 1.0; b.delp = 0.0;
    if (!heat && b.retry_ts) t9rhofind<STRICT>(s, s->tt, &s->ntt, &s->t9t, &s->rhot, s->nt);
  }
  __syncthreads();
  for (int kit = 1; kit <= max_it_nrslv; ++kit) {
    const bool iterate = b.iterate != 0, refactor = b.refactor != 0, rebuild = b.rebuild != 0;
    __syncthreads();
    if (tid == 0 && iterate) { b.nit_nr += 1; b.nit_nrslv += 1; }
    const bool eval_rates = heat ? iterate : (kit == 1 && iterate && s->nh > 1);
    XB_PHASE_BEGIN(zm);
    if (eval_rates) cross_sect_cta<STRICT>();
    XB_PHASE_END(zm, PH_RATES);
    if (iterate) yderiv_cta<STRICT>();
    XB_PHASE_END(zm, PH_YDERIV);
    // Newton matrix: diag = 1, mult = -gam.  With max_j_age = max_p_age = -1 a refactorisation without a rebuild
    // (refactor && !rebuild) re-creates the matrix of an unchanged state from an unchanged gam: it is assembled again.
    if (refactor) {
      if (tid == 0) { s->rdt = 1.0; s->mult = -b.gam; }
      __syncthreads();
      jacobian_build_cta<STRICT>();
      if (tid == 0 && !rebuild) s->ktot[2] -= 1;       // jacobian_scale does not count as a build
      XB_PHASE_END(zm, PH_JAC);
      if (tid == 0) {
        b.gamhat = b.gam; b.gamratio = 1.0; b.crate = 1.0; b.p_age = 0; b.refactor = 0;
        if (rebuild) { b.j_age = 0; b.rebuild = 0; }
      }
      __syncthreads();
    }
    if (iterate) {
      // right-hand side (uses the gamratio the refactorisation has just reset), in slot order
      const double cstar = A::div(2.0, A::add(1.0, b.gamratio)), gam = b.gam, l1 = b.lvec[1];
      for (int k = tid; k < ny; k += nt)
        zm.rhs[zm.pos[k]] = A::mul(cstar, A::sub(A::sub(A::mul(gam, zm.ydot[k]), A::div(B.Z(B.zt0, 1, k), l1)), B.acor[k]));
      if (heat && tid == 0)
        zm.rhs[zm.pos[ny]] = A::mul(cstar, A::sub(A::sub(A::mul(gam, s->t9dot), A::div(B.Z(B.zt0, 1, neq - 1), l1)), B.acor[neq - 1]));
      __syncthreads();
    }
    if (refactor) {
      lu_cta<STRICT>(zm, N, iterate, true);
      XB_PHASE_END(zm, PH_LU);
    }
    if (iterate) {
      bksub_cta<STRICT>(zm, N, refactor && (zm.tile || s->rhs_in_jac));
      XB_PHASE_END(zm, PH_BKSUB);
      // apply the correction (dy = rhs in equation order)
      const double ymin = ctl.ymin; const double *aa = net.aa;
      for (int k = tid; k < ny; k += nt) {
        const double zt = B.Z(B.zt0, 0, k), ac = B.acor[k];
        double dy = zm.rhs[k], yt = A::add(A::add(zt, ac), dy);
        if (yt < ymin) { yt = 0.0; dy = A::sub(-zt, ac); }
        else if (A::mul(yt, aa[k]) > 1.0) { yt = A::div(1.0, aa[k]); dy = A::sub(A::sub(A::div(1.0, aa[k]), zt), ac); }
        zm.yt[k] = yt; B.acor[k] = A::add(ac, dy); B.dvec[k] = dy;
      }
      if (heat && tid == 0) {
        const double dt9 = zm.rhs[ny];
        B.dvec[neq - 1] = dt9; B.acor[neq - 1] = A::add(B.acor[neq - 1], dt9);
        s->t9t = A::add(B.Z(B.zt0, 0, neq - 1), B.acor[neq - 1]);
      }
      __syncthreads();
      const double del = bdf_normw<STRICT>(zm, B.dvec, B.ewt, neq);
      if (tid == 0) {
        if (b.nit_nr > 1) b.crate = fmax(A::mul(kBdfCrDown, b.crate), A::div(del, b.delp));
        const double dcon = A::mul(A::mul(del, fmin(1.0, b.crate)), b.tq[1]);
        b.conv = dcon < kBdfTolNr; b.del = del;
      }
      __syncthreads();
      if (b.conv && b.nit_nr > 1) {
        const double an = bdf_normw<STRICT>(zm, B.acor, B.ewt, neq);
        if (tid == 0) b.acnrm = an;
      } else if (b.conv && tid == 0) b.acnrm = del;
      __syncthreads();
      bool reset_iter = false;
      if (b.conv) { if (tid == 0) b.iterate = 0; }
      else if (b.nit_nrslv == max_it_nrslv) { if (tid == 0) { b.iterate = 0; b.nit_nr = kBdfMaxItNr + 1; } }
      else if (b.nit_nr == kBdfMaxItNr || (b.nit_nr > 1 && del > A::mul(b.delp, kBdfCrDiverge))) {
        if (b.j_age > 0) reset_iter = true;
        else if (tid == 0) { b.iterate = 0; b.nit_nr = kBdfMaxItNr + 1; }
      }
      __syncthreads();
      if (reset_iter) {
        for (int k = tid; k < neq; k += nt) { B.acor[k] = 0.0; if (k < ny) zm.yt[k] = B.Z(B.zt0, 0, k); }
        if (tid == 0) {
          if (heat) s->t9t = B.Z(B.zt0, 0, neq - 1);
          b.refactor = 1; b.rebuild = fabs(A::sub(b.gamratio, 1.0)) < kBdfDgmaxj; b.nit_nr = 0;
        }
      }
      if (tid == 0) b.delp = del;
      __syncthreads();
    }
    XB_PHASE_END(zm, PH_OTHER);
    if (!b.iterate) break;
  }
  __syncthreads();
}

template <bool STRICT>
__device__ __forceinline__ void bdf_check(const ZoneMem &zm, const BdfMem &B, const bool active) {
  using A = Ar<STRICT>;
  ZoneScalars *s = zm.s; BdfScalars &b = *reinterpret_cast<BdfScalars *>(s->g_bdf);
  const int tid = threadIdx.x, nt = blockDim.x, ny = c_net.ny, neq = B.neq;
  const bool heat = c_ctl.iheat > 0;
  __syncthreads();
  if (tid == 0) {
    const double biasq = 6.0, addon = 1.0e-6;
    b.restore = 0; b.rescale = 0; b.retry_ts = 0; b.reload = 0;
    if (active) {
      const double sp = bdf_spacing(s->t);
      if (b.nit_nr > kBdfMaxItNr) {
        b.ncf_ts += 1; b.eta_next = 1.0;
        if (b.nit_ts >= kBdfMaxItTs || s->tdel <= sp || b.ncf_ts == kBdfMaxNcfTs) { b.ierr_nr = kBdfSkip; b.ierr_ts = kBdfFail; }
        else {
          b.ierr_nr = kBdfFail; b.ierr_ts = kBdfSuccess; b.retry_ts = 1; b.restore = 1;
          b.eta = fmax(kBdfEtaReduce, A::div(sp, s->tdel)); b.rescale = 1;
          s->kmon[1] = c_ctl.kitmx + 1;
        }
      } else {
        const double error = A::mul(b.tq[1], b.acnrm);
        if (error <= 1.0) { b.ierr_nr = kBdfSuccess; b.ierr_ts = kBdfSuccess; }
        else {
          b.nerr_ts += 1; b.eta_next = 1.0;
          if (b.nerr_ts == kBdfMaxNerrTs || s->tdel <= sp) { b.ierr_nr = kBdfSkip; b.ierr_ts = kBdfFail; }
          else if (b.nerr_ts > kBdfNerrTsQreduce) {
            b.ierr_nr = kBdfFail; b.ierr_ts = kBdfSuccess; b.retry_ts = 1;
            b.eta = fmax(kBdfEtaMin, A::div(sp, s->tdel)); b.rescale = 1;
            if (b.q == 1) { b.q_age = b.q - kBdfNstepQwait; b.reload = 1; b.restore = 0; }
            else { b.restore = 1; b.deltaq = -1; }
          } else {
            b.ierr_nr = kBdfFail; b.ierr_ts = kBdfSuccess; b.retry_ts = 1; b.restore = 1;
            double e = A::div(1.0, A::add(bdf_pow<STRICT>(A::mul(biasq, error), A::div(1.0, (double)(b.q + 1))), addon));
            e = fmin(kBdfEtaMax, fmax(kBdfEtaMin, fmax(A::div(sp, s->tdel), e)));
            if (b.nerr_ts >= kBdfNerrTsSmall) e = fmin(kBdfEtaSmall, e);
            b.eta = e; b.rescale = 1;
          }
        }
      }
      s->ktot[1] += b.nit_nrslv;
      b.nit_ts += 1;
    }
  }
  __syncthreads();
  if (b.reload) {       // order 1 and repeated error test failures: reload the Nordsieck vector from the start of the step
    const double tdel = s->tdel;
    for (int k = tid; k < neq; k += nt) {
      if (k < ny) { B.Z(B.z, 0, k) = zm.y[k]; B.Z(B.z, 1, k) = A::mul(B.ydot0[k], tdel); }
      else if (heat) { B.Z(B.z, 0, k) = s->t9; B.Z(B.z, 1, k) = A::mul(b.t9dot0, tdel); }
    }
  }
  if (b.restore)
    for (int k = tid; k < neq; k += nt)
      for (int j = 0; j <= kBdfMaxOrder; ++j) B.Z(B.z, j, k) = B.Z(B.z0, j, k);
  bdf_adjust_order<STRICT>(zm, B);
  if (b.rescale) bdf_rescale<STRICT>(zm, B);
  __syncthreads();
}

template <bool STRICT>
__device__ __forceinline__ void bdf_correct_eta(const ZoneMem &zm, const BdfMem &B) {     // bdf_correct + bdf_eta
  using A = Ar<STRICT>;
  ZoneScalars *s = zm.s; BdfScalars &b = *reinterpret_cast<BdfScalars *>(s->g_bdf);
  const int tid = threadIdx.x, nt = blockDim.x, ny = c_net.ny, neq = B.neq;
  const bool heat = c_ctl.iheat > 0;
  __syncthreads();
  if (b.ierr_nr != kBdfSuccess) return;
  const int q = b.q;
  for (int i = tid; i < neq; i += nt) {
    B.Z(B.z, 0, i) = (i < ny) ? zm.yt[i] : (heat ? s->t9t : 0.0);
    const double ac = B.acor[i];
    for (int j = 1; j <= q; ++j) B.Z(B.z, j, i) = A::add(B.Z(B.zt0, j, i), A::mul(ac, b.lvec[j]));
  }
  if (tid == 0) {
    for (int j = kBdfMaxOrder; j >= 1; --j) b.hvec[j] = b.hvec[j - 1];
    b.q_age += 1; b.j_age += 1; b.p_age += 1;
  }
  __syncthreads();
  // bdf_eta
  const double biasqm1 = 6.0, biasq = 6.0, biasqp1 = 10.0, addon = 1.0e-6;
  if (b.eta_next > 1.0) {
    const bool consider = b.q_age > q;
    double nm1 = 0.0, np1 = 0.0;
    if (consider && q > 1) nm1 = bdf_normw<STRICT>(zm, B.z + (size_t)q * neq, B.ewt, neq);
    if (consider && q < kBdfMaxOrder) {
      const double c = A::mul(A::div(b.tq[3], b.tq2save), bdf_powi<STRICT>(A::div(b.hvec[1], b.hvec[2]), q + 1));
      for (int i = tid; i < neq; i += nt) B.acorhat[i] = A::sub(B.acor[i], A::mul(c, B.acorp[i]));
      __syncthreads();
      np1 = bdf_normw<STRICT>(zm, B.acorhat, B.ewt, neq);
    }
    if (tid == 0) {
      b.etaq[0] = b.etaq[1] = b.etaq[2] = 1.0;
      const double e0 = A::mul(b.tq[1], b.acnrm);
      b.etaq[1] = A::div(1.0, A::add(bdf_pow<STRICT>(A::mul(biasq, e0), A::div(1.0, (double)(q + 1))), addon));
      if (consider) {
        if (q > 1) {
          const double em = A::mul(b.tq[0], nm1);
          b.etaq[0] = A::div(1.0, A::add(bdf_pow<STRICT>(A::mul(biasqm1, em), A::div(1.0, (double)q)), addon));
        }
        if (q < kBdfMaxOrder) {
          const double ep = A::mul(b.tq[2], np1);
          b.etaq[2] = A::div(1.0, A::add(bdf_pow<STRICT>(A::mul(biasqp1, ep), A::div(1.0, (double)(q + 2))), addon));
        }
        b.q_age = q - 1;
      }
      b.eta = 1.0;
      if (b.etaq[1] > b.eta) { b.eta = b.etaq[1]; b.deltaq = 0; }
      if (b.etaq[0] > b.eta) { b.eta = b.etaq[0]; b.deltaq = -1; }
      if (b.etaq[2] > b.eta) { b.eta = b.etaq[2]; b.deltaq = 1; }
      if (b.eta <= kBdfEtaThresh) { b.eta = 1.0; b.deltaq = 0; }
      b.eta = fmin(b.eta, kBdfEtaMax);
    }
  } else if (tid == 0) b.q_age = min(b.q_age, q - 1);
  __syncthreads();
  for (int i = tid; i < neq; i += nt) B.acorp[i] = B.acor[i];
  if (tid == 0) b.tq2save = b.tq[3];
  __syncthreads();
}

// solve_bdf for one zone (all threads); s->its in/out as for solve_be
template <bool STRICT>
__device__ __noinline__ void solve_bdf_cta() {
  using A = Ar<STRICT>;
  XB_ZONE_PROLOGUE
  ZoneScalars *s = zm.s; BdfScalars &b = *reinterpret_cast<BdfScalars *>(s->g_bdf);
  const int tid = threadIdx.x, nt = blockDim.x, ny = net.ny, neq = net.msize;
  const bool heat = ctl.iheat > 0;
  const BdfMem B = bdf_mem(s, neq);
  const int max_it_nrslv = max(ctl.kitmx, kBdfMaxItNr);
  __syncthreads();
  const bool first = b.started == 0;
  __syncthreads();
  if (first) bdf_reset<STRICT>(zm, B);
  for (int k = tid; k < ny; k += nt) B.ydot0[k] = zm.ydot[k];
  if (tid == 0) {
    b.started = 1;
    b.ierr_ts = s->its;
    b.ierr_nr = (b.ierr_ts != kBdfSuccess) ? kBdfSkip : kBdfFail;
    if (heat) b.t9dot0 = s->t9dot;
    b.eta_next = kBdfEtaMax; b.nit_ts = 0; b.nerr_ts = 0; b.ncf_ts = 0;
    b.rescale = 0;
    if (!first && b.ierr_ts == kBdfSuccess && fabs(A::sub(A::div(s->tdel, s->tdel_old), 1.0)) > 2.220446049250313e-16) {
      b.eta = A::div(s->tdel, s->tdel_old); b.rescale = 1;
    }
  }
  __syncthreads();
  if (!first) {         // bdf_adjust
    bdf_adjust_order<STRICT>(zm, B);
    if (b.rescale) bdf_rescale<STRICT>(zm, B);
  }
  for (int kts = 1; kts <= kBdfMaxItTs; ++kts) {
    __syncthreads();
    const bool active = b.ierr_nr == kBdfFail;
    if (active) { bdf_update<STRICT>(zm, B, first); bdf_predict<STRICT>(zm, B); }
    bdf_step<STRICT>(active, max_it_nrslv);
    bdf_check<STRICT>(zm, B, active);
    if (b.ierr_nr != kBdfFail) break;
  }
  bdf_correct_eta<STRICT>(zm, B);
  __syncthreads();
  const bool ok = b.ierr_nr == kBdfSuccess;
  if (ok) for (int k = tid; k < ny; k += nt) { const double v = B.Z(B.z, 0, k); zm.yt[k] = v; zm.yo[k] = zm.y[k]; zm.y[k] = v; }
  if (tid == 0) {
    s->tdel_next = A::mul(b.eta, s->tdel);
    if (ok) {
      b.ierr_ts = kBdfSuccess;
      if (heat) s->t9t = B.Z(B.z, 0, neq - 1);
      s->nto = s->nt; s->nt = s->ntt; s->to = s->t; s->t = s->tt; s->t9o = s->t9; s->t9 = s->t9t;
      s->rhoo = s->rho; s->rho = s->rhot; s->yeo = s->ye; s->ye = s->yet;
    } else if (b.ierr_nr == kBdfFail || b.ierr_ts == kBdfFail) b.ierr_ts = kBdfFail;
    s->kmon[0] = b.nit_ts; s->ktot[0] += b.nit_ts; s->kmon[1] = b.nit_nrslv;
    s->its = b.ierr_ts;
  }
  __syncthreads();
}

}  // namespace xb
