/* Oracle (TEST INFRASTRUCTURE ONLY) -- plain-C CPU restatement of XNet's per-zone implicit step.
 *
 * This file is the checker for the CUDA path and the "port" CPU baseline of bench.py.  Product
 * code (xnet_b200/) must never include, link or call it.
 *
 * Restates, with the same operation order as the Fortran (paths relative to /root/reference):
 *   full_net                       source/xnet_evolve.F90:19-269
 *   solve_be / step_be             source/xnet_integrate_be.F90:32-482
 *   timestep / cross_sect / yderiv source/xnet_integrate.F90:13-984
 *   update_eos (Ye from y_moment)  source/xnet_integrate.F90:319-362, xnet_eos_starkiller.F90:107-204,
 *                                  source/xnet_abundances.F90:93-185
 *   partf / benuc / enudot         source/xnet_data.F90:119-283, 958-1023
 *   ffn_rate (table branch)        source/xnet_ffn.F90:139-324
 *   t9rhofind                      source/xnet_conditions.F90:50-102
 *   jacobian_build/scale/decomp/bksub  source/xnet_jacobian_dense.F90:136-434, 549-732
 *   DGETRF/DGETRF2/DGETRS/DLASWP/DTRSM/DGEMM/IDAMAX/DSCAL  tools/LAPACK/{dgetrf,dgetrf2,...}.f (Reference-LAPACK 3.8.0)
 *   safe_exp                       source/xnet_util.F90:17-18,55-72
 * Screening/EOS/self-heating (iscrn>0, iheat>0) live in ox_eos.c.
 *
 * Parity status: the reference cannot be built in this environment (no Fortran compiler) and ships
 * no numeric golden trajectories (SURVEY F1,F3).  Integer/index work is pinned by the net_desc
 * goldens (oracle/xnet_loader.py); the LU is pinned against LAPACK (scipy) in tests; the
 * floating-point trajectory itself is "parity unpinned" beyond code review.
 */
#include "xnet_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "ox_math.h"

#define NG 24
#define NT9GRID 13
#define NENEGRID 11
#define NGRID 143

/* xnet_constants.F90 */
static const double avn = 6.022140857e+23, epmev = 1.6021766208e-06;

/* eos/screening hooks implemented in ox_eos.c */
int ox_eos_interface(const ox_net *net, const ox_ctl *ctl, double t9, double rho, const double *y,
                     double xext, double aext, double zext, double *ye, double *cv, double *etae,
                     double *detaedt9);
int ox_screening(const ox_net *net, const ox_ctl *ctl, double t9, double rho, const double *y,
                 double xext, double aext, double zext, double etae, double detaedt9,
                 double *h[4], double *dh[4]);
void ox_effphase(double t9, double chem, double qec, double *phasei, double *dphaseidt9);

typedef struct {
  const ox_net *net; const ox_ctl *ctl;
  int ny, msize;
  /* trajectory */
  int nh; const double *th, *t9h, *rhoh; double rho_scale; int has_scale;
  double tstart, tstop, tdelstart;
  /* conditions */
  double t, tt, to, tdel, tdel_next, tdel_old;
  double t9, t9t, t9o, t9dot, rho, rhot, rhoo, ye, yet, yeo, cv, etae, detaedt9;
  int nt, ntt, nto, ints, intso, iweak;
  double xext, aext, zext;
  double *y, *yo, *yt, *ydot;
  double *csect[4], *dcsect[4], *h[4], *dh[4], *b[4];
  double *gg, *dlngdt9;             /* 0..ny */
  double *rffn, *dlnrffndt9, *qkffn; /* qkffn is 0..nffn */
  double *phasei, *dphaseidt9;
  double *dydotdy, *jac, *rhs; int *indx; int info;
  double *yrhs, *dy, *reldy; double t9rhs, dt9, relt9;
  int *n_n[4][4];                   /* expanded reactant indices n_cj [nan_c], 1-based */
  int kmon[5], ktot[5];
  double *rtau;
  /* BDF integrator (xnet_integrate_bdf.F90): Nordsieck arrays [6][neq] and per-zone scalars */
  int neq, bq, deltaq, q_age, j_age, p_age, nit_nrslv, nit_nr, nit_ts, nerr_ts, ncf_ts, ierr_nr, ierr_ts, retry_ts;
  int bdf_started;
  double *bz, *bz0, *bzt0, *ewt, *acor, *acorp, *acorhat, *dvec, *ydot0, *brtol, *batol;
  double hvec[6], lvec[6], dt_scale, etaq[3], eta, eta_next, gam, gamhat, gamratio, t9dot0, acnrm, crate, tq[4], tq2save;
} zone_t;

/* -------------------------------------------------------------------------------------------- */
static double safe_exp(double x) {
  /* xnet_util.F90:17-18,69 */
  const double ln_2 = log(2.0);
  const double exp_max = 1024.0 * ln_2 * 0.9, exp_min = -1021.0 * ln_2 * 0.9;
  double v = x > exp_min ? x : exp_min;
  v = v < exp_max ? v : exp_max;
  return M_EXP(v);
}

static double rhoh_at(const zone_t *z, int n) { /* n 1-based */
  return z->has_scale ? z->rhoh[n - 1] * z->rho_scale : z->rhoh[n - 1];
}

static void t9rhofind(const zone_t *z, double tf, int *nf, double *t9f, double *rhof) {
  /* xnet_conditions.F90:72-99 */
  int ns = z->nh, n;
  if (ns == 1) { *t9f = z->t9h[0]; *rhof = rhoh_at(z, 1); *nf = 1; return; }
  for (n = 1; n <= ns; n++) if (tf <= z->th[n - 1]) break;
  *nf = n;
  if (n > 1 && n <= ns) {
    double rdt = 1.0 / (z->th[n - 1] - z->th[n - 2]);
    double dt = tf - z->th[n - 2];
    double dt9 = z->t9h[n - 1] - z->t9h[n - 2];
    double drho = rhoh_at(z, n) - rhoh_at(z, n - 1);
    *t9f = dt * rdt * dt9 + z->t9h[n - 2];
    *rhof = dt * rdt * drho + rhoh_at(z, n - 1);
  } else if (n == 1) {
    *t9f = z->t9h[0]; *rhof = rhoh_at(z, 1);
  } else {
    *t9f = z->t9h[ns - 1]; *rhof = rhoh_at(z, ns);
  }
}

static void partf(zone_t *z, double t9) {
  /* xnet_data.F90:217-267 */
  const ox_net *net = z->net; int ny = z->ny, i, k, ii;
  const double *g = net->g, *t9i = net->t9i;
  for (i = 1; i <= NG; i++) if (t9 <= t9i[i - 1]) break;
  ii = i;
  z->gg[0] = 1.0;
  if (ii == 1) {
    for (k = 1; k <= ny; k++) z->gg[k] = g[(k - 1) * NG + 0];
  } else if (ii == NG + 1) {
    for (k = 1; k <= ny; k++) z->gg[k] = g[(k - 1) * NG + NG - 1];
  } else {
    double rdt9 = (t9 - t9i[ii - 2]) / (t9i[ii - 1] - t9i[ii - 2]);
    for (k = 1; k <= ny; k++)
      z->gg[k] = safe_exp(rdt9 * log(g[(k - 1) * NG + ii - 1]) + (1.0 - rdt9) * log(g[(k - 1) * NG + ii - 2]));
  }
  if (z->ctl->iheat > 0) {
    z->dlngdt9[0] = 0.0;
    if (ii == 1) {
      for (k = 1; k <= ny; k++) z->dlngdt9[k] = log(g[(k - 1) * NG + 1] / g[(k - 1) * NG + 0]) / (t9i[1] - t9i[0]);
    } else if (ii == NG + 1) {
      for (k = 1; k <= ny; k++)
        z->dlngdt9[k] = log(g[(k - 1) * NG + NG - 1] / g[(k - 1) * NG + NG - 2]) / (t9i[NG - 1] - t9i[NG - 2]);
    } else {
      for (k = 1; k <= ny; k++)
        z->dlngdt9[k] = log(g[(k - 1) * NG + ii - 1] / g[(k - 1) * NG + ii - 2]) / (t9i[ii - 1] - t9i[ii - 2]);
    }
  }
}

static void ffn_rate(zone_t *z, double t9, double ene) {
  /* xnet_ffn.F90:139-301 */
  static const double t9grid[NT9GRID] = {0.01, 0.1, 0.2, 0.4, 0.7, 1.0, 1.5, 2.0, 3.0, 5.0, 10.0, 30.0, 100.0};
  static const double enegrid[NENEGRID] = {1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0, 10.0, 11.0};
  const double lqkmin = -90.0, lrfmin = -30.0, rfmin = 1.0e-30;
  const double ln_2 = log(2.0), ln_10 = log(10.0);
  const double bok = 8.6173303e-02, m_e = 0.5109989461;
  const ox_net *net = z->net; int nffn = net->nffn, k, i, lt1;
  for (k = 1; k <= nffn; k++) {
    if (net->has_logft[k - 1] > 0) {
      double cheme = z->etae * bok * t9 + m_e;
      /* sign(cheme, -has_logft-1.5): magnitude of cheme with negative sign */
      cheme = -fabs(cheme);
      ox_effphase(t9, cheme, net->ffn_qval[k - 1], &z->phasei[k - 1], &z->dphaseidt9[k - 1]);
    }
  }
  for (i = 1; i <= NT9GRID; i++) if (t9 <= t9grid[i - 1]) break;
  lt1 = (i - 1 > 1) ? i - 1 : 1;
  for (k = 1; k <= nffn; k++) {
    double enel = M_LOG10(ene);
    int le1 = (int)enel; if (le1 < 1) le1 = 1;
    double dt9 = t9grid[lt1] - t9grid[lt1 - 1];
    double rdt9 = (t9 - t9grid[lt1 - 1]) / dt9;
    double dene = enegrid[le1] - enegrid[le1 - 1];
    double rdene = (enel - enegrid[le1 - 1]) / dene;
    int i1 = NT9GRID * (le1 - 1) + lt1, i2 = i1 + 1, i3 = NT9GRID * le1 + lt1, i4 = i3 + 1;
    const double *enu = net->ffnenu + (size_t)(k - 1) * NGRID;
    const double *sum = net->ffnsum + (size_t)(k - 1) * NGRID;
    double qdr1 = enu[i2 - 1] - enu[i1 - 1], qdr2 = enu[i4 - 1] - enu[i3 - 1];
    double qr1 = enu[i1 - 1] + rdt9 * qdr1, qr2 = enu[i3 - 1] + rdt9 * qdr2;
    double qk = qr1 + rdene * (qr2 - qr1);
    z->qkffn[k] = (qk > lqkmin) ? M_POW10(qk) : 0.0;
    if (net->has_logft[k - 1] > 0) {
      const double *ec = net->ffn_ec + (size_t)(k - 1) * NGRID;
      const double *be = net->ffn_beta + (size_t)(k - 1) * NGRID;
      double dr1_ec = ec[i2 - 1] - ec[i1 - 1], dr2_ec = ec[i4 - 1] - ec[i3 - 1];
      double r1 = ec[i1 - 1] + rdt9 * dr1_ec, r2 = ec[i3 - 1] + rdt9 * dr2_ec;
      double rf_ec = r1 + rdene * (r2 - r1), rf_beta, dr1, dr2, rf;
      if (z->phasei[k - 1] > rfmin && rf_ec > lrfmin) rf_ec = ln_2 * z->phasei[k - 1] / M_POW10(rf_ec);
      else rf_ec = 0.0;
      dr1 = be[i2 - 1] - be[i1 - 1]; dr2 = be[i4 - 1] - be[i3 - 1];
      r1 = be[i1 - 1] + rdt9 * dr1; r2 = be[i3 - 1] + rdt9 * dr2;
      rf_beta = r1 + rdene * (r2 - r1);
      rf_beta = (rf_beta > lrfmin) ? M_POW10(rf_beta) : 0.0;
      rf = rf_beta + rf_ec;
      if (rf_beta < rfmin && rf_ec < rfmin) { z->rffn[k - 1] = 0.0; z->dlnrffndt9[k - 1] = 0.0; }
      else {
        z->rffn[k - 1] = rf;
        if (rf_ec < rfmin) z->dlnrffndt9[k - 1] = ln_10 * (rdene * dr2 + (1.0 - rdene) * dr1) / dt9;
        else {
          double drbeta_dt = rf_beta * ln_10 * (rdene * dr2 + (1.0 - rdene) * dr1) / dt9;
          double drec_dt = rf_ec * (-ln_10 * (rdene * dr2_ec + (1.0 - rdene) * dr1_ec) / dt9
                                    + z->dphaseidt9[k - 1] / z->phasei[k - 1]);
          z->dlnrffndt9[k - 1] = (drbeta_dt + drec_dt) / rf;
        }
      }
    } else {
      double dr1 = sum[i2 - 1] - sum[i1 - 1], dr2 = sum[i4 - 1] - sum[i3 - 1];
      double r1 = sum[i1 - 1] + rdt9 * dr1, r2 = sum[i3 - 1] + rdt9 * dr2;
      double rf = r1 + rdene * (r2 - r1);
      if (rf < lrfmin) { z->rffn[k - 1] = 0.0; z->dlnrffndt9[k - 1] = 0.0; }
      else { z->rffn[k - 1] = M_POW10(rf); z->dlnrffndt9[k - 1] = ln_10 * (rdene * dr2 + (1.0 - rdene) * dr1) / dt9; }
    }
  }
}

static int cross_sect(zone_t *z) {
  /* xnet_integrate.F90:569-927 for one zone (mask true) */
  const ox_net *net = z->net; const ox_ctl *ctl = z->ctl;
  int iheat = ctl->iheat, c, k, j, rc_;
  double t09[7], dt09[7], ene, ascrn, t9t = z->t9t, rhot = z->rhot;
  /* update_eos: Ye always from y_moment; cv/eta only if iscrn>0 or iheat>0 */
  rc_ = ox_eos_interface(net, ctl, t9t, rhot, z->yt, z->xext, z->aext, z->zext,
                         &z->yet, &z->cv, &z->etae, &z->detaedt9);
  if (rc_) return rc_;
  if (ctl->iscrn >= 1) {
    ascrn = 1.0;
    rc_ = ox_screening(net, ctl, t9t, rhot, z->yt, z->xext, z->aext, z->zext, z->etae, z->detaedt9, z->h, z->dh);
    if (rc_) return rc_;
  } else ascrn = 0.0;
  /* update_iweak :300-311 */
  if (t9t < 0.01 && ctl->iweak0 != 0) z->iweak = -abs(ctl->iweak0); else z->iweak = ctl->iweak0;
  partf(z, t9t);
  ene = z->yet * rhot;
  t09[0] = 1.0; t09[1] = 1.0 / t9t; t09[4] = t9t;
#ifdef OX_PMATH
  { double l = M_LOG(t9t);   /* pow(x,y) evaluated as exp(y*log(x)) in the portable-math build */
    t09[2] = M_EXP((-1.0 / 3.0) * l); t09[3] = M_EXP((1.0 / 3.0) * l); t09[5] = M_EXP((5.0 / 3.0) * l); t09[6] = l; }
#else
  t09[2] = pow(t9t, -1.0 / 3.0); t09[3] = pow(t9t, 1.0 / 3.0); t09[5] = pow(t9t, 5.0 / 3.0); t09[6] = log(t9t);
#endif
  dt09[0] = 0.0; dt09[1] = -(1.0 / (t9t * t9t)); dt09[2] = -M_POW(t9t, -4.0 / 3.0) / 3.0;
  dt09[3] = M_POW(t9t, -2.0 / 3.0) / 3.0; dt09[4] = 1.0; dt09[5] = M_POW(t9t, 2.0 / 3.0) * 5.0 / 3.0;
  dt09[6] = 1.0 / t9t;
  if (net->nffn > 0) ffn_rate(z, t9t, ene);
  /* nnu_rate: ineutrino=0 -> rnnu = 0 (xnet_nnu.F90:222-233) */
  for (c = 0; c < 4; c++) {
    int nr = net->nr[c];
    for (k = 0; k < nr; k++) {
      double lam = 0.0, dlam = 0.0;
      for (j = 0; j < 7; j++) { lam = lam + net->rc[c][k * 7 + j] * t09[j]; dlam = dlam + net->rc[c][k * 7 + j] * dt09[j]; }
      z->h[c][k] = ascrn * z->h[c][k] + lam;
      z->dh[c][k] = ascrn * z->dh[c][k] + dlam;
    }
  }
  for (c = 0; c < 4; c++) {
    int nr = net->nr[c], nreact = c + 1;
    double rhofac = 1.0;
    if (c == 1) rhofac = rhot; else if (c == 2) rhofac = rhot * rhot; else if (c == 3) rhofac = rhot * rhot * rhot;
    for (k = 0; k < nr; k++) {
      int iwk = net->iwk[c][k];
      const int *ni = net->ni[c] + k * 6;
      if ((z->iweak < 0 && iwk == 0) || (z->iweak == 0 && iwk != 0)) {
        z->csect[c][k] = 0.0; z->dcsect[c][k] = 0.0; continue;
      }
      double rpf = 1.0, dlnrpf = 0.0;
      if (net->irev[c][k] == 1) {
        int nslot = (c == 0) ? 5 : 6;
        double p, s;
        if (c == 0) { p = z->gg[ni[0]]; s = z->dlngdt9[ni[0]]; }
        else { p = 1.0; s = 0.0; for (j = 0; j < nreact; j++) { p = p * z->gg[ni[j]]; s = s + z->dlngdt9[ni[j]]; } }
        for (j = nreact; j < nslot; j++) { rpf = rpf * z->gg[ni[j]]; dlnrpf = dlnrpf + z->dlngdt9[ni[j]]; }
        rpf = rpf / p; dlnrpf = dlnrpf - s;
      }
      double cs;
      if (c == 0) {
        if (iwk == 1) cs = ene * rpf * safe_exp(z->h[0][k]);
        else if (iwk == 2 || iwk == 3) cs = z->rffn[net->iffn[k] - 1];
        else if (iwk == 7 || iwk == 8) cs = rpf * 0.0;
        else cs = rpf * safe_exp(z->h[0][k]);
      } else {
        if (iwk == 1) cs = rhofac * ene * rpf * safe_exp(z->h[c][k]);
        else cs = rhofac * rpf * safe_exp(z->h[c][k]);
        if (c == 2 && cs < 1.0e-20) cs = 0.0;
      }
      z->csect[c][k] = cs;
      if (iheat > 0) {
        if (c == 0 && (iwk == 2 || iwk == 3)) z->dcsect[c][k] = cs * z->dlnrffndt9[net->iffn[k] - 1];
        else if (c == 0 && (iwk == 7 || iwk == 8)) z->dcsect[c][k] = cs * dlnrpf;
        else z->dcsect[c][k] = cs * (z->dh[c][k] + dlnrpf);
      }
    }
  }
  z->ktot[4] += 1;
  return 0;
}

static void yderiv(zone_t *z) {
  /* xnet_integrate.F90:409-495 */
  const ox_net *net = z->net; int ny = z->ny, i0, i1, c;
  const double *yt = z->yt;
  for (i0 = 0; i0 < ny; i0++) {
    double s[4];
    for (c = 0; c < 4; c++) {
      int la = net->la[c * ny + i0], le = net->le[c * ny + i0];
      double sc = 0.0;
      for (i1 = la; i1 <= le; i1++) {
        double bb = net->a[c][i1 - 1] * z->csect[c][net->mu[c][i1 - 1] - 1];
        double t;
        z->b[c][i1 - 1] = bb;
        t = bb * yt[z->n_n[c][0][i1 - 1] - 1];
        if (c >= 1) t = t * yt[z->n_n[c][1][i1 - 1] - 1];
        if (c >= 2) t = t * yt[z->n_n[c][2][i1 - 1] - 1];
        if (c >= 3) t = t * yt[z->n_n[c][3][i1 - 1] - 1];
        sc = sc + t;
      }
      s[c] = sc;
    }
    z->ydot[i0] = s[0] + s[1] + s[2] + s[3];
  }
  if (z->ctl->iheat > 0) {
    double sdot = 0.0;
    for (i0 = 0; i0 < ny; i0++) sdot = sdot - net->mex[i0] * z->ydot[i0] / z->cv;
    z->t9dot = sdot;
  }
  z->ktot[3] += 1;
}

static void jacobian_scale(zone_t *z, double diag, double mult) {
  /* xnet_jacobian_dense.F90:183-200: jac(i0,j1) = mult*dydotdy(j1,i0) (+diag); jac column-major */
  int m = z->msize, i0, j1; const double *d = z->dydotdy;
  for (j1 = 0; j1 < m; j1++)
    for (i0 = 0; i0 < m; i0++) {
      double v = mult * d[(size_t)i0 * m + j1];
      if (i0 == j1) v = v + diag;
      z->jac[(size_t)j1 * m + i0] = v;
    }
}

static void jacobian_build(zone_t *z, double diag, double mult) {
  /* xnet_jacobian_dense.F90:303-399 then jacobian_scale :183-200 */
  const ox_net *net = z->net; int ny = z->ny, m = z->msize, i0, j1, c;
  const double *yt = z->yt; double *d = z->dydotdy;
  int iheat = z->ctl->iheat;
  /* d[(i0)*m + j1] == dydotdy(j1+1,i0+1) */
  for (i0 = 0; i0 < ny; i0++) {
    double *row = d + (size_t)i0 * m;
    for (j1 = 0; j1 < m; j1++) row[j1] = 0.0;
    {
      int la = net->la[0 * ny + i0], le = net->le[0 * ny + i0];
      for (j1 = la; j1 <= le; j1++) row[z->n_n[0][0][j1 - 1] - 1] += z->b[0][j1 - 1];
    }
    {
      int la = net->la[1 * ny + i0], le = net->le[1 * ny + i0];
      for (j1 = la; j1 <= le; j1++) {
        int n1 = z->n_n[1][0][j1 - 1] - 1, n2 = z->n_n[1][1][j1 - 1] - 1; double bb = z->b[1][j1 - 1];
        row[n1] += bb * yt[n2];
        row[n2] += bb * yt[n1];
      }
    }
    {
      int la = net->la[2 * ny + i0], le = net->le[2 * ny + i0];
      for (j1 = la; j1 <= le; j1++) {
        int n1 = z->n_n[2][0][j1 - 1] - 1, n2 = z->n_n[2][1][j1 - 1] - 1, n3 = z->n_n[2][2][j1 - 1] - 1;
        double bb = z->b[2][j1 - 1];
        row[n1] += bb * yt[n2] * yt[n3];
        row[n2] += bb * yt[n3] * yt[n1];
        row[n3] += bb * yt[n1] * yt[n2];
      }
    }
    {
      int la = net->la[3 * ny + i0], le = net->le[3 * ny + i0];
      for (j1 = la; j1 <= le; j1++) {
        int n1 = z->n_n[3][0][j1 - 1] - 1, n2 = z->n_n[3][1][j1 - 1] - 1, n3 = z->n_n[3][2][j1 - 1] - 1,
            n4 = z->n_n[3][3][j1 - 1] - 1;
        double bb = z->b[3][j1 - 1];
        row[n1] += bb * yt[n2] * yt[n3] * yt[n4];
        row[n2] += bb * yt[n3] * yt[n4] * yt[n1];
        row[n3] += bb * yt[n4] * yt[n1] * yt[n2];
        row[n4] += bb * yt[n1] * yt[n2] * yt[n3];
      }
    }
    if (iheat > 0) {
      double s[4];
      for (c = 0; c < 4; c++) {
        int la = net->la[c * ny + i0], le = net->le[c * ny + i0]; double sc = 0.0;
        for (j1 = la; j1 <= le; j1++) {
          double t = net->a[c][j1 - 1] * z->dcsect[c][net->mu[c][j1 - 1] - 1] * yt[z->n_n[c][0][j1 - 1] - 1];
          if (c >= 1) t = t * yt[z->n_n[c][1][j1 - 1] - 1];
          if (c >= 2) t = t * yt[z->n_n[c][2][j1 - 1] - 1];
          if (c >= 3) t = t * yt[z->n_n[c][3][j1 - 1] - 1];
          sc = sc + t;
        }
        s[c] = sc;
      }
      row[ny] = s[0] + s[1] + s[2] + s[3];
    }
  }
  if (iheat > 0) {
    for (j1 = 0; j1 < m; j1++) {
      double sdot = 0.0;
      for (i0 = 0; i0 < ny; i0++) sdot = sdot - net->mex[i0] * d[(size_t)i0 * m + j1] / z->cv;
      d[(size_t)ny * m + j1] = sdot;
    }
  }
  z->ktot[2] += 1;
  jacobian_scale(z, diag, mult);
}

/* ------------------------------- LAPACK 3.8.0 restatement ----------------------------------- */
#define A_(i, j) a[(size_t)((j)-1) * lda + ((i)-1)]
static void dlaswp(int n, double *a, int lda, int k1, int k2, const int *ipiv) {
  /* incx = 1; the 32-column blocking of dlaswp.f does not change results */
  int i, k;
  for (i = k1; i <= k2; i++) {
    int ip = ipiv[i - 1];
    if (ip != i)
      for (k = 1; k <= n; k++) { double t = A_(i, k); A_(i, k) = A_(ip, k); A_(ip, k) = t; }
  }
}
static void dtrsm_llnu(int m, int n, const double *a, int lda, double *b, int ldb) {
  int i, j, k;
  for (j = 1; j <= n; j++)
    for (k = 1; k <= m; k++) {
      double bkj = b[(size_t)(j - 1) * ldb + k - 1];
      if (bkj != 0.0)
        for (i = k + 1; i <= m; i++) b[(size_t)(j - 1) * ldb + i - 1] -= bkj * A_(i, k);
    }
}
static void dtrsm_lunn(int m, int n, const double *a, int lda, double *b, int ldb) {
  int i, j, k;
  for (j = 1; j <= n; j++)
    for (k = m; k >= 1; k--) {
      double *bkj = &b[(size_t)(j - 1) * ldb + k - 1];
      if (*bkj != 0.0) {
        *bkj = *bkj / A_(k, k);
        for (i = 1; i <= k - 1; i++) b[(size_t)(j - 1) * ldb + i - 1] -= *bkj * A_(i, k);
      }
    }
}
static void dgemm_nn_m1(int m, int n, int k, const double *a, int lda, const double *b, int ldb, double *c, int ldc) {
  /* C := C - A*B  (alpha=-1, beta=1), reference loop order j,l,i */
  int i, j, l;
  for (j = 1; j <= n; j++)
    for (l = 1; l <= k; l++) {
      double temp = -1.0 * b[(size_t)(j - 1) * ldb + l - 1];
      for (i = 1; i <= m; i++) c[(size_t)(j - 1) * ldc + i - 1] += temp * A_(i, l);
    }
}
static void dgetrf2(int m, int n, double *a, int lda, int *ipiv, int *info) {
  *info = 0;
  if (m == 0 || n == 0) return;
  if (m == 1) {
    ipiv[0] = 1;
    if (A_(1, 1) == 0.0) *info = 1;
  } else if (n == 1) {
    const double sfmin = DBL_MIN; /* dlamch('S') */
    int i, imax = 1; double dmax = fabs(A_(1, 1));
    for (i = 2; i <= m; i++) if (fabs(A_(i, 1)) > dmax) { imax = i; dmax = fabs(A_(i, 1)); }
    ipiv[0] = imax;
    if (A_(imax, 1) != 0.0) {
      if (imax != 1) { double t = A_(1, 1); A_(1, 1) = A_(imax, 1); A_(imax, 1) = t; }
      if (fabs(A_(1, 1)) >= sfmin) {
        double r = 1.0 / A_(1, 1);
        for (i = 2; i <= m; i++) A_(i, 1) = r * A_(i, 1);
      } else {
        for (i = 1; i <= m - 1; i++) A_(1 + i, 1) = A_(1 + i, 1) / A_(1, 1);
      }
    } else *info = 1;
  } else {
    int mn = m < n ? m : n, n1 = mn / 2, n2 = n - n1, iinfo, i;
    dgetrf2(m, n1, a, lda, ipiv, &iinfo);
    if (*info == 0 && iinfo > 0) *info = iinfo;
    dlaswp(n2, &A_(1, n1 + 1), lda, 1, n1, ipiv);
    dtrsm_llnu(n1, n2, a, lda, &A_(1, n1 + 1), lda);
    dgemm_nn_m1(m - n1, n2, n1, &A_(n1 + 1, 1), lda, &A_(1, n1 + 1), lda, &A_(n1 + 1, n1 + 1), lda);
    dgetrf2(m - n1, n2, &A_(n1 + 1, n1 + 1), lda, ipiv + n1, &iinfo);
    if (*info == 0 && iinfo > 0) *info = iinfo + n1;
    for (i = n1 + 1; i <= mn; i++) ipiv[i - 1] += n1;
    dlaswp(n1, a, lda, n1 + 1, mn, ipiv);
  }
}
void ox_dgetrf(int m, int n, double *a, int lda, int *ipiv, int *info) {
  /* dgetrf.f with NB = 64 (ilaenv.f:272-279) */
  const int nb = 64; int mn = m < n ? m : n, j, i, iinfo;
  *info = 0;
  if (m == 0 || n == 0) return;
  if (nb <= 1 || nb >= mn) { dgetrf2(m, n, a, lda, ipiv, info); return; }
  for (j = 1; j <= mn; j += nb) {
    int jb = (mn - j + 1 < nb) ? mn - j + 1 : nb;
    dgetrf2(m - j + 1, jb, &A_(j, j), lda, ipiv + j - 1, &iinfo);
    if (*info == 0 && iinfo > 0) *info = iinfo + j - 1;
    for (i = j; i <= ((m < j + jb - 1) ? m : j + jb - 1); i++) ipiv[i - 1] = j - 1 + ipiv[i - 1];
    dlaswp(j - 1, a, lda, j, j + jb - 1, ipiv);
    if (j + jb <= n) {
      dlaswp(n - j - jb + 1, &A_(1, j + jb), lda, j, j + jb - 1, ipiv);
      dtrsm_llnu(jb, n - j - jb + 1, &A_(j, j), lda, &A_(j, j + jb), lda);
      if (j + jb <= m)
        dgemm_nn_m1(m - j - jb + 1, n - j - jb + 1, jb, &A_(j + jb, j), lda, &A_(j, j + jb), lda,
                    &A_(j + jb, j + jb), lda);
    }
  }
}
void ox_dgetrs(int n, int nrhs, const double *a, int lda, const int *ipiv, double *b, int ldb) {
  /* dgetrs.f, TRANS='N' */
  dlaswp(nrhs, b, ldb, 1, n, ipiv);
  dtrsm_llnu(n, nrhs, a, lda, b, ldb);
  dtrsm_lunn(n, nrhs, a, lda, b, ldb);
}
#undef A_

/* -------------------------------------------------------------------------------------------- */
static int step_be(zone_t *z, int *inr) {
  /* xnet_integrate_be.F90:233-482 for one zone; *inr: 0 active on input; >0 its on convergence */
  const ox_net *net = z->net; const ox_ctl *ctl = z->ctl;
  int ny = z->ny, k, kit, iterate, iheat = ctl->iheat, rc_;
  double xtot_init, rdt, mult, s1;
  if (*inr != 0) return 0;
  iterate = 1;
  s1 = 0.0;
  for (k = 0; k < ny; k++) s1 = s1 + net->aa[k] * z->y[k];
  xtot_init = s1 + z->xext - 1.0;
  rdt = 1.0 / z->tdel; mult = -1.0;
  for (kit = 1; kit <= ctl->kitmx; kit++) {
    int rebuild = iterate && ((kit - 1) % ctl->ijac == 0);
    int eval_rates;
    double s2, s3, testc, testc2, xtot, testm, testn, toln;
    if (iheat > 0) eval_rates = iterate; else if (kit == 1) eval_rates = iterate && z->nh > 1; else eval_rates = 0;
    if (eval_rates) { rc_ = cross_sect(z); if (rc_) return rc_; }
    yderiv(z);
    if (rebuild) {
      jacobian_build(z, rdt, mult);
      ox_dgetrf(z->msize, z->msize, z->jac, z->msize, z->indx, &z->info);
    }
    if (kit > 1) {
      for (k = 0; k < ny; k++) z->yrhs[k] = (z->y[k] - z->yt[k]) * rdt + z->ydot[k];
      if (iheat > 0) z->t9rhs = (z->t9 - z->t9t) * rdt + z->t9dot;
    } else {
      for (k = 0; k < ny; k++) z->yrhs[k] = z->ydot[k];
      if (iheat > 0) z->t9rhs = z->t9dot;
    }
    for (k = 0; k < ny; k++) z->rhs[k] = z->yrhs[k];
    if (iheat > 0) z->rhs[ny] = z->t9rhs;
    ox_dgetrs(z->msize, 1, z->jac, z->msize, z->indx, z->rhs, z->msize);
    for (k = 0; k < ny; k++) z->dy[k] = z->rhs[k];
    if (iheat > 0) z->dt9 = z->rhs[ny];
    s1 = 0.0; s2 = 0.0; s3 = 0.0;
    for (k = 0; k < ny; k++) {
      double ytmp = z->yt[k] + z->dy[k];
      if (ytmp < ctl->ymin) { z->dy[k] = -z->yt[k]; z->yt[k] = 0.0; z->reldy[k] = 0.0; }
      else { z->yt[k] = ytmp; z->reldy[k] = fabs(z->dy[k] / z->yt[k]); }
      s1 = s1 + z->reldy[k];
      s2 = s2 + net->aa[k] * z->dy[k];
      s3 = s3 + net->aa[k] * z->yt[k];
    }
    testc = s1; testc2 = s2; (void)testc2;
    xtot = s3 + z->xext - 1.0;
    testm = xtot - xtot_init;
    if (iheat > 0) {
      z->t9t = z->t9t + z->dt9;
      if (fabs(z->t9t) > DBL_MIN) z->relt9 = fabs(z->dt9 / z->t9t); else z->relt9 = 0.0;
    } else z->relt9 = 0.0;
    if (ctl->iconvc != 0) { testn = testc; toln = ctl->tolc; } else { testn = testm; toln = ctl->tolm; }
    if (fabs(testn) <= toln && z->relt9 < ctl->tolt9) { *inr = kit; iterate = 0; } else iterate = 1;
    if (!iterate) break;
  }
  return 0;
}

static int timestep(zone_t *z, int kstep);

static int solve_be(zone_t *z, int kstep, int *its) {
  /* xnet_integrate_be.F90:32-231 for one zone */
  const ox_ctl *ctl = z->ctl; const int ktsmx = 10;
  int ny = z->ny, inr, mykts = 0, kts, k, lzstep, rc_;
  if (*its != 0) return 0;
  inr = 0; lzstep = 1;
  for (kts = 1; kts <= ktsmx; kts++) {
    rc_ = step_be(z, &inr); if (rc_) return rc_;
    if (inr == 0) {
      z->tdel = z->tdel / ctl->tdel_maxmult;
      z->tt = z->t + z->tdel;
      z->yet = z->ye;
      mykts = kts + 1;
      for (k = 0; k < ny; k++) z->yt[k] = z->y[k];
      z->kmon[1] = ctl->kitmx + 1; z->ktot[1] += ctl->kitmx + 1;
    } else if (inr > 0) {
      mykts = kts;
      z->kmon[1] = inr; z->ktot[1] += inr;
    }
    lzstep = (inr == 0);
    if (lzstep) {
      t9rhofind(z, z->tt, &z->ntt, &z->t9t, &z->rhot);
      if (ctl->iheat > 0) z->t9t = z->t9;
      if (kts == ktsmx - 1) {
        z->tdel = 0.0; /* tdelstart = 0 too */
        z->tdelstart = 0.0;
        rc_ = timestep(z, kstep); if (rc_) return rc_;
      }
    }
    if (!lzstep) break;
  }
  z->kmon[0] = mykts; z->ktot[0] += mykts;
  if (inr > 0) {
    z->tdel_next = z->tdel * ctl->tdel_maxmult;
    z->nto = z->nt; z->nt = z->ntt;
    z->to = z->t; z->t = z->tt;
    z->t9o = z->t9; z->t9 = z->t9t;
    z->rhoo = z->rho; z->rho = z->rhot;
    z->yeo = z->ye; z->ye = z->yet;
    for (k = 0; k < ny; k++) { z->yo[k] = z->y[k]; z->y[k] = z->yt[k]; }
  } else *its = 1;
  return 0;
}

static int timestep(zone_t *z, int kstep) {
  /* xnet_integrate.F90:13-268 for one zone (mask true) */
  const ox_ctl *ctl = z->ctl; const double changeth = 0.1;
  int ny = z->ny, k, mask_init, rc_;
  double changest, changey, changet9, tdel_init, tdel_stop, tdel_dy = 0.0, tdel_dt9 = 0.0, rtau_t9;
  (void)kstep;
  z->tdel_old = z->tdel;
  mask_init = (fabs(z->tdel_old) < DBL_MIN);
  if (mask_init) { rc_ = cross_sect(z); if (rc_) return rc_; yderiv(z); }
  tdel_stop = z->tstop - z->t;
  if (mask_init) {
    z->intso = 0;
    changest = (z->tdelstart < 0.0) ? 0.5 : 0.01;
    changey = changest * (0.1 < ctl->changemx ? 0.1 : ctl->changemx);
    changet9 = changest * (0.01 < ctl->changemxt ? 0.01 : ctl->changemxt);
    tdel_init = changest * fabs(z->tdelstart); if (tdel_stop < tdel_init) tdel_init = tdel_stop;
    if (z->nh > 2) z->tdel_next = 1.0e-4 * tdel_stop; else z->tdel_next = tdel_stop;
  } else { changey = ctl->changemx; changet9 = ctl->changemxt; tdel_init = 0.0; }
  if (z->tdel_old >= 0.0) {
    int imax = 0; double vmax;
    for (k = 0; k < ny; k++) {
      if (z->y[k] > ctl->ymin) {
        double yfloor = z->y[k] > ctl->yacc ? z->y[k] : ctl->yacc;
        if (mask_init) z->rtau[k] = fabs(z->ydot[k] / yfloor);
        else z->rtau[k] = fabs((z->y[k] - z->yo[k]) / yfloor) / z->tdel_old;
      } else z->rtau[k] = 0.0;
    }
    vmax = z->rtau[0];
    for (k = 1; k < ny; k++) if (z->rtau[k] > vmax) { vmax = z->rtau[k]; imax = k; }
    z->ints = imax + 1;
    if (fabs(z->rtau[imax]) > 0.0) tdel_dy = changey / z->rtau[imax]; else tdel_dy = z->tdel_next;
    if (ctl->iheat > 0) {
      if (mask_init) rtau_t9 = fabs(z->t9dot / z->t9);
      else rtau_t9 = fabs((z->t9 - z->t9o) / z->t9) / z->tdel_old;
    } else rtau_t9 = 0.0;
    if (rtau_t9 > 0.0) tdel_dt9 = changet9 / rtau_t9; else tdel_dt9 = z->tdel_next;
    {
      double m = tdel_stop;
      if (z->tdel_next < m) m = z->tdel_next;
      if (tdel_dy < m) m = tdel_dy;
      if (tdel_dt9 < m) m = tdel_dt9;
      z->tdel = m;
    }
    if (tdel_init > z->tdel) z->tdel = tdel_init;
  } else z->tdel = -z->tdel_old;
  { double v = z->t + z->tdel; z->tt = v < z->tstop ? v : z->tstop; }
  if (ctl->iheat == 0 && z->nh > 1) {
    int j, i; double dtherm;
    t9rhofind(z, z->tt, &z->ntt, &z->t9t, &z->rhot);
    if (z->ntt - 1 > z->nt) {
      for (j = z->nt; j <= z->ntt - 1; j++) {
        double t9j = z->t9h[j - 1], rhoj = rhoh_at(z, j);
        if ((t9j > z->t9 && t9j > z->t9t) || (t9j < z->t9 && t9j < z->t9t) ||
            (rhoj > z->rho && rhoj > z->rhot) || (rhoj < z->rho && rhoj < z->rhot)) {
          z->tdel = z->th[j - 1] - z->t; break;
        }
      }
      { double v = z->t + z->tdel; z->tt = v < z->tstop ? v : z->tstop; }
    }
    dtherm = 0.0;
    for (i = 1; i <= 10; i++) {
      t9rhofind(z, z->tt, &z->ntt, &z->t9t, &z->rhot);
      if (z->t9 > 0.0) dtherm = 10.0 * fabs(z->t9t - z->t9) / z->t9 + fabs(z->rhot - z->rho) / z->rho;
      if (dtherm < changeth) break;
      z->tdel = 0.5 * z->tdel;
      { double v = z->t + z->tdel; z->tt = v < z->tstop ? v : z->tstop; }
    }
  }
  return 0;
}


/* ------------------------------- BDF integrator (xnet_integrate_bdf.F90) ----------------------------------------
 * Per-zone restatement of solve_bdf and its helpers (the zone masks of the reference become plain conditions).
 * check_stability = .false. in the reference: bdf_stab / bdf_stab_detect are dead code and are not restated.
 * Real powers go through M_POW (libm, or the portable exp/log pair of the -DOX_PMATH build), integer powers through
 * the square-and-multiply loop of libgfortran's pow_r8_i4. */
#define BDF_MAX_ORDER 5
#define BDF_SKIP (-1)
#define BDF_SUCCESS 0
#define BDF_FAIL 1
static const int bdf_max_it_ts = 20, bdf_max_it_nr = 3, bdf_max_p_age = -1, bdf_max_j_age = -1, bdf_max_nerr_ts = 10,
                 bdf_max_ncf_ts = 10, bdf_nerr_ts_small = 2, bdf_nerr_ts_qreduce = 3, bdf_nstep_qwait = 10;
static const double bdf_eta_min = 0.1, bdf_eta_small = 0.2, bdf_eta_reduce = 0.25, bdf_eta_max = 10.0, bdf_eta_thresh = 1.5,
                    bdf_dgmaxp = 0.3, bdf_dgmaxj = 0.2, bdf_tol_nr = 0.1, bdf_cr_down = 0.3, bdf_cr_diverge = 2.0;
#define BZ(a, j, i) (a)[(size_t)(j) * z->neq + (i)]

static double bdf_powi(double a, int n) {
  double pw = 1.0, x = a; unsigned u = (unsigned)(n < 0 ? -n : n);
  if (n == 0) return 1.0;
  for (;;) { if (u & 1u) pw = pw * x; u >>= 1; if (u) x = x * x; else break; }
  return n < 0 ? 1.0 / pw : pw;
}
static double bdf_spacing(double x) { /* Fortran SPACING */
  int e; if (x == 0.0) return DBL_MIN; (void)frexp(fabs(x), &e); return ldexp(1.0, e - 53);
}
static int bdf_pascal(int i, int j) { /* binomial(i, j), i >= j (pascal_build :2014-2046) */
  static const int A[6][6] = {{1,0,0,0,0,0},{1,1,0,0,0,0},{1,2,1,0,0,0},{1,3,3,1,0,0},{1,4,6,4,1,0},{1,5,10,10,5,1}};
  return A[i][j];
}
static void bdf_error_weights(const zone_t *z, const double *y, double *wt) { /* :2048-2085 */
  const ox_ctl *ctl = z->ctl; int i, n = z->neq;
  if (ctl->iconvc == 0 || ctl->iconvc == 1) {
    for (i = 0; i < n; i++) {
      if (y[i] < ctl->ymin) wt[i] = 0.0;
      else if (y[i] < z->batol[i]) wt[i] = 1.0 / (z->brtol[i] * z->batol[i]);
      else wt[i] = 1.0 / (z->brtol[i] * fabs(y[i]));
    }
  } else {
    for (i = 0; i < n; i++) { double a = fabs(y[i]); if (a < ctl->ymin) a = ctl->ymin; wt[i] = 1.0 / (z->brtol[i] * a + z->batol[i]); }
  }
}
static double bdf_normw(const zone_t *z, const double *x, const double *wt) { /* :2087-2140 */
  int i, n = z->neq, ic = z->ctl->iconvc; double r = 0.0;
  if (ic == 0) { for (i = 0; i < n; i++) { double v = fabs(x[i] * wt[i]); if (v > r) r = v; } }
  else if (ic == 1) { for (i = 0; i < n; i++) r = r + fabs(x[i] * wt[i]); }
  else if (ic == 2) { for (i = 0; i < n; i++) { double v = x[i] * wt[i]; r = r + v * v; } r = sqrt(r); }
  else if (ic == 3) { for (i = 0; i < n; i++) { double v = x[i] * wt[i]; r = r + v * v; } r = sqrt(r / (double)n); }
  return r;
}
static void bdf_reset(zone_t *z) { /* :490-593 (tolerances: bdf_init :333-340) */
  const ox_ctl *ctl = z->ctl; int i, j, ny = z->ny, neq = z->neq;
  for (i = 0; i < ny; i++) { z->brtol[i] = ctl->tolc; z->batol[i] = ctl->yacc; }
  if (ctl->iheat > 0) { z->brtol[neq - 1] = ctl->tolt9; z->batol[neq - 1] = 1.0e-99; }
  z->retry_ts = 0; z->nit_nrslv = 0; z->nit_nr = 0; z->nit_ts = 0; z->ierr_nr = 0; z->ierr_ts = 0; z->nerr_ts = 0; z->ncf_ts = 0;
  z->bq = 1; z->deltaq = 0; z->j_age = bdf_max_j_age + 1; z->p_age = bdf_max_p_age + 1; z->q_age = 0;
  for (j = 0; j <= BDF_MAX_ORDER; j++) for (i = 0; i < neq; i++) { BZ(z->bz, j, i) = 0.0; BZ(z->bz0, j, i) = 0.0; BZ(z->bzt0, j, i) = 0.0; }
  for (i = 0; i < ny; i++) { BZ(z->bz, 0, i) = z->yt[i]; BZ(z->bz, 1, i) = z->ydot[i] * z->tdel; }
  if (ctl->iheat > 0) { BZ(z->bz, 0, neq - 1) = z->t9t; BZ(z->bz, 1, neq - 1) = z->t9dot * z->tdel; }
  for (i = 0; i <= BDF_MAX_ORDER; i++) { z->hvec[i] = z->tdel; z->lvec[i] = 0.0; }
  z->dt_scale = z->tdel; z->etaq[0] = z->etaq[1] = z->etaq[2] = 0.0; z->eta = 0.0; z->eta_next = bdf_eta_max;
  z->gam = z->tdel; z->gamhat = z->tdel; z->gamratio = 1.0;
  for (i = 0; i < neq; i++) { z->ewt[i] = 0.0; z->acor[i] = 0.0; z->acorp[i] = 0.0; }
  z->acnrm = 0.0; z->crate = 0.0; z->tq[0] = z->tq[1] = z->tq[2] = z->tq[3] = 0.0; z->tq2save = 0.0;
  for (i = 0; i < ny; i++) z->ydot0[i] = z->ydot[i];
  z->t9dot0 = z->t9dot;
}
static void bdf_adjust_order(zone_t *z) { /* :1844-1937 */
  int i, j, neq = z->neq, q = z->bq; double hsum, xi;
  if (z->deltaq == -1) {
    for (i = 0; i <= BDF_MAX_ORDER; i++) z->lvec[i] = 0.0;
    z->lvec[2] = 1.0; hsum = 0.0;
    for (j = 1; j <= q - 2; j++) {
      hsum = hsum + z->hvec[j]; xi = hsum / z->tdel;
      for (i = j + 2; i >= 2; i--) z->lvec[i] = z->lvec[i] * xi + z->lvec[i - 1];
    }
    for (j = 2; j <= q - 1; j++) for (i = 0; i < neq; i++) BZ(z->bz, j, i) = BZ(z->bz, j, i) - BZ(z->bz, q, i) * z->lvec[j];
    for (i = 0; i < neq; i++) BZ(z->bz, q, i) = 0.0;
    z->bq = q - 1; z->q_age = 0;
  } else if (z->deltaq == 1) {
    double alpha0 = -1.0, alpha1 = 1.0, prod = 1.0, xiold = 1.0, a1;
    for (i = 0; i <= BDF_MAX_ORDER; i++) z->lvec[i] = 0.0;
    z->lvec[2] = 1.0; hsum = z->tdel;
    for (j = 1; j <= q - 1; j++) {
      hsum = hsum + z->hvec[j + 1]; xi = hsum / z->tdel; prod = prod * xi;
      alpha0 = alpha0 - 1.0 / (double)(j + 1); alpha1 = alpha1 + 1.0 / xi;
      for (i = j + 2; i >= 2; i--) z->lvec[i] = z->lvec[i] * xiold + z->lvec[i - 1];
      xiold = xi;
    }
    a1 = -(alpha0 + alpha1) / prod;
    for (j = 2; j <= q; j++) for (i = 0; i < neq; i++) BZ(z->bz, j, i) = BZ(z->bz, j, i) + a1 * z->acor[i] * z->lvec[j];
    for (i = 0; i < neq; i++) BZ(z->bz, q + 1, i) = a1 * z->acor[i];
    z->bq = q + 1; z->q_age = 0;
  }
  z->deltaq = 0;
}
static void bdf_rescale(zone_t *z) { /* :1939-1978 (mask true) */
  int i, j, neq = z->neq;
  z->tdel = z->eta * z->dt_scale; z->hvec[0] = z->tdel; z->tt = z->t + z->tdel;
  for (j = 1; j <= z->bq; j++) { double f = bdf_powi(z->eta, j); for (i = 0; i < neq; i++) BZ(z->bz, j, i) = BZ(z->bz, j, i) * f; }
  z->dt_scale = z->tdel;
}
static void bdf_adjust(zone_t *z) { /* :595-629 */
  int rescale = 0;
  if (z->ierr_ts == BDF_SUCCESS) {
    if (fabs(z->tdel / z->tdel_old - 1.0) > DBL_EPSILON) { z->eta = z->tdel / z->tdel_old; rescale = 1; }
  }
  bdf_adjust_order(z);
  if (rescale) bdf_rescale(z);
}
static void bdf_update(zone_t *z, int kstep) { /* :688-831 (zone active) */
  int i, j, q = z->bq; double hsum, xi_inv = 1.0, xistar_inv = 1.0, alpha0 = -1.0, alpha0_hat = -1.0, a1, a2, a3, a4, a5, a6, c, cinv;
  double *l = z->lvec;
  l[0] = 1.0; l[1] = 1.0; for (i = 2; i <= BDF_MAX_ORDER; i++) l[i] = 0.0;
  hsum = z->tdel;
  if (q > 1) {
    for (j = 2; j <= q - 1; j++) {
      hsum = hsum + z->hvec[j - 1]; xi_inv = z->tdel / hsum; alpha0 = alpha0 - 1.0 / (double)j;
      for (i = j; i >= 1; i--) l[i] = l[i] + l[i - 1] * xi_inv;
    }
    alpha0 = alpha0 - 1.0 / (double)q;
    xistar_inv = -l[1] - alpha0;
    hsum = hsum + z->hvec[q - 1]; xi_inv = z->tdel / hsum;
    alpha0_hat = -l[1] - xi_inv;
    for (i = q; i >= 1; i--) l[i] = l[i] + l[i - 1] * xistar_inv;
  }
  a1 = 1.0 - alpha0_hat + alpha0; a2 = 1.0 + q * a1;
  z->tq[1] = fabs(a1 / (alpha0 * a2));                          /* tq(0) */
  z->tq[3] = fabs(a2 * xistar_inv / (l[q] * xi_inv));           /* tq(2) */
  if (q > 1) {
    c = xistar_inv / l[q]; a3 = alpha0 + 1.0 / q; a4 = alpha0_hat + xi_inv;
    z->tq[0] = fabs(c * (1.0 - a4 + a3) / a3);                  /* tq(-1) */
  } else z->tq[0] = 1.0;
  hsum = hsum + z->hvec[q]; xi_inv = z->tdel / hsum;
  a5 = alpha0 - 1.0 / (double)(q + 1); a6 = alpha0_hat - xi_inv;
  cinv = (1.0 - a6 + a5) / a2;
  z->tq[2] = fabs(cinv / (xi_inv * (double)(q + 2) * a5));      /* tq(1) */
  bdf_error_weights(z, &BZ(z->bz, 0, 0), z->ewt);
  z->gam = z->tdel / l[1];
  if (kstep > 1) z->gamratio = z->gam / z->gamhat; else { z->gamhat = z->gam; z->gamratio = 1.0; }
}
static void bdf_predict(zone_t *z) { /* :631-686 */
  int i, j, k, neq = z->neq, q = z->bq;
  for (j = 0; j <= BDF_MAX_ORDER; j++) for (i = 0; i < neq; i++) { BZ(z->bz0, j, i) = BZ(z->bz, j, i); BZ(z->bzt0, j, i) = BZ(z->bz, j, i); }
  for (i = 0; i <= q; i++) for (j = i + 1; j <= q; j++) for (k = 0; k < neq; k++)
    BZ(z->bzt0, i, k) = BZ(z->bzt0, i, k) + BZ(z->bz, j, k) * (double)bdf_pascal(j, i);
}
static int bdf_step(zone_t *z, int kstep, int active, int max_it_nrslv) { /* :833-1120 */
  const ox_net *net = z->net; const ox_ctl *ctl = z->ctl;
  int ny = z->ny, neq = z->neq, iheat = ctl->iheat, k, kit, refactor, rebuild, iterate, converged = 0, rc_;
  double delp = 0.0, del = 0.0, dcon = 0.0, diag = 1.0, mult, cstar;
  (void)kstep; (void)converged;
  if (active) {
    for (k = 0; k < ny; k++) { z->acor[k] = 0.0; z->yt[k] = BZ(z->bzt0, 0, k); }
    if (iheat > 0) { z->acor[neq - 1] = 0.0; z->t9t = BZ(z->bzt0, 0, neq - 1); }
    z->nit_nr = 0; z->nit_nrslv = 0;
  }
  refactor = active && (z->retry_ts || z->p_age > bdf_max_p_age || fabs(z->gamratio - 1.0) > bdf_dgmaxp);
  rebuild = refactor && z->j_age > bdf_max_j_age;
  iterate = active; z->crate = 1.0; mult = -z->gam;
  if (iheat == 0 && z->retry_ts) t9rhofind(z, z->tt, &z->ntt, &z->t9t, &z->rhot);
  for (kit = 1; kit <= max_it_nrslv; kit++) {
    int eval_rates;
    if (iterate) { z->nit_nr += 1; z->nit_nrslv += 1; }
    if (iheat > 0) eval_rates = iterate; else if (kit == 1) eval_rates = iterate && z->nh > 1; else eval_rates = 0;
    if (eval_rates) { rc_ = cross_sect(z); if (rc_) return rc_; }
    if (iterate) yderiv(z);
    if (rebuild) jacobian_build(z, diag, mult);           /* builds dydotdy and scales */
    else if (refactor) jacobian_scale(z, diag, mult);     /* refactor without rebuild: dydotdy of the last build */
    if (refactor) ox_dgetrf(z->msize, z->msize, z->jac, z->msize, z->indx, &z->info);
    if (refactor) { z->gamhat = z->gam; z->gamratio = 1.0; z->crate = 1.0; z->p_age = 0; refactor = 0; }
    if (rebuild) { z->j_age = 0; rebuild = 0; }
    if (iterate) {
      cstar = 2.0 / (1.0 + z->gamratio);
      for (k = 0; k < ny; k++) z->yrhs[k] = cstar * (z->gam * z->ydot[k] - BZ(z->bzt0, 1, k) / z->lvec[1] - z->acor[k]);
      if (iheat > 0) z->t9rhs = cstar * (z->gam * z->t9dot - BZ(z->bzt0, 1, neq - 1) / z->lvec[1] - z->acor[neq - 1]);
      for (k = 0; k < ny; k++) z->rhs[k] = z->yrhs[k];
      if (iheat > 0) z->rhs[ny] = z->t9rhs;
      ox_dgetrs(z->msize, 1, z->jac, z->msize, z->indx, z->rhs, z->msize);
      for (k = 0; k < ny; k++) z->dy[k] = z->rhs[k];
      if (iheat > 0) z->dt9 = z->rhs[ny];
      for (k = 0; k < ny; k++) {
        z->yt[k] = BZ(z->bzt0, 0, k) + z->acor[k] + z->dy[k];
        if (z->yt[k] < ctl->ymin) { z->yt[k] = 0.0; z->dy[k] = -BZ(z->bzt0, 0, k) - z->acor[k]; }
        else if (z->yt[k] * net->aa[k] > 1.0) { z->yt[k] = 1.0 / net->aa[k]; z->dy[k] = 1.0 / net->aa[k] - BZ(z->bzt0, 0, k) - z->acor[k]; }
        z->acor[k] = z->acor[k] + z->dy[k];
        z->dvec[k] = z->dy[k];
      }
      if (iheat > 0) {
        z->dvec[neq - 1] = z->dt9; z->acor[neq - 1] = z->acor[neq - 1] + z->dt9;
        z->t9t = BZ(z->bzt0, 0, neq - 1) + z->acor[neq - 1];
      }
      del = bdf_normw(z, z->dvec, z->ewt);
      if (z->nit_nr > 1) { double r = del / delp, c2 = bdf_cr_down * z->crate; z->crate = c2 > r ? c2 : r; }
      dcon = del * (z->crate < 1.0 ? z->crate : 1.0) * z->tq[1];
      converged = dcon < bdf_tol_nr;
      if (converged) {
        iterate = 0;
        if (z->nit_nr > 1) z->acnrm = bdf_normw(z, z->acor, z->ewt); else z->acnrm = del;
      } else if (z->nit_nrslv == max_it_nrslv) { iterate = 0; z->nit_nr = bdf_max_it_nr + 1; }
      else if (z->nit_nr == bdf_max_it_nr || (z->nit_nr > 1 && del > delp * bdf_cr_diverge)) {
        if (z->j_age > 0) {
          for (k = 0; k < ny; k++) { z->yt[k] = BZ(z->bzt0, 0, k); z->acor[k] = 0.0; }
          if (iheat > 0) { z->t9t = BZ(z->bzt0, 0, neq - 1); z->acor[neq - 1] = 0.0; }
          refactor = 1; rebuild = fabs(z->gamratio - 1.0) < bdf_dgmaxj; z->nit_nr = 0;
        } else { iterate = 0; z->nit_nr = bdf_max_it_nr + 1; }
      }
      delp = del;
    }
    if (!iterate) break;
  }
  return 0;
}
static void bdf_check(zone_t *z, int active) { /* :1122-1324 */
  const ox_ctl *ctl = z->ctl; const double biasq = 6.0, addon = 1.0e-6;
  int k, ny = z->ny, neq = z->neq, restore = 0, rescale = 0; double error;
  z->retry_ts = 0;
  if (active) {
    if (z->nit_nr > bdf_max_it_nr) {
      z->ncf_ts += 1; z->eta_next = 1.0;
      if (z->nit_ts >= bdf_max_it_ts) { z->ierr_nr = BDF_SKIP; z->ierr_ts = BDF_FAIL; }
      else if (z->tdel <= bdf_spacing(z->t)) { z->ierr_nr = BDF_SKIP; z->ierr_ts = BDF_FAIL; }
      else if (z->ncf_ts == bdf_max_ncf_ts) { z->ierr_nr = BDF_SKIP; z->ierr_ts = BDF_FAIL; }
      else {
        double e2 = bdf_spacing(z->t) / z->tdel;
        z->ierr_nr = BDF_FAIL; z->ierr_ts = BDF_SUCCESS; z->retry_ts = 1; restore = 1;
        z->eta = bdf_eta_reduce > e2 ? bdf_eta_reduce : e2; rescale = 1;
        z->kmon[1] = ctl->kitmx + 1;
      }
    } else {
      error = z->tq[1] * z->acnrm;
      if (error <= 1.0) { z->ierr_nr = BDF_SUCCESS; z->ierr_ts = BDF_SUCCESS; }
      else {
        z->nerr_ts += 1; z->eta_next = 1.0;
        if (z->nerr_ts == bdf_max_nerr_ts) { z->ierr_nr = BDF_SKIP; z->ierr_ts = BDF_FAIL; }
        else if (z->tdel <= bdf_spacing(z->t)) { z->ierr_nr = BDF_SKIP; z->ierr_ts = BDF_FAIL; }
        else if (z->nerr_ts > bdf_nerr_ts_qreduce) {
          double e2 = bdf_spacing(z->t) / z->tdel;
          z->ierr_nr = BDF_FAIL; z->ierr_ts = BDF_SUCCESS; z->retry_ts = 1;
          z->eta = bdf_eta_min > e2 ? bdf_eta_min : e2; rescale = 1;
          if (z->bq == 1) {
            z->q_age = z->bq - bdf_nstep_qwait;
            for (k = 0; k < ny; k++) { BZ(z->bz, 0, k) = z->y[k]; BZ(z->bz, 1, k) = z->ydot0[k] * z->tdel; }
            if (ctl->iheat > 0) { BZ(z->bz, 0, neq - 1) = z->t9; BZ(z->bz, 1, neq - 1) = z->t9dot0 * z->tdel; }
            restore = 0;
          } else { restore = 1; z->deltaq = -1; }
        } else {
          double e2 = bdf_spacing(z->t) / z->tdel, e;
          z->ierr_nr = BDF_FAIL; z->ierr_ts = BDF_SUCCESS; z->retry_ts = 1; restore = 1;
          e = 1.0 / (M_POW(biasq * error, 1.0 / (double)(z->bq + 1)) + addon);
          if (e2 > e) e = e2;                       /* max( eta_min, spacing/tdel, eta ) */
          if (bdf_eta_min > e) e = bdf_eta_min;
          if (e > bdf_eta_max) e = bdf_eta_max;
          if (z->nerr_ts >= bdf_nerr_ts_small && e > bdf_eta_small) e = bdf_eta_small;
          z->eta = e; rescale = 1;
        }
      }
    }
    z->ktot[1] += z->nit_nrslv;
    z->nit_ts += 1;
  }
  if (restore) { int i, j; for (j = 0; j <= BDF_MAX_ORDER; j++) for (i = 0; i < neq; i++) BZ(z->bz, j, i) = BZ(z->bz0, j, i); }
  bdf_adjust_order(z);
  if (rescale) bdf_rescale(z);
}
static void bdf_correct(zone_t *z) { /* :1326-1395 */
  int i, j, k, ny = z->ny, neq = z->neq;
  if (z->ierr_nr != BDF_SUCCESS) return;
  for (k = 0; k < ny; k++) BZ(z->bz, 0, k) = z->yt[k];
  if (z->ctl->iheat > 0) BZ(z->bz, 0, neq - 1) = z->t9t;
  for (j = 1; j <= z->bq; j++) for (i = 0; i < neq; i++) BZ(z->bz, j, i) = BZ(z->bzt0, j, i) + z->acor[i] * z->lvec[j];
  for (j = BDF_MAX_ORDER; j >= 1; j--) z->hvec[j] = z->hvec[j - 1];
  z->q_age += 1; z->j_age += 1; z->p_age += 1;
}
static void bdf_eta(zone_t *z) { /* :1397-1522 */
  const double biasqm1 = 6.0, biasq = 6.0, biasqp1 = 10.0, addon = 1.0e-6;
  int i, neq = z->neq, q = z->bq; double errq, c;
  if (z->ierr_nr != BDF_SUCCESS) return;
  if (z->eta_next > 1.0) {
    z->etaq[0] = z->etaq[1] = z->etaq[2] = 1.0;
    errq = z->tq[1] * z->acnrm;
    z->etaq[1] = 1.0 / (M_POW(biasq * errq, 1.0 / (double)(q + 1)) + addon);
    if (z->q_age > q) {
      if (q > 1) {
        errq = z->tq[0] * bdf_normw(z, &BZ(z->bz, q, 0), z->ewt);
        z->etaq[0] = 1.0 / (M_POW(biasqm1 * errq, 1.0 / (double)q) + addon);
      }
      if (q < BDF_MAX_ORDER) {
        c = (z->tq[3] / z->tq2save) * bdf_powi(z->hvec[1] / z->hvec[2], q + 1);
        for (i = 0; i < neq; i++) z->acorhat[i] = z->acor[i] - c * z->acorp[i];
        errq = z->tq[2] * bdf_normw(z, z->acorhat, z->ewt);
        z->etaq[2] = 1.0 / (M_POW(biasqp1 * errq, 1.0 / (double)(q + 2)) + addon);
      }
      z->q_age = q - 1;
    }
    z->eta = 1.0;
    if (z->etaq[1] > z->eta) { z->eta = z->etaq[1]; z->deltaq = 0; }
    if (z->etaq[0] > z->eta) { z->eta = z->etaq[0]; z->deltaq = -1; }
    if (z->etaq[2] > z->eta) { z->eta = z->etaq[2]; z->deltaq = 1; }
    if (z->eta <= bdf_eta_thresh) { z->eta = 1.0; z->deltaq = 0; }
    if (z->eta > bdf_eta_max) z->eta = bdf_eta_max;
  } else {
    if (q - 1 < z->q_age) z->q_age = q - 1;
  }
  for (i = 0; i < neq; i++) z->acorp[i] = z->acor[i];
  z->tq2save = z->tq[3];
}
static int solve_bdf(zone_t *z, int kstep, int *its) { /* :155-310 */
  const ox_ctl *ctl = z->ctl; int ny = z->ny, neq = z->neq, k, kts, rc_;
  const int max_it_nrslv = ctl->kitmx > bdf_max_it_nr ? ctl->kitmx : bdf_max_it_nr;
  const int first = !z->bdf_started;       /* kstep == 1 of this zone's run */
  if (first) { bdf_reset(z); z->bdf_started = 1; }
  z->ierr_ts = *its;
  z->ierr_nr = (z->ierr_ts != BDF_SUCCESS) ? BDF_SKIP : BDF_FAIL;
  for (k = 0; k < ny; k++) z->ydot0[k] = z->ydot[k];
  if (ctl->iheat > 0) z->t9dot0 = z->t9dot;
  z->eta_next = bdf_eta_max; z->nit_ts = 0; z->nerr_ts = 0; z->ncf_ts = 0;
  if (!first) bdf_adjust(z);
  for (kts = 1; kts <= bdf_max_it_ts; kts++) {
    const int active = z->ierr_nr == BDF_FAIL;
    if (active) bdf_update(z, first ? 1 : 2);
    if (active) bdf_predict(z);
    rc_ = bdf_step(z, kstep, active, max_it_nrslv); if (rc_) return rc_;
    bdf_check(z, active);
    if (z->ierr_nr != BDF_FAIL) break;
  }
  bdf_correct(z);
  bdf_eta(z);
  z->tdel_next = z->eta * z->tdel;
  if (z->ierr_nr == BDF_SUCCESS) {
    z->ierr_ts = BDF_SUCCESS;
    for (k = 0; k < ny; k++) { z->yt[k] = BZ(z->bz, 0, k); z->yo[k] = z->y[k]; z->y[k] = z->yt[k]; }
    if (ctl->iheat > 0) z->t9t = BZ(z->bz, 0, neq - 1);
    z->nto = z->nt; z->nt = z->ntt; z->to = z->t; z->t = z->tt; z->t9o = z->t9; z->t9 = z->t9t;
    z->rhoo = z->rho; z->rho = z->rhot; z->yeo = z->ye; z->ye = z->yet;
  } else if (z->ierr_nr == BDF_FAIL || z->ierr_ts == BDF_FAIL) z->ierr_ts = BDF_FAIL;
  z->kmon[0] = z->nit_ts; z->ktot[0] += z->nit_ts; z->kmon[1] = z->nit_nrslv;
  *its = z->ierr_ts;
  return 0;
}

static void benuc(const zone_t *z, const double *y, double *enb, double *enm) {
  /* xnet_data.F90:157-175 */
  const ox_net *net = z->net; double ztot = 0.0, btot = 0.0; int k;
  for (k = 0; k < z->ny; k++) { ztot = ztot + y[k] * net->zz[k]; btot = btot + y[k] * net->be[k]; }
  *enb = btot;
  *enm = net->mex_p * ztot + net->mex_n * (1.0 - ztot) - btot;
  *enb = epmev * avn * *enb;
  *enm = epmev * avn * *enm;
}

static void enudot(const zone_t *z, const double *y, double *sqnu) {
  /* xnet_data.F90:990-1020 */
  const ox_net *net = z->net; int k;
  if (net->nffn > 0) {
    double s = 0.0;
    for (k = 0; k < net->nr[0]; k++) s = s + y[net->ni[0][k * 6] - 1] * z->qkffn[net->iffn[k]];
    *sqnu = epmev * avn * s;
  }
}

static double *dalloc(size_t n) { double *p = (double *)calloc(n ? n : 1, sizeof(double)); return p; }

static zone_t *zone_new(const ox_net *net, const ox_ctl *ctl) {
  zone_t *z = (zone_t *)calloc(1, sizeof(zone_t)); int c, r, j, ny = net->ny;
  z->net = net; z->ctl = ctl; z->ny = ny; z->msize = ny + (ctl->iheat > 0 ? 1 : 0);
  z->y = dalloc(ny); z->yo = dalloc(ny); z->yt = dalloc(ny); z->ydot = dalloc(ny);
  z->yrhs = dalloc(ny); z->dy = dalloc(ny); z->reldy = dalloc(ny); z->rtau = dalloc(ny);
  for (c = 0; c < 4; c++) {
    z->csect[c] = dalloc(net->nr[c]); z->dcsect[c] = dalloc(net->nr[c]);
    z->h[c] = dalloc(net->nr[c]); z->dh[c] = dalloc(net->nr[c]); z->b[c] = dalloc(net->nan[c]);
    for (r = 0; r <= c; r++) {
      z->n_n[c][r] = (int *)calloc(net->nan[c] ? net->nan[c] : 1, sizeof(int));
      for (j = 0; j < net->nan[c]; j++) z->n_n[c][r][j] = net->ni[c][(net->mu[c][j] - 1) * 6 + r];
    }
  }
  z->gg = dalloc(ny + 1); z->dlngdt9 = dalloc(ny + 1);
  z->rffn = dalloc(net->nffn); z->dlnrffndt9 = dalloc(net->nffn); z->qkffn = dalloc(net->nffn + 1);
  z->phasei = dalloc(net->nffn); z->dphaseidt9 = dalloc(net->nffn);
  z->dydotdy = dalloc((size_t)z->msize * z->msize); z->jac = dalloc((size_t)z->msize * z->msize);
  z->rhs = dalloc(z->msize); z->indx = (int *)calloc(z->msize, sizeof(int));
  z->neq = z->msize;
  z->bz = dalloc((size_t)6 * z->neq); z->bz0 = dalloc((size_t)6 * z->neq); z->bzt0 = dalloc((size_t)6 * z->neq);
  z->ewt = dalloc(z->neq); z->acor = dalloc(z->neq); z->acorp = dalloc(z->neq); z->acorhat = dalloc(z->neq);
  z->dvec = dalloc(z->neq); z->ydot0 = dalloc(ny); z->brtol = dalloc(z->neq); z->batol = dalloc(z->neq);
  return z;
}
static void zone_free(zone_t *z) {
  int c, r;
  free(z->y); free(z->yo); free(z->yt); free(z->ydot); free(z->yrhs); free(z->dy); free(z->reldy); free(z->rtau);
  for (c = 0; c < 4; c++) {
    free(z->csect[c]); free(z->dcsect[c]); free(z->h[c]); free(z->dh[c]); free(z->b[c]);
    for (r = 0; r <= c; r++) free(z->n_n[c][r]);
  }
  free(z->gg); free(z->dlngdt9); free(z->rffn); free(z->dlnrffndt9); free(z->qkffn); free(z->phasei); free(z->dphaseidt9);
  free(z->bz); free(z->bz0); free(z->bzt0); free(z->ewt); free(z->acor); free(z->acorp); free(z->acorhat); free(z->dvec);
  free(z->ydot0); free(z->brtol); free(z->batol);
  free(z->dydotdy); free(z->jac); free(z->rhs); free(z->indx); free(z);
}
static void zone_reset_scratch(zone_t *z) {
  const ox_net *net = z->net; int c;
  for (c = 0; c < 4; c++) {
    memset(z->csect[c], 0, sizeof(double) * net->nr[c]); memset(z->dcsect[c], 0, sizeof(double) * net->nr[c]);
    memset(z->h[c], 0, sizeof(double) * net->nr[c]); memset(z->dh[c], 0, sizeof(double) * net->nr[c]);
    memset(z->b[c], 0, sizeof(double) * net->nan[c]);
  }
  memset(z->qkffn, 0, sizeof(double) * (net->nffn + 1));
  memset(z->rffn, 0, sizeof(double) * (net->nffn ? net->nffn : 1));
  memset(z->ydot, 0, sizeof(double) * z->ny);
  z->cv = z->etae = z->detaedt9 = 0.0; z->t9dot = 0.0;
}

static void flux(const zone_t *z, double *flx) {
  /* xnet_flux.F90:104-131: flx(ifl) += dcflx(dir,ifl) * sign(csect * prod yt, ifl_n(i)), classes and reactions in order */
  const ox_net *net = z->net; int c, i, m;
  for (m = 0; m < net->mflx; m++) flx[m] = 0.0;
  for (c = 0; c < 4; c++) {
    for (i = 0; i < net->nr[c]; i++) {
      const int *ni = net->ni[c] + (size_t)i * 6;
      int f = net->ifl[c][i], ifl = f < 0 ? -f : f, idc = f < 0 ? 1 : 0;
      double flt;
      if (c == 0) flt = z->csect[0][i] * z->yt[ni[0] - 1];
      else if (c == 1) flt = z->csect[1][i] * z->yt[ni[0] - 1] * z->yt[ni[1] - 1];
      else if (c == 2) flt = z->csect[2][i] * ((z->yt[ni[0] - 1] * z->yt[ni[1] - 1]) * z->yt[ni[2] - 1]);
      else flt = z->csect[3][i] * (((z->yt[ni[0] - 1] * z->yt[ni[1] - 1]) * z->yt[ni[2] - 1]) * z->yt[ni[3] - 1]);
      if (ifl == 0) continue;
      flx[ifl - 1] = flx[ifl - 1] + net->dcflx[2 * (ifl - 1) + idc] * (f < 0 ? -fabs(flt) : fabs(flt));
    }
  }
}

static void hist_put(const ox_out *out, int iz, int ny, int kstep, const zone_t *z, double edot, double sqnu) {
  int n;
  if (out->hist_cap <= 0) return;
  n = out->hist_n[iz];
  if (n < out->hist_cap) {
    size_t o = (size_t)iz * out->hist_cap + n;
    out->hist_kstep[o] = kstep; out->hist_kmon[2 * o] = z->kmon[0]; out->hist_kmon[2 * o + 1] = z->kmon[1];
    out->hist_t[o] = z->t; out->hist_t9[o] = z->t9; out->hist_rho[o] = z->rho; out->hist_tdel[o] = z->tdel;
    out->hist_edot[o] = edot; out->hist_sqnu[o] = sqnu;
    memcpy(out->hist_y + o * ny, z->y, sizeof(double) * ny);
    if (out->hist_flx && z->net->mflx > 0) {   /* ts_output: flux() for kstep > 0, zeros in the initial record */
      double *f = out->hist_flx + o * (size_t)z->net->mflx;
      if (kstep > 0) flux(z, f); else memset(f, 0, sizeof(double) * z->net->mflx);
    }
  }
  out->hist_n[iz] = n + 1;
}

static int full_net_zone(zone_t *z, const ox_zones *zn, int iz, ox_out *out) {
  /* net.F90:246-272 (zone load) + xnet_evolve.F90:19-269 */
  const ox_ctl *ctl = z->ctl; int ny = z->ny, fx = zn->zfix[iz], k, kstep, its, lzsolve, mykstep = 0, rc_ = 0, kstep0 = 0;
  double enb, enm, enold, edot, sqnu;
  zone_reset_scratch(z);
  z->nh = zn->nh[fx]; z->th = zn->th + zn->off[fx]; z->t9h = zn->t9h + zn->off[fx]; z->rhoh = zn->rhoh + zn->off[fx];
  z->has_scale = zn->rho_scale != NULL; z->rho_scale = z->has_scale ? zn->rho_scale[iz] : 1.0;
  z->tstart = zn->tstart[fx]; z->tstop = zn->tstop[fx]; z->tdelstart = zn->tdelstart[fx];
  z->xext = zn->xext[fx]; z->aext = zn->aext[fx]; z->zext = zn->zext[fx];
  z->tdel = 0.0; z->nt = zn->nstart[fx]; z->t = z->tstart; z->t9 = zn->t9start[fx];
  z->rho = z->has_scale ? zn->rhostart[fx] * z->rho_scale : zn->rhostart[fx];
  z->ye = zn->yestart[fx];
  for (k = 0; k < ny; k++) z->y[k] = zn->ystart[(size_t)fx * ny + k];
  its = 0; lzsolve = 1; z->bdf_started = 0;
  for (k = 0; k < 5; k++) { z->kmon[k] = 0; z->ktot[k] = 0; }
  z->ints = 0; z->intso = 0;
  if (out->hist_cap > 0) out->hist_n[iz] = 0;
  if (zn->rs_t != NULL) {
    /* restart from a recorded accepted step (state after solve_be's roll, xnet_integrate_be.F90:181-196) */
    double t9x, rhox;
    kstep0 = zn->rs_kstep0[iz];
    z->t = zn->rs_t[iz]; z->tdel = zn->rs_tdel[iz]; z->tdel_next = z->tdel * ctl->tdel_maxmult;
    z->t9 = zn->rs_t9[iz]; z->rho = zn->rs_rho[iz]; z->t9o = zn->rs_t9o[iz];
    for (k = 0; k < ny; k++) { z->y[k] = zn->rs_y[(size_t)iz * ny + k]; z->yo[k] = zn->rs_yo[(size_t)iz * ny + k]; z->yt[k] = z->y[k]; }
    t9rhofind(z, z->t, &z->nt, &t9x, &rhox);
    z->tdel_old = z->tdel; z->nto = z->nt; z->ntt = z->nt; z->to = z->t; z->tt = z->t;
    z->t9t = z->t9; z->rhoo = z->rho; z->rhot = z->rho; z->yeo = z->ye; z->yet = z->ye;
    benuc(z, z->y, &enb, &enm);
    enold = enm; edot = 0.0; sqnu = 0.0;
    /* constant thermodynamic conditions without self-heating: the reference evaluates the rates once, in the first
     * timestep call (tdel = 0, xnet_integrate.F90:58-62), and step_be never again (xnet_integrate_be.F90:307-313);
     * a restarted zone has tdel != 0, so the rates of its state are evaluated here */
    if (z->nh == 1 && ctl->iheat == 0) { rc_ = cross_sect(z); if (rc_) return rc_; }
  } else {
  benuc(z, z->y, &enb, &enm);
  z->tdel_old = z->tdel; z->tdel_next = z->tdel; z->nto = z->nt; z->ntt = z->nt; z->to = z->t; z->tt = z->t;
  z->t9o = z->t9; z->t9t = z->t9; z->rhoo = z->rho; z->rhot = z->rho; z->yeo = z->ye; z->yet = z->ye;
  for (k = 0; k < ny; k++) { z->yo[k] = z->y[k]; z->yt[k] = z->y[k]; }
  enold = 0.0; edot = 0.0; sqnu = 0.0;
  hist_put(out, iz, ny, 0, z, edot, sqnu);
  }
  for (kstep = kstep0 + 1; kstep <= kstep0 + ctl->kstmx; kstep++) {
    int lzoutput;
    if (lzsolve) { rc_ = timestep(z, kstep); if (rc_) break; }
    if (ctl->isolv == 3) rc_ = solve_bdf(z, kstep, &its); else rc_ = solve_be(z, kstep, &its);
    if (rc_) break;
    lzoutput = 1;
    if (its == 0) {
      enold = enm;
      if (z->t >= z->tstop) { mykstep = kstep; its = -1; lzsolve = 0; }
    } else if (its == 1) { its = 2; lzsolve = 0; lzoutput = 0; }
    else { lzsolve = 0; lzoutput = 0; }
    if (lzoutput) {
      benuc(z, z->yt, &enb, &enm);
      enudot(z, z->yt, &sqnu);
      edot = -(enm - enold) / z->tdel;
      hist_put(out, iz, ny, kstep, z, edot, sqnu);
    }
    if (!lzsolve) break;
  }
  for (k = 0; k < ny; k++) out->y[(size_t)iz * ny + k] = z->y[k];
  out->t[iz] = z->t; out->t9[iz] = z->t9; out->rho[iz] = z->rho; out->tdel[iz] = z->tdel;
  out->kstep[iz] = mykstep; out->status[iz] = its;
  for (k = 0; k < 5; k++) out->ktot[(size_t)iz * 5 + k] = z->ktot[k];
  return rc_;
}

int ox_full_net(const ox_net *net, const ox_ctl *ctl, const ox_zones *zn, ox_out *out, int nthreads) {
  int rc_all = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    zone_t *z = zone_new(net, ctl);
    int iz;
#pragma omp for schedule(dynamic, 1)
    for (iz = 0; iz < zn->nzone; iz++) {
      int r = full_net_zone(z, zn, iz, out);
      if (r) {
#pragma omp critical
        rc_all = r;
      }
    }
    zone_free(z);
  }
  return rc_all;
}

int ox_eval_point(const ox_net *net, const ox_ctl *ctl, double t9, double rho, const double *y,
                  double xext, double aext, double zext, double diag, double mult,
                  double *csect, double *ydot, double *jac, double *ye_out) {
  zone_t *z = zone_new(net, ctl); int k, c, o = 0, rc_;
  zone_reset_scratch(z);
  z->t9t = t9; z->rhot = rho; z->xext = xext; z->aext = aext; z->zext = zext; z->nh = 2;
  for (k = 0; k < z->ny; k++) { z->yt[k] = y[k]; z->y[k] = y[k]; }
  rc_ = cross_sect(z);
  if (!rc_) {
    yderiv(z);
    jacobian_build(z, diag, mult);
    for (c = 0; c < 4; c++) { memcpy(csect + o, z->csect[c], sizeof(double) * net->nr[c]); o += net->nr[c]; }
    memcpy(ydot, z->ydot, sizeof(double) * z->ny);
    memcpy(jac, z->jac, sizeof(double) * z->msize * z->msize);
    *ye_out = z->yet;
  }
  zone_free(z);
  return rc_;
}

/* Same chain with the self-heating outputs (iheat > 0: msize = ny+1): additionally dcsect (concatenated classes),
 * eos[3] = cv, etae, detaedt9 (update_eos inside cross_sect) and t9dot (yderiv, xnet_integrate.F90:471-487).
 * jac is msize x msize. */
int ox_eval_point_heat(const ox_net *net, const ox_ctl *ctl, double t9, double rho, const double *y,
                       double xext, double aext, double zext, double diag, double mult,
                       double *csect, double *dcsect, double *ydot, double *jac, double *ye_out, double *eos,
                       double *t9dot) {
  zone_t *z = zone_new(net, ctl); int k, c, o = 0, rc_;
  zone_reset_scratch(z);
  z->t9t = t9; z->rhot = rho; z->xext = xext; z->aext = aext; z->zext = zext; z->nh = 2;
  for (k = 0; k < z->ny; k++) { z->yt[k] = y[k]; z->y[k] = y[k]; }
  rc_ = cross_sect(z);
  if (!rc_) {
    yderiv(z);
    jacobian_build(z, diag, mult);
    for (c = 0; c < 4; c++) {
      memcpy(csect + o, z->csect[c], sizeof(double) * net->nr[c]);
      memcpy(dcsect + o, z->dcsect[c], sizeof(double) * net->nr[c]);
      o += net->nr[c];
    }
    memcpy(ydot, z->ydot, sizeof(double) * z->ny);
    memcpy(jac, z->jac, sizeof(double) * z->msize * z->msize);
    *ye_out = z->yet; eos[0] = z->cv; eos[1] = z->etae; eos[2] = z->detaedt9; *t9dot = z->t9dot;
  }
  zone_free(z);
  return rc_;
}

int ox_flux_point(const ox_net *net, const ox_ctl *ctl, double t9, double rho, const double *y,
                  double xext, double aext, double zext, double *flx) {
  zone_t *z = zone_new(net, ctl); int k, rc_;
  zone_reset_scratch(z);
  z->t9t = t9; z->rhot = rho; z->xext = xext; z->aext = aext; z->zext = zext; z->nh = 2;
  for (k = 0; k < z->ny; k++) { z->yt[k] = y[k]; z->y[k] = y[k]; }
  rc_ = cross_sect(z);
  if (!rc_) flux(z, flx);
  zone_free(z);
  return rc_;
}

/* One call of step_be (xnet_integrate_be.F90:233-482) on an explicit state: y, tdel, t9 at the start of the step;
 * yt/t9t first guess in, last iterate out; nh > 1 makes the rates be evaluated at kit = 1.  Returns inr through
 * *inr_out (0 = no convergence), ydot/t9dot of the last evaluation, kcount[5] = counter increments. */
int ox_step_be_point(const ox_net *net, const ox_ctl *ctl, const double *y, double *yt, double tdel, double t9,
                     double *t9t, double rhot, int nh, double xext, double aext, double zext,
                     int *inr_out, double *ydot, double *t9dot, int *kcount) {
  zone_t *z = zone_new(net, ctl); int k, rc_, inr = 0;
  zone_reset_scratch(z);
  z->t9 = t9; z->t9t = *t9t; z->rhot = rhot; z->tdel = tdel; z->nh = nh;
  z->xext = xext; z->aext = aext; z->zext = zext;
  for (k = 0; k < 5; k++) z->ktot[k] = 0;
  for (k = 0; k < z->ny; k++) { z->y[k] = y[k]; z->yt[k] = yt[k]; }
  rc_ = step_be(z, &inr);
  if (!rc_) {
    for (k = 0; k < z->ny; k++) { yt[k] = z->yt[k]; ydot[k] = z->ydot[k]; }
    *t9t = z->t9t; *t9dot = z->t9dot; *inr_out = inr;
    for (k = 0; k < 5; k++) kcount[k] = z->ktot[k];
  }
  zone_free(z);
  return rc_;
}
