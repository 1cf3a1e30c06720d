/* Oracle (TEST INFRASTRUCTURE ONLY): CPU restatement of XNet's batched implicit step.
 * Nothing in the product (xnet_b200/) may include or link this.  See xnet_oracle.c. */
#ifndef XNET_ORACLE_H
#define XNET_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int ny, nffn, nnnu;
  int nr[4], nan[4];
  const double *rc[4];   /* [nr][7]   == Fortran rc_n(7,nr)                        */
  const int *ni[4];      /* [nr][6]   == n_ni(:,k); 1-based species, 0 = blank    */
  const int *iwk[4];     /* [nr]                                                   */
  const int *irev[4];    /* [nr]                                                   */
  const int *iffn;       /* [nr1]     1-based FFN table index (0 = none)           */
  const int *la, *le;    /* [4][ny]   la[c*ny+i], 1-based inclusive extents        */
  const int *mu[4];      /* [nan]     1-based reaction index                       */
  const double *a[4];    /* [nan]                                                  */
  const double *aa, *zz, *mex, *be; /* [ny] */
  double mex_n, mex_p;
  const double *g;       /* [ny][24]  == Fortran g(24,ny)                          */
  const double *t9i;     /* [24]                                                   */
  const double *ffnsum, *ffnenu, *ffn_ec, *ffn_beta; /* [nffn][143] */
  const double *ffn_qval; /* [nffn] */
  const int *has_logft;   /* [nffn] */
  /* Helmholtz table (only when iscrn>0 or iheat>0), see ox_eos.c; NULL otherwise */
  const double *helm;     /* 13 planes [13][201][541] : f,fd,ft,fdd,ftt,fdt,fddt,fdtt,fddtt,ef,efd,eft,efdt */
  /* reaction-pair (flux) matching, xnet_preprocess.F90:108-289 + xnet_flux.F90:20-65; mflx = 0: no flux output */
  int mflx;
  const int *ifl[4];      /* [nr] signed 1-based flux-set index (+ forward, - reverse)                  */
  const double *dcflx;    /* [mflx][2] double-counting factor of the forward / reverse direction         */
} ox_net;

typedef struct {
  int iweak0, iscrn, iheat, isolv, kstmx, kitmx, ijac, iconvc;
  double changemx, changemxt, yacc, tolm, tolc, tolt9, ymin, tdel_maxmult;
} ox_ctl;

/* K distinct input fixtures (trajectory + initial abundances); zone z uses fixture zfix[z]. */
typedef struct {
  int nfix;
  const int *nh;          /* [nfix] trajectory points                               */
  const int *off;         /* [nfix] offset of fixture's trajectory in th/t9h/rhoh   */
  const double *th, *t9h, *rhoh;
  const double *tstart, *tstop, *tdelstart;      /* [nfix] */
  const int *nstart;                             /* [nfix] */
  const double *t9start, *rhostart, *yestart;    /* [nfix] */
  const double *ystart;                          /* [nfix][ny] */
  const double *xext, *aext, *zext;              /* [nfix] */
  int nzone;
  const int *zfix;        /* [nzone] */
  const double *rho_scale;/* [nzone] or NULL: rhoh multiplied by this (decorrelated variant) */
  /* optional restart state (all NULL = start from the fixture's initial conditions).  Per zone:
   * state after an accepted step, as recorded in a history: the run continues with step
   * kstep0+1 and executes at most ctl->kstmx further steps. */
  const double *rs_t, *rs_tdel, *rs_t9, *rs_rho, *rs_t9o; /* [nzone] */
  const double *rs_y, *rs_yo;                              /* [nzone][ny] */
  const int *rs_kstep0;                                    /* [nzone] */
} ox_zones;

typedef struct {
  double *y;        /* [nzone][ny] final abundances                                */
  double *t, *t9, *rho, *tdel; /* [nzone] final                                    */
  int *kstep;       /* [nzone] step at which the zone finished (mykstep)           */
  int *ktot;        /* [nzone][5] TS / NR / Jac / Deriv / CrossSect counters       */
  int *status;      /* [nzone] its: -1 finished, 2 failed, 0 ran out of kstmx      */
  /* optional per-step history (ts_output records); hist_cap = 0 disables */
  int hist_cap;
  int *hist_n;      /* [nzone] records written (incl. kstep=0)                     */
  int *hist_kstep;  /* [nzone][cap] */
  int *hist_kmon;   /* [nzone][cap][2] */
  double *hist_t, *hist_t9, *hist_rho, *hist_tdel, *hist_edot, *hist_sqnu; /* [nzone][cap] */
  double *hist_y;   /* [nzone][cap][ny] */
  double *hist_flx; /* [nzone][cap][mflx] reaction fluxes of the record (xnet_flux.F90:67-150); NULL = skip */
} ox_out;

/* full_net for all zones, OpenMP over zones (nthreads<=0: default).  Returns 0. */
int ox_full_net(const ox_net *net, const ox_ctl *ctl, const ox_zones *zn, ox_out *out, int nthreads);

/* LAPACK 3.8.0 restatement, column-major, 1-based ipiv (for unit tests vs scipy) */
void ox_dgetrf(int m, int n, double *a, int lda, int *ipiv, int *info);
void ox_dgetrs(int n, int nrhs, const double *a, int lda, const int *ipiv, double *b, int ldb);

/* One evaluation of the rate/derivative/Jacobian chain for a single state (kernel-level parity
 * checks): inputs t9,rho,y[ny]; outputs csect (concatenated classes), ydot[ny], jac[ny*ny]
 * (col-major, jac = mult*dydotdy^T + diag, no T9 row: iheat must be 0), ye. */
int ox_eval_point(const ox_net *net, const ox_ctl *ctl, double t9, double rho, const double *y,
                  double xext, double aext, double zext, double diag, double mult,
                  double *csect, double *ydot, double *jac, double *ye_out);

/* ox_eval_point with the self-heating outputs (jac is msize x msize, msize = ny + (iheat>0)) */
int ox_eval_point_heat(const ox_net *net, const ox_ctl *ctl, double t9, double rho, const double *y,
                       double xext, double aext, double zext, double diag, double mult,
                       double *csect, double *dcsect, double *ydot, double *jac, double *ye_out, double *eos,
                       double *t9dot);
/* flux (xnet_flux.F90:67-150) of one state: cross_sect at (t9, rho, y), then the matched-pair fluxes flx[mflx] */
int ox_flux_point(const ox_net *net, const ox_ctl *ctl, double t9, double rho, const double *y,
                  double xext, double aext, double zext, double *flx);
/* one step_be call (Newton loop of one backward-Euler step) on an explicit state */
int ox_step_be_point(const ox_net *net, const ox_ctl *ctl, const double *y, double *yt, double tdel, double t9,
                     double *t9t, double rhot, int nh, double xext, double aext, double zext,
                     int *inr_out, double *ydot, double *t9dot, int *kcount);

#ifdef __cplusplus
}
#endif
#endif
