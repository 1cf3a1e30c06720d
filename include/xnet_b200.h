/* xnet_b200 -- C ABI of the B200-native batched per-zone implicit step for XNet.
 *
 * This is the drop-in boundary: everything a Fortran `xnet_jacobian_b200.F90` /
 * `xnet_integrate_b200.F90` shim (iso_c_binding) binds.  Plain C types only;
 * every pointer is a HOST pointer to caller-owned memory unless the name starts with `d_`.
 * Fortran arrays are passed as they lie in memory (column-major, zone index last), Fortran
 * LOGICAL(4) masks as int32 (non-zero = .true.), species/reaction indices stay 1-based.
 *
 * Return value of every function: 0 = ok; <0 = error (message via xb_last_error()); numerical
 * singularity is NOT an error (reported per zone in `info`, LAPACK meaning).  There is no CPU
 * fallback anywhere: without a usable CUDA device every compute entry point returns XB_ERR_CUDA.
 *
 * Reference interfaces replaced (paths relative to the XNet tree, /root/reference):
 *   xb_net_from_arrays    <- read_reaction_data / read_nuclear_data module state
 *                            (source/xnet_data.F90:391-513, 654-956)
 *   xb_net_load           <- net_preprocess + read_* from files (source/xnet_preprocess.F90:505-1011)
 *   xb_net_load_helm      <- actual_eos_init (tools/starkiller-helmholtz/actual_eos.F90:1520-1700): reads
 *                            helm_table.dat (ASCII, reference format) or the packed helm_table.bin
 *   xb_ctx_create         <- read_jacobian_data (source/xnet_jacobian_dense.F90:54-119), be_init
 *                            (source/xnet_integrate_be.F90:484-551)
 *   xb_ctx_destroy        <- jacobian_finalize (source/xnet_jacobian_dense.F90:121-134)
 *   xb_cross_sect         <- cross_sect (source/xnet_integrate.F90:569-984)
 *   xb_yderiv             <- yderiv (source/xnet_integrate.F90:364-567)
 *   xb_jacobian_build     <- jacobian_build (source/xnet_jacobian_dense.F90:257-434)
 *   xb_jacobian_scale     <- jacobian_scale (source/xnet_jacobian_dense.F90:136-255)
 *   xb_jacobian_decomp    <- jacobian_decomp -> LUDecompBatched_GPU/cublasDgetrfBatched
 *                            (source/xnet_jacobian_dense.F90:549-616, source/xnet_linalg.F90:1814-1870)
 *   xb_jacobian_bksub     <- jacobian_bksub -> cublasDgetrsBatched
 *                            (source/xnet_jacobian_dense.F90:618-732, source/xnet_linalg.F90:1895-1960)
 *   xb_jacobian_solve     <- jacobian_solve (source/xnet_jacobian_dense.F90:436-547)
 *   xb_step_be            <- step_be Newton loop (source/xnet_integrate_be.F90:233-482)
 *   xb_evolve             <- full_net + solve_be + timestep, device resident
 *                            (source/xnet_evolve.F90:19-269, xnet_integrate_be.F90:32-231,
 *                             xnet_integrate.F90:13-268)
 */
#ifndef XNET_B200_H
#define XNET_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XB_OK 0
#define XB_ERR_ARG (-1)
#define XB_ERR_IO (-2)
#define XB_ERR_CUDA (-3)
#define XB_ERR_UNSUPPORTED (-4)

typedef struct xb_net xb_net;   /* zone-invariant network tables (host copy)     */
typedef struct xb_ctx xb_ctx;   /* device context: tables + per-zone state in HBM */

/* Network tables as the Fortran modules hold them (nuclear_data, reaction_data, xnet_ffn).
 * c = 0..3 indexes 1..4-reactant reactions.  All index values 1-based, 0 = blank slot. */
typedef struct {
  int32_t ny, nffn, nnnu;
  int32_t nreac[4], nan[4];
  const double *rc[4];       /* rc_n(7,nreac)                                        */
  const int32_t *ni[4];      /* n1i(5,nr1), n2i..n4i(6,nr)  (leading dim given below) */
  int32_t ni_ld[4];          /* 5,6,6,6 for Fortran arrays                           */
  const int32_t *iwk[4], *irev[4];
  const int32_t *la, *le;    /* la(4,ny), le(4,ny) as in Fortran (class index fastest) */
  const int32_t *mu[4];      /* mu_n(nan)                                            */
  const double *a[4];        /* a_n(nan)                                             */
  const double *aa, *zz, *mex, *be; /* (ny) */
  double mex_n, mex_p;
  const double *g;           /* g(24,ny)                                             */
  const double *t9i;         /* t9i(24)                                              */
  const double *ffnsum, *ffnenu; /* (nffn,143) Fortran order: rate index fastest     */
  const double *ffn_ec, *ffn_beta, *ffn_qval; const int32_t *has_logft; /* may be NULL */
  /* optional reaction-pair matching of Module xnet_match (source/xnet_match.F90): mflx = 0 / NULL = no flux output.
   * ifl_n(nreac_n) signed 1-based pair index, nflx(8,mflx), iwflx(mflx), qflx(mflx) (may be NULL); the double-counting
   * factors dcflx(2,mflx) of xnet_flux may be NULL: they are then formed as flux_init does (xnet_flux.F90:36-50). */
  int32_t mflx;
  const int32_t *ifl[4];
  const int32_t *nflx, *iwflx;
  const double *qflx, *dcflx;
} xb_net_arrays;

/* Controls (xnet_controls.F90:20-62) */
typedef struct {
  int32_t iweak0, iscrn, iheat, isolv, kstmx, kitmx, ijac, iconvc;
  double changemx, changemxt, yacc, tolm, tolc, tolt9, ymin, tdel_maxmult;
  int32_t arith;             /* 0 = fast (FMA contraction, CUDA libm); 1 = strict (no FMA,
                                portable elementary functions: bit-reproducible vs the oracle) */
  int32_t solver;            /* linear solver of the fused evolve (xb_evolve*) in fast arithmetic:
                                0 = dense LU with (threshold) partial pivoting, DGETRF's algorithm;
                                1 = structured LU: species reordered light-particles-last, skyline envelope from the
                                    reaction maps (what sparse_check finds, source/xnet_preprocess.F90:291-503; the
                                    reference's answer is xnet_jacobian_MA48.F90), elimination without interchanges
                                    verified column by column (tau |a_ik| <= |a_kk|), partial pivoting inside the
                                    light-particle border, dense partial-pivoting fallback for a factorisation that
                                    fails the test.  Ignored (dense) with arith = 1 and by the module-level drop-ins. */
} xb_ctl;

const char *xb_last_error(void);
int xb_version(void);
/* number of kernels launched by this library since load / since last reset (bench's gpu_launches) */
long long xb_launch_count(int reset);

/* ---- network ---------------------------------------------------------------------------- */
int xb_net_load(const char *data_dir, int iweak0, xb_net **net);
int xb_net_from_arrays(const xb_net_arrays *arr, xb_net **net);
/* Helmholtz table for screening / self-heating (required before xb_ctx_create with iscrn>0 or iheat>0).
 * path: XNet's ASCII helm_table.dat, or the packed 13-plane float64 file written by tools/make_helm_table.py */
int xb_net_load_helm(xb_net *net, const char *path);
void xb_net_free(xb_net *net);
/* introspection (tests check these bit-exactly): what = "ny","nffn","nnnu","nreac1".."nreac4",
 * "nan1".."nan4","izmax" */
int xb_net_int(const xb_net *net, const char *what, int32_t *value);
/* copies an integer table; what = "la","le" (4*ny, class-major: la[c*ny+i]), "mu1".."mu4",
 * "n1i".."n4i" (6 per reaction), "iwk1".., "irev1".., "ires1".., "iffn", "chap_la","chap_le" (11) */
int xb_net_itable(const xb_net *net, const char *what, int32_t *out, int64_t cap, int64_t *n);
/* what = "a1".."a4","rc1".."rc4" (7 per reaction),"q1".."q4","aa","zz","mex","be","g","t9i",
 * "ffnsum","ffnenu" (143 per rate) */
int xb_net_dtable(const xb_net *net, const char *what, double *out, int64_t cap, int64_t *n);

/* ---- per-zone input fixtures (host readers; model_input_ascii.F90) ------------------------ */
/* thermo trajectory: returns nh and fills th/t9h/rhoh/yeh (cap entries each) + header scalars
 * hdr[0..2] = tstart,tstop,tdelstart(=min(0,file value)) */
int xb_read_thermo(const char *path, int32_t cap, int32_t *nh, double *th, double *t9h, double *rhoh,
                   double *yeh, double *hdr);
/* initial abundances: y[ny], scal[0..3] = yein,xext,aext,zext */
int xb_read_inab(const xb_net *net, const char *path, double *y, double *scal);

/* start-of-run conditions of a fixture (net.F90:249-268): scal = output of xb_read_inab;
 * out4 = t9start, rhostart, yestart, 0 */
int xb_zone_start(const xb_net *net, int32_t nh, const double *th, const double *t9h, const double *rhoh,
                  const double *yeh, double tstart, const double *y, const double *scal, int32_t *nstart,
                  double *out4);

/* ---- device context ------------------------------------------------------------------------ */
int xb_ctx_create(const xb_net *net, const xb_ctl *ctl, int32_t nzevolve, int32_t device, xb_ctx **ctx);
void xb_ctx_destroy(xb_ctx *ctx);

/* ---- module-level drop-ins (zone slice [zlo,zhi], 1-based inclusive like zb_lo:zb_hi) ------- */
/* cross_sect: inputs yt(ny,nz), t9t, rhot (+xext,aext,zext may be NULL = 0,1,0); outputs yet;
 * csect/h stay device resident.  mask NULL = all. */
int xb_cross_sect(xb_ctx *ctx, int32_t zlo, int32_t zhi, const int32_t *mask, const double *yt,
                  const double *t9t, const double *rhot, const double *xext, const double *aext,
                  const double *zext, double *yet);
/* copies the device-resident csect back: csect[(nr1+nr2+nr3+nr4), nz] */
int xb_get_csect(xb_ctx *ctx, int32_t zlo, int32_t zhi, double *csect);
int xb_yderiv(xb_ctx *ctx, int32_t zlo, int32_t zhi, const int32_t *mask, const double *yt, double *ydot);
/* jacobian_build (xnet_jacobian_dense.F90:257-434): dydotdy from the device-resident yt / csect (kept on the device,
 * as the reference keeps its module array), then jacobian_scale.  diag / mult NULL = diag0 (0) / mult1 (1). */
int xb_jacobian_build(xb_ctx *ctx, int32_t zlo, int32_t zhi, const int32_t *mask, const double *diag,
                      const double *mult);
/* jacobian_scale (xnet_jacobian_dense.F90:136-255): jac = mult * dydotdy^T + diag on the diagonal from the dydotdy of
 * the last xb_jacobian_build, without re-assembly (BDF refactors this way, xnet_integrate_bdf.F90:925-927). */
int xb_jacobian_scale(xb_ctx *ctx, int32_t zlo, int32_t zhi, const int32_t *mask, const double *diag,
                      const double *mult);
/* copies jac(msize,msize,nz) back (idiag>=5 style inspection) */
int xb_get_jac(xb_ctx *ctx, int32_t zlo, int32_t zhi, double *jac);
/* uploads jac(msize,msize,nz) (lets a caller use decomp/bksub as a pure batched solver) */
int xb_set_jac(xb_ctx *ctx, int32_t zlo, int32_t zhi, const double *jac);
int xb_jacobian_decomp(xb_ctx *ctx, int32_t zlo, int32_t zhi, const int32_t *mask, int32_t *info);
int xb_get_indx(xb_ctx *ctx, int32_t zlo, int32_t zhi, int32_t *indx);
int xb_jacobian_bksub(xb_ctx *ctx, int32_t zlo, int32_t zhi, const int32_t *mask, const double *yrhs,
                      const double *t9rhs, double *dy, double *dt9);
int xb_jacobian_solve(xb_ctx *ctx, int32_t zlo, int32_t zhi, const int32_t *mask, const double *yrhs,
                      const double *t9rhs, double *dy, double *dt9, int32_t *info);
/* Self-heating (iheat > 0, msize = ny+1) at module level: the T9 equation is row/column ny of jac, t9rhs/dt9 are
 * required by bksub/solve, and the module variables the reference keeps between calls are mirrored on the device:
 * dcsect*(nr,nz) (xnet_integrate.F90:807-815), cv/etae/detaedt9 (xnet_conditions, set by update_eos inside
 * cross_sect) and t9dot (set by yderiv, xnet_integrate.F90:471-487).  NULL outputs are skipped. */
int xb_get_dcsect(xb_ctx *ctx, int32_t zlo, int32_t zhi, double *dcsect);
int xb_get_eos(xb_ctx *ctx, int32_t zlo, int32_t zhi, double *cv, double *etae, double *detaedt9, double *t9dot);

/* flux (source/xnet_flux.F90:67-150): net fluxes of the mflx matched forward/reverse reaction pairs
 * (match_react, source/xnet_preprocess.F90:108-289; mflx, nflx, ifl_orig, ifl_term through xb_net_int/xb_net_itable)
 * from the device-resident rates of the last xb_cross_sect / xb_step_be call and yt (NULL = the abundances already
 * on the device).  flx(mflx,nz) host; masked-out zones keep the caller's values. */
int xb_flux(xb_ctx *ctx, int32_t zlo, int32_t zhi, const int32_t *mask, const double *yt, double *flx);
/* fused path: record flx in the per-step history (xb_out.hist_flx); call before xb_zones_upload.  xb_evolve switches
 * it on by itself when out->hist_flx is set. */
int xb_set_flux_output(xb_ctx *ctx, int32_t on);

/* step_be (source/xnet_integrate_be.F90:233-482): the complete Newton-Raphson loop of ONE backward-Euler step in one
 * launch -- no host synchronisation per iteration (the reference has one, :430).  Module-variable interface of the
 * reference spelled out as arguments, all host arrays over the slice:
 *   inr(nz)    in: 0 = active, anything else = inactive (left untouched); out: > 0 = Newton iterations to convergence,
 *              0 = not converged in kitmx iterations
 *   y(ny,nz)   abundances at the start of the step;  yt(ny,nz) in: first guess, out: last iterate
 *   tdel(nz)   step size (non-zero for active zones);  t9(nz) temperature at the start of the step
 *   t9t(nz)    in: trial temperature, out: updated when iheat > 0;  rhot(nz) trial density
 *   nh(nz)     trajectory length (rates are re-evaluated at kit = 1 only when nh > 1, :300-307); NULL = 2
 *   ydot, t9dot (optional out) derivative of the last iterate's evaluation; kcount(5,nz) (optional out) increments
 *   of ktot(3:5) = Jacobian builds, yderiv calls, cross_sect calls made by this call (ktot(1:2) belong to solve_be).
 * With iheat = 0 and nh <= 1 the rates of the previous xb_cross_sect call are used, as in the reference. */
int xb_step_be(xb_ctx *ctx, int32_t zlo, int32_t zhi, int32_t kstep, int32_t *inr, const double *y, double *yt,
               const double *tdel, const double *t9, double *t9t, const double *rhot, const int32_t *nh,
               double *ydot, double *t9dot, int32_t *kcount);

/* ---- fused evolve (full_net on device) ------------------------------------------------------ */
typedef struct {
  int32_t nfix;              /* distinct fixtures                                          */
  const int32_t *nh, *off;   /* [nfix] points / offset into th,t9h,rhoh                    */
  const double *th, *t9h, *rhoh;
  const double *tstart, *tstop, *tdelstart;  /* [nfix]                                      */
  const int32_t *nstart;                     /* [nfix]                                      */
  const double *t9start, *rhostart, *yestart;/* [nfix]                                      */
  const double *ystart;                      /* [nfix][ny]                                  */
  const double *xext, *aext, *zext;          /* [nfix]                                      */
  int32_t nzone;
  const int32_t *zfix;       /* [nzone] fixture of each zone                               */
  const double *rho_scale;   /* [nzone] or NULL                                            */
  /* optional restart state (NULL = start from fixture): see DESIGN.md "restart parity"     */
  const double *rs_t, *rs_tdel, *rs_t9, *rs_rho, *rs_t9o, *rs_y, *rs_yo; const int32_t *rs_kstep0;
} xb_zones;

typedef struct {
  double *y;                 /* [nzone][ny]                                                */
  double *t, *t9, *rho, *tdel; /* [nzone]                                                  */
  int32_t *kstep, *ktot, *status; /* [nzone], [nzone][5], [nzone]                          */
  int32_t hist_cap;          /* 0 = no per-step history                                    */
  int32_t *hist_n, *hist_kstep, *hist_kmon;
  double *hist_t, *hist_t9, *hist_rho, *hist_tdel, *hist_edot, *hist_sqnu, *hist_y;
  double *hist_flx;          /* [nzone][hist_cap][mflx] matched-pair fluxes of each record (flx of ts_output,
                                source/xnet_output.F90:78-79); NULL = not wanted                          */
} xb_out;

/* host buffers in, host buffers out (H2D + kernel + D2H): the end-to-end call */
int xb_evolve(xb_ctx *ctx, const xb_zones *zones, xb_out *out);
/* split form for device-resident timing: upload once, run many times, download */
int xb_zones_upload(xb_ctx *ctx, const xb_zones *zones, int32_t hist_cap);
int xb_evolve_resident(xb_ctx *ctx, float *kernel_ms);   /* runs on the uploaded zones */
int xb_results_download(xb_ctx *ctx, xb_out *out);
/* ---- ts_<zone> output (source/net.F90:281-301 header, source/xnet_output.F90:78-79 records) ---------------
 * Fortran sequential-unformatted file exactly as the reference writes it (gfortran record markers), so the
 * reference's tools (tools/python/read_ts_file.py, test/compare_ts.py) read it.  Strings are blank-padded to 80. */
typedef struct {
  const char *descript[3];   /* run description lines of the control file              */
  const char *data_desc;     /* description line of the network data (net_desc)        */
  const xb_ctl *ctl;         /* kstmx,kitmx,iweak0,iscrn,iconvc,changemx,tolm,tolc,yacc,ymin,tdel_maxmult,iheat,isolv */
  const char *inab_file, *abund_desc, *thermo_file, *thermo_desc;
  int32_t ny; const double *zz, *aa;
  int32_t mflx; const int32_t *ifl_orig, *ifl_term;
} xb_ts_header;
/* nrec records: kstep, t, t9, rho, sqnu, tdel, edot, y(ny) [nrec][ny], flx(mflx) [nrec][mflx] (NULL = zeros) */
int xb_ts_write(const char *path, const xb_ts_header *hdr, int32_t nrec, const int32_t *kstep, const double *t,
                const double *t9, const double *rho, const double *sqnu, const double *tdel, const double *edot,
                const double *y, const double *flx);

/* ASCII evolution file ev_<zone> (itsout >= 2): header of source/net.F90:311-314, one line per output step in the
 * reference's format (i4,1es15.8,2es10.3,3es10.2,<nout>es9.2,2i2) (source/xnet_output.F90:72,83-85): kstep, t, t9, rho,
 * edot, sqnu, tdel, the mass fractions xout[nrec][nout] = aa*y of the nout selected species (names5: nout x 5
 * characters, as nname), kmon[nrec][2]. */
int xb_ev_write(const char *path, int32_t nrec, const int32_t *kstep, const double *t, const double *t9,
                const double *rho, const double *edot, const double *sqnu, const double *tdel, int32_t nout,
                const char *names5, const double *xout, const int32_t *kmon);

/* per-phase cycle counters of the last resident run (sum over CTAs), 24 values:
 * rates,yderiv,jac,lu,bksub,other,total,zones, lu_panel,lu_swap_trsm,lu_gemm,pivoting-panel fallbacks,
 * then 12 fine-grained development timers */
int xb_phase_cycles(xb_ctx *ctx, double *cycles24);
/* bytes moved host->device / device->host by the last xb_evolve call */
int xb_last_copy_bytes(xb_ctx *ctx, int64_t *h2d, int64_t *d2h);
/* duration (ms, CUDA events on the launching stream) of the last kernel this context launched: a module-level
 * drop-in (xb_cross_sect ... xb_step_be) or the evolve kernel */
int xb_last_kernel_ms(xb_ctx *ctx, float *ms);
/* structured solver of this context: info8 = active (0/1), msize, doubles of the skyline store, flops of one
 * factorisation (2 per multiply-add), zones resident per SM, rows/columns of the pivoted border, warps that
 * factor, reaction rates cached in shared memory */
int xb_solver_info(xb_ctx *ctx, double *info8);
/* structured-LU index tables of a network (heat = 1: with the T9 equation); what = "perm","pos","rf","cf","cb","cbot",
 * "nrb","ncb","sizes" (n, nstore, nbord, flops, bo, bs, flops of the stored envelope) -- see SluTables in
 * xnet_b200/csrc/xb_host.h */
int xb_slu_table(const xb_net *net, int32_t heat, const char *what, int32_t *out, int64_t cap, int64_t *n);
/* nz dense systems a[nz][n*n] (column-major, equation order, zero outside the structural envelope) solved with the
 * structured LU of a solver = 1 context: mode 0 = right-hand side carried through the factorisation, 1 = factor, then
 * forward + backward solve, 2 / 3 = the dense fallback likewise.  flags[nz]: bit 0 = threshold test failed, bit 1 = a
 * non-zero outside the envelope was dropped, bit 2 = singular (dense fallback). */
int xb_slu_solve(xb_ctx *ctx, int32_t nz, const double *a, const double *b, double *x, int32_t *flags, int32_t mode);
/* cfg6 = grid, threads per CTA, dynamic smem bytes, jacobian-in-smem flag, SM count, msize */
int xb_launch_config(xb_ctx *ctx, int32_t *cfg6);
/* measured FP64 FMA peak of this device (microbenchmark, FLOP/s) */
int xb_measure_fp64_peak(int32_t device, double *flops);

#ifdef __cplusplus
}
#endif
#endif
