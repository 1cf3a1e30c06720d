/* Oracle (TEST INFRASTRUCTURE ONLY): EOS interface, screening and logft phase space.
 * Product code must never include or link this file.
 *
 * ox_eos_interface: xnet_eos_starkiller.F90:107-204 (Ye from y_moment always; cv/eta only when
 *                   iscrn>0 or iheat>0 -- xnet_eos_starkiller.F90:89)
 * y_moment:         xnet_abundances.F90:137-166
 */
#include "xnet_oracle.h"
#include <math.h>

static const double thbim1 = 1.58;

void ox_y_moment(const ox_net *net, const double *y, double xext, double aext, double zext,
                 double *ye, double *ytot, double *abar, double *zbar, double *z2bar, double *zibar) {
  double yext = xext / aext, ntot = 0.0, atot = 0.0, ztot = 0.0, z2tot = 0.0, zitot = 0.0;
  int k;
  for (k = 0; k < net->ny; k++) {
    double zz = net->zz[k];
    ntot = ntot + y[k];
    atot = atot + y[k] * net->aa[k];
    ztot = ztot + y[k] * zz;
    z2tot = z2tot + y[k] * (zz * zz);
    zitot = zitot + y[k] * pow(zz, thbim1);
  }
  ntot = ntot + yext;
  atot = atot + xext;
  ztot = ztot + yext * zext;
  z2tot = z2tot + yext * zext * zext;
  zitot = zitot + yext * pow(zext, thbim1);
  *ye = ztot / atot; *ytot = ntot; *abar = atot / ntot; *zbar = ztot / ntot;
  *z2bar = z2tot / ntot; *zibar = zitot / ntot;
}

int ox_eos_interface(const ox_net *net, const ox_ctl *ctl, double t9, double rho, const double *y,
                     double xext, double aext, double zext, double *ye, double *cv, double *etae,
                     double *detaedt9) {
  double ytot, abar, zbar, z2bar, zibar;
  (void)t9; (void)rho;
  ox_y_moment(net, y, xext, aext, zext, ye, &ytot, &abar, &zbar, &z2bar, &zibar);
  *cv = 0.0; *etae = 0.0; *detaedt9 = 0.0;
  if (ctl->iscrn > 0 || ctl->iheat > 0) return -100; /* Helmholtz EOS not restated yet */
  return 0;
}

int ox_screening(const ox_net *net, const ox_ctl *ctl, double t9, double rho, const double *y,
                 double xext, double aext, double zext, double etae, double detaedt9,
                 double *h[4], double *dh[4]) {
  (void)net; (void)ctl; (void)t9; (void)rho; (void)y; (void)xext; (void)aext; (void)zext;
  (void)etae; (void)detaedt9; (void)h; (void)dh;
  return -101; /* screening not restated yet */
}

void ox_effphase(double t9, double chem, double qec, double *phasei, double *dphaseidt9) {
  (void)t9; (void)chem; (void)qec; *phasei = 0.0; *dphaseidt9 = 0.0; /* logft not restated yet */
}
