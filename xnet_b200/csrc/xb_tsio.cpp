// Writer of XNet's binary time-series file `ts_<zone>` (Fortran sequential-unformatted: every record is
// int32 length | payload | int32 length, little-endian as written by gfortran).  Record order and payloads follow
// the Write statements of the reference: header net.F90:288-301, one record per output step xnet_output.F90:78-79.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/xnet_b200.h"

namespace {

struct Record {
  std::vector<unsigned char> b;
  void put(const void *p, size_t n) { const unsigned char *c = static_cast<const unsigned char *>(p); b.insert(b.end(), c, c + n); }
  void i4(int32_t v) { put(&v, 4); }
  void f8(double v) { put(&v, 8); }
  // Character(80): left-justified, blank padded, truncated at 80
  void c80(const char *s) {
    char buf[80];
    std::memset(buf, ' ', 80);
    if (s) { size_t n = std::strlen(s); if (n > 80) n = 80; std::memcpy(buf, s, n); }
    put(buf, 80);
  }
  bool flush(FILE *f) {
    const int32_t len = (int32_t)b.size();
    const bool ok = std::fwrite(&len, 4, 1, f) == 1 && (b.empty() || std::fwrite(b.data(), 1, b.size(), f) == b.size()) &&
                    std::fwrite(&len, 4, 1, f) == 1;
    b.clear();
    return ok;
  }
};

}  // namespace

extern "C" int xb_ts_write(const char *path, const xb_ts_header *h, int32_t nrec, const int32_t *kstep, const double *t,
                           const double *t9, const double *rho, const double *sqnu, const double *tdel, const double *edot,
                           const double *y, const double *flx) {
  if (!path || !h || !h->ctl || nrec < 0 || (nrec > 0 && (!kstep || !t || !t9 || !rho || !sqnu || !tdel || !edot || !y)))
    return XB_ERR_ARG;
  if (h->ny < 1 || !h->zz || !h->aa || h->mflx < 0 || (h->mflx > 0 && (!h->ifl_orig || !h->ifl_term))) return XB_ERR_ARG;
  FILE *f = std::fopen(path, "wb");
  if (!f) return XB_ERR_IO;
  Record r;
  bool ok = true;
  // 1: descript(1:3), data_desc
  for (int i = 0; i < 3; ++i) r.c80(h->descript[i]);
  r.c80(h->data_desc);
  ok = ok && r.flush(f);
  // 2: kstmx,kitmx,iweak0,iscrn,iconvc,changemx,tolm,tolc,yacc,ymin,tdel_maxmult,iheat,isolv
  const xb_ctl *c = h->ctl;
  r.i4(c->kstmx); r.i4(c->kitmx); r.i4(c->iweak0); r.i4(c->iscrn); r.i4(c->iconvc);
  r.f8(c->changemx); r.f8(c->tolm); r.f8(c->tolc); r.f8(c->yacc); r.f8(c->ymin); r.f8(c->tdel_maxmult);
  r.i4(c->iheat); r.i4(c->isolv);
  ok = ok && r.flush(f);
  // 3, 4: input file names and their description lines
  r.c80(h->inab_file); r.c80(h->abund_desc); ok = ok && r.flush(f);
  r.c80(h->thermo_file); r.c80(h->thermo_desc); ok = ok && r.flush(f);
  // 5: ny, zz, aa
  r.i4(h->ny); r.put(h->zz, (size_t)h->ny * 8); r.put(h->aa, (size_t)h->ny * 8); ok = ok && r.flush(f);
  // 6: mflx, ifl_orig, ifl_term
  r.i4(h->mflx);
  if (h->mflx > 0) { r.put(h->ifl_orig, (size_t)h->mflx * 4); r.put(h->ifl_term, (size_t)h->mflx * 4); }
  ok = ok && r.flush(f);
  // 7...: kstep, t, t9, rho, sqnu, tdel, edot, y(1:ny), flx(1:mflx)
  const std::vector<double> zero((size_t)h->mflx, 0.0);
  for (int32_t n = 0; ok && n < nrec; ++n) {
    r.i4(kstep[n]); r.f8(t[n]); r.f8(t9[n]); r.f8(rho[n]); r.f8(sqnu[n]); r.f8(tdel[n]); r.f8(edot[n]);
    r.put(y + (size_t)n * h->ny, (size_t)h->ny * 8);
    if (h->mflx > 0) r.put(flx ? flx + (size_t)n * h->mflx : zero.data(), (size_t)h->mflx * 8);
    ok = r.flush(f);
  }
  ok = (std::fclose(f) == 0) && ok;
  return ok ? XB_OK : XB_ERR_IO;
}
