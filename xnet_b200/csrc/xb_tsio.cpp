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

// ---- ASCII evolution file `ev_<zone>` (itsout >= 2) --------------------------------------------------------------
// Header: net.F90:311-314, format (a4,a15,5a10,<n>a9,a4); one line per output step: xnet_output.F90:72 and :83-85,
// format (i4,1es15.8,2es10.3,3es10.2,<n>es9.2,2i2) = kstep, t, t9, rho, edot, sqnu, tdel, aa*y of the nnucout selected
// species, kmon(1:2).
namespace {

// Fortran ESw.d edit descriptor: one digit before the point, exponent E+ee, or +eee without the letter when it needs
// three digits; right-justified in w columns, asterisks when it does not fit.
std::string fortran_es(double v, int w, int d) {
  char buf[64];
  std::string s;
  if (v != v) s = "NaN";
  else if (v > 1.7976931348623157e308) s = "Infinity";
  else if (v < -1.7976931348623157e308) s = "-Infinity";
  else {
    std::snprintf(buf, sizeof buf, "%.*E", d, v);
    s = buf;
    const size_t e = s.find('E');
    if (e != std::string::npos && s.size() - e - 2 > 2) s.erase(e, 1);   // three exponent digits: the E is dropped
  }
  if ((int)s.size() > w) return std::string((size_t)w, '*');
  return std::string((size_t)w - s.size(), ' ') + s;
}
std::string fortran_i(long v, int w) {
  std::string s = std::to_string(v);
  if ((int)s.size() > w) return std::string((size_t)w, '*');
  return std::string((size_t)w - s.size(), ' ') + s;
}
// Fortran Aw output of a string of length n: right-justified when shorter than w, the leftmost w characters otherwise
std::string fortran_a(const std::string &s, int w) {
  if ((int)s.size() >= w) return s.substr(0, (size_t)w);
  return std::string((size_t)w - s.size(), ' ') + s;
}

}  // namespace

extern "C" int xb_ev_write(const char *path, int32_t nrec, const int32_t *kstep, const double *t, const double *t9,
                           const double *rho, const double *edot, const double *sqnu, const double *tdel, int32_t nout,
                           const char *names5, const double *xout, const int32_t *kmon) {
  if (!path || nrec < 0 || nout < 0 || (nout > 0 && (!names5 || !xout)) ||
      (nrec > 0 && (!kstep || !t || !t9 || !rho || !edot || !sqnu || !tdel || !kmon)))
    return XB_ERR_ARG;
  FILE *f = std::fopen(path, "w");
  if (!f) return XB_ERR_IO;
  std::string line = fortran_a("k ", 4) + fortran_a(" Time ", 15) + fortran_a(" T(GK) ", 10) + fortran_a(" Density ", 10) +
                     fortran_a(" dE/dt ", 10) + fortran_a(" sqnu ", 10) + fortran_a(" Timestep ", 10);
  for (int i = 0; i < nout; ++i) line += fortran_a(std::string(names5 + (size_t)5 * i, 5), 9);
  line += fortran_a(" It ", 4);
  bool ok = std::fprintf(f, "%s\n", line.c_str()) > 0;
  for (int32_t n = 0; ok && n < nrec; ++n) {
    line = fortran_i(kstep[n], 4) + fortran_es(t[n], 15, 8) + fortran_es(t9[n], 10, 3) + fortran_es(rho[n], 10, 3) +
           fortran_es(edot[n], 10, 2) + fortran_es(sqnu[n], 10, 2) + fortran_es(tdel[n], 10, 2);
    for (int i = 0; i < nout; ++i) line += fortran_es(xout[(size_t)n * nout + i], 9, 2);
    line += fortran_i(kmon[2 * n], 2) + fortran_i(kmon[2 * n + 1], 2);
    ok = std::fprintf(f, "%s\n", line.c_str()) > 0;
  }
  ok = (std::fclose(f) == 0) && ok;
  return ok ? XB_OK : XB_ERR_IO;
}
