// Host-side state shared by the translation units of libxnet_b200.so (xb_api.cu: dense kernels and the C ABI;
// xb_slu.cu: the structured-LU evolve kernel).  Include after xb_zone.cuh.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/xnet_b200.h"
#include "xb_host.h"

using namespace xb;

// ------------------------------------------------------------------------------------------------
// error text / launch counter of the library (defined in xb_api.cu)
int xb_fail(int code, const std::string &msg);
void xb_count_launch();
static inline int fail(int code, const std::string &msg) { return xb_fail(code, msg); }
#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return fail(XB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));             \
  } while (0)

struct xb_net { Network h; };

template <class T>
struct DevBuf {
  T *p = nullptr; size_t n = 0;
  cudaError_t put(const T *src, size_t count) {
    cudaError_t e = resize(count);
    if (e != cudaSuccess) return e;
    if (count == 0) return cudaSuccess;
    return cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice);
  }
  cudaError_t put(const std::vector<T> &v) { return put(v.data(), v.size()); }
  cudaError_t resize(size_t count) {
    if (count <= n && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) n = std::max<size_t>(count, 1);
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct xb_ctx {
  const xb_net *net = nullptr;
  xb_ctl ctl{};
  int device = 0, nzevolve = 0, nsm = 0;
  DevNet dn{};
  DevCtl dc{};
  // network tables
  DevBuf<double> rcT, a[4], aa, zz, mex, be, g, lng, dlng, t9i, ffnsum, ffnenu;
  DevBuf<int> ni, iwk, irev, iffn, la, le, mu[4], nn[4][4], ffr_n1, ffr_iffn;
  DevBuf<double> helm, helm_t, helm_d, zzi, zseq53, sc_zw, sc_zi, z_zexti, ffn_ec, ffn_beta, ffn_qval;
  DevBuf<int> has_logft;
  DevBuf<int2> sc_iz;
  // term plans (ydot, Jacobian)
  struct PlanBufs { DevBuf<double> a; DevBuf<uint2> nn, mc; DevBuf<int> seg_start, seg_dest, multi_start, multi_dest; } pby, pbj;
  int ld = 0, npart = 0;
  // launch configuration
  int threads = 256, grid = 0, jac_in_smem = 1;
  size_t smem_bytes = 0;
  DevScratch scr{};
  DevBuf<double> scratch;
  DevBuf<unsigned long long> phase;
  DevBuf<int> counter;
  // module-level per-zone state (HBM)
  DevBuf<double> m_yt, m_t9t, m_rhot, m_ext, m_yet, m_csect, m_qk, m_ydot, m_jac, m_rhs, m_diag, m_mult;
  DevBuf<double> m_dyd;                        // dydotdy^T of the last xb_jacobian_build (the reference's module array; allocated on first use)
  DevBuf<double> m_dcsect, m_eos, m_t9dot, m_y, m_t9, m_tdel;
  DevBuf<int> m_indx, m_info, m_mask, m_nh, m_inr, m_kcnt;
  // evolve: zones + outputs
  DevZones dz{};
  DevOut dout{};
  int nzone = 0, hist_cap = 0;
  DevBuf<int> z_nh, z_off, z_nstart, z_zfix, z_kstep0;
  DevBuf<double> z_th, z_t9h, z_rhoh, z_tstart, z_tstop, z_tdelstart, z_t9start, z_rhostart, z_yestart, z_ystart,
      z_xext, z_aext, z_zext, z_rho_scale, z_rs_t, z_rs_tdel, z_rs_t9, z_rs_rho, z_rs_t9o, z_rs_y, z_rs_yo;
  DevBuf<double> o_y, o_t, o_t9, o_rho, o_tdel, o_ht, o_ht9, o_hrho, o_htdel, o_hedot, o_hsqnu, o_hy, o_hflx;
  DevBuf<int> fl_start, fl_k;
  DevBuf<double> fl_dc, m_flx;
  int want_flux = 0;
  DevBuf<int> o_kstep, o_ktot, o_status, o_hn, o_hkstep, o_hkmon;
  int64_t h2d = 0, d2h = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // structured-LU variant of the fused evolve (xb_slu.cu): its own DevNet (plans chunked for its thread count, skyline
  // destinations, shared-memory layout), tables, launch configuration and per-CTA scratch
  int slu_on = 0, slu_threads = 0, slu_grid = 0, slu_per_sm = 0, slu_npart = 0;
  size_t slu_smem = 0;
  DevNet dn_slu{};
  SluTables slu_tab;
  PlanBufs pby_slu, pbj_slu;
  DevBuf<uint2> slu_tabd;
  DevBuf<int> slu_pos;
  DevBuf<uint32_t> slu_schw, slu_schl, slu_soll;
  DevBuf<int> slu_solo;        // pivot schedule of the structured LU (SluSchedule, xb_host.h)
  DevScratch scr_slu{};
  DevBuf<double> scratch_slu;
};

// ---- term plans ------------------------------------------------------------------------------------
struct PlanRec { double a; int mu; int n[4]; int fseg; int fdest; int sseg; };
struct PlanHost {
  int nrec = 0, L = 0, nseg = 0, nmulti = 0, npart = 0;
  std::vector<double> a; std::vector<uint2> nn, mc;
  std::vector<int> seg_start, seg_dest, multi_start, multi_dest;
};

// recs sorted by segment (fseg non-decreasing, sseg non-decreasing); fdest = (destination<<1)|diag of
// the record's fast segment; strict segments are [sseg_start[i], sseg_start[i+1]) with sseg_dest[i].
static inline PlanHost make_plan(const std::vector<PlanRec> &recs, int NT, int dummy, std::vector<int> sseg_start,
                   std::vector<int> sseg_dest) {
  PlanHost P;
  P.nrec = (int)recs.size();
  P.L = std::max(1, (P.nrec + NT - 1) / NT);
  P.L = (P.L + kPlanU - 1) / kPlanU * kPlanU;
  const size_t tot = (size_t)P.L * NT;
  const unsigned dd = (unsigned)dummy | ((unsigned)dummy << 16);
  P.a.assign(tot, 0.0); P.nn.assign(tot, make_uint2(dd, dd)); P.mc.assign(tot, make_uint2(0u, 0u));
  P.nseg = (int)sseg_dest.size();
  P.seg_start = std::move(sseg_start); P.seg_dest = std::move(sseg_dest);
  // extent of every fast segment
  std::vector<int> first, last;   // by fseg
  for (int r = 0; r < P.nrec; ++r) {
    const int g = recs[r].fseg;
    if (g >= (int)first.size()) { first.resize(g + 1, -1); last.resize(g + 1, -1); }
    if (first[g] < 0) first[g] = r;
    last[g] = r;
  }
  int open_multi = -1;
  for (int t = 0; t < NT; ++t) {
    const int b = t * P.L, e = std::min(P.nrec, (t + 1) * P.L);
    for (int r = b; r < e; ++r) {
      const PlanRec &R = recs[r];
      const size_t at = (size_t)(r - b) * NT + t;
      P.a[at] = R.a;
      P.nn[at] = make_uint2((unsigned)R.n[0] | ((unsigned)R.n[1] << 16), (unsigned)R.n[2] | ((unsigned)R.n[3] << 16));
      unsigned ctl = 0;
      const bool piece_end = (r + 1 == e) || recs[r + 1].fseg != R.fseg;
      if (piece_end) {
        int pb = r;
        while (pb > b && recs[pb - 1].fseg == R.fseg) --pb;
        const bool complete = (pb == first[R.fseg]) && (r == last[R.fseg]);
        if (complete) ctl = 1u | ((unsigned)(R.fdest & 1) << 2) | ((unsigned)(R.fdest >> 1) << 3);
        else {
          if (open_multi != R.fseg) { P.multi_start.push_back(P.npart); P.multi_dest.push_back(R.fdest); open_multi = R.fseg; }
          ctl = 2u | ((unsigned)P.npart << 3);
          ++P.npart;
        }
      }
      P.mc[at] = make_uint2((unsigned)R.mu, ctl);
    }
  }
  P.nmulti = (int)P.multi_dest.size();
  P.multi_start.push_back(P.npart);
  return P;
}

static inline int upload_plan(xb_ctx::PlanBufs &B, DevPlan &D, const PlanHost &P) {
  CUDA_TRY(B.a.put(P.a)); CUDA_TRY(B.nn.put(P.nn)); CUDA_TRY(B.mc.put(P.mc));
  std::vector<int> ss(P.seg_start), sd(P.seg_dest), ms(P.multi_start), md(P.multi_dest);
  if (sd.empty()) sd.push_back(0);
  if (md.empty()) md.push_back(0);
  CUDA_TRY(B.seg_start.put(ss)); CUDA_TRY(B.seg_dest.put(sd)); CUDA_TRY(B.multi_start.put(ms)); CUDA_TRY(B.multi_dest.put(md));
  D.nrec = P.nrec; D.L = P.L; D.nseg = P.nseg; D.nmulti = P.nmulti; D.npart = P.npart;
  D.a = B.a.p; D.nn = B.nn.p; D.mc = B.mc.p; D.seg_start = B.seg_start.p; D.seg_dest = B.seg_dest.p;
  D.multi_start = B.multi_start.p; D.multi_dest = B.multi_dest.p;
  return XB_OK;
}

// ydot plan: species-major, reaction class, ascending la..le (xnet_integrate.F90:409-469).
// Jacobian plan: contributions generated in the order of the reference scatter loop
// (xnet_jacobian_dense.F90:303-341), then grouped by element (stable): element (i0, col) of jac.
// slu == nullptr: destinations (column << 10 | row) of the dense tile; else the linear address of the element in the skyline
// store of the structured LU (equations in the light-particles-last order).
static inline int build_plans(xb_ctx *c, DevNet &dn_out, xb_ctx::PlanBufs &pby, xb_ctx::PlanBufs &pbj, int NT, int ld,
                              int *npart_out, const SluTables *slu) {
  const Network &h = c->net->h;
  const DevNet &d = dn_out;
  const int ny = h.ny, N = d.msize;
  std::vector<PlanRec> ry, rj;
  std::vector<int> ys_start, ys_dest;
  for (int i = 0; i < ny; ++i)
    for (int cI = 0; cI < 4; ++cI) {
      ys_start.push_back((int)ry.size()); ys_dest.push_back(4 * i + cI);
      for (int j = h.la[(size_t)cI * ny + i] - 1; j < h.le[(size_t)cI * ny + i]; ++j) {
        const int k = h.mu[cI][j] - 1;
        PlanRec R{}; R.a = h.a[cI][j]; R.mu = d.roff[cI] + k;
        for (int r = 0; r < 4; ++r) R.n[r] = (r <= cI) ? h.ni[cI][(size_t)k * 6 + r] - 1 : ny;
        R.fseg = i; R.fdest = i << 1; R.sseg = 4 * i + cI;
        ry.push_back(R);
        // Jacobian contributions of this term, slot by slot
        for (int sl = 0; sl <= cI; ++sl) {
          PlanRec Q{}; Q.a = R.a; Q.mu = R.mu;
          for (int r = 0; r < 4; ++r) Q.n[r] = ny;
          for (int r = 0; r < cI; ++r) Q.n[r] = R.n[(sl + 1 + r) % (cI + 1)];
          const int col = R.n[sl];
          Q.fseg = i * N + col;   // element key, replaced by a dense id below
          rj.push_back(Q);
        }
      }
    }
  ys_start.push_back((int)ry.size());
  for (int i = 0; i < ny; ++i) {   // every diagonal element exists (jac = mult*d + diag)
    bool have = false;
    for (const PlanRec &Q : rj) if (Q.fseg == i * N + i) { have = true; break; }
    if (!have) { PlanRec Q{}; Q.a = 0.0; Q.mu = 0; for (int r = 0; r < 4; ++r) Q.n[r] = ny; Q.fseg = i * N + i; rj.push_back(Q); }
  }
  std::stable_sort(rj.begin(), rj.end(), [](const PlanRec &x, const PlanRec &y) { return x.fseg < y.fseg; });
  std::vector<int> js_start, js_dest;
  int id = -1, prev = -1;
  for (size_t r = 0; r < rj.size(); ++r) {
    const int key = rj[r].fseg;
    if (key != prev) {
      ++id; prev = key;
      const int i0 = key / N, col = key % N;
      int dest = (col << 10) | i0;   // device: column*ld + slot(row)
      if (slu) dest = slu->addr(slu->pos[i0], slu->pos[col]);
      js_start.push_back((int)r); js_dest.push_back((dest << 1) | (i0 == col ? 1 : 0));
    }
    rj[r].fseg = id; rj[r].sseg = id; rj[r].fdest = js_dest.back();
  }
  js_start.push_back((int)rj.size());
  PlanHost py = make_plan(ry, NT, ny, ys_start, ys_dest), pj = make_plan(rj, NT, ny, js_start, js_dest);
  int rc = upload_plan(pby, dn_out.py, py); if (rc) return rc;
  rc = upload_plan(pbj, dn_out.pj, pj); if (rc) return rc;
  dn_out.ld = ld;
  *npart_out = std::max(std::max(py.npart, pj.npart), std::max(2 * std::max(ny, dn_out.nffr) + 2, 128));
  return XB_OK;
}


// xb_slu.cu
int xb_slu_configure(xb_ctx *c);
int xb_slu_evolve(xb_ctx *c, float *kernel_ms);
void xb_slu_release(xb_ctx *c);
int xb_slu_solve_dense_test(xb_ctx *c, int nz, const double *a, const double *b, double *x, int32_t *flags, int mode);
