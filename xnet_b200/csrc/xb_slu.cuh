// Structured LU of the backward-Euler matrix: the B200 answer to the reference's sparse plug-ins
// (XNet tree: source/xnet_jacobian_MA48.F90, pattern from sparse_check, source/xnet_preprocess.F90:291-503).
//
// Equations are reordered with the light particles (n, p, d, t, he3, he4) and, with self-heating, T9 LAST.  In that
// order the matrix is a variable band plus a dense border, and an elimination without row interchanges never leaves
// its envelope (host tables: structured_lu_tables, xb_loader.cpp; layout: SluTables, xb_host.h).  Storage in shared
// memory, S[nstore], everything COLUMN by column:
//     heavy rows i < nh  : element (i,j) at cb[j] + i            (column j holds rows cf[j]..cbot[j])
//     border rows i >= nh: element (i,j) at bo + j * bs + (i-nh)  (bs odd >= nb; column n = the right-hand side)
// The multipliers are kept UNSCALED (l(i,k) = S(i,k) * rinv[k], rinv = reciprocals of U's diagonal): the pivot row is
// scaled instead, a(i,j) -= a(i,k) * (u(k,j) * rinv[k]), which takes the reciprocal off every row's dependency chain.
// SN160: 6.0k doubles (47 KB) instead of 25.6k, 1.1e5 flops instead of 2.7e6 per factorisation -- three zones fit on
// one SM where the dense tile allows one.
//
// Work distribution of one pivot k: LANES = COLUMNS (the band columns k+1..k+ncb, then the border columns, then the
// right-hand side), each lane walks down its column over the band rows k+1..k+nrb and the border rows.  A lane's
// addresses are base + immediate (no per-element index arithmetic), the multipliers a(i,k) are broadcast loads shared
// by the warp, and the column bases are padded (cb[j] mod 16 distinct over 8 neighbours; odd stride in the border) so
// that one row of consecutive columns is nearly bank-conflict free.  Rows are split in chunks of 8 over the nwl warps
// that factor.  With three zones of four warps on an SM every instruction of a pivot's dependent chain costs its full
// latency, so the pivot loop is written for instruction count: all index arithmetic is done once per network on the
// host (the pivot schedule, structured_lu_schedule in xb_loader.cpp), 93 instructions per warp and round remain.
//
// Numerical safety: every heavy column is verified, tau*|a(i,k)| <= |a(k,k)| for all rows below (the test the dense
// fast path applies to its predicted pivot; tau = 1e-3, see kSluTauDefault below).  The border block at the end keeps
// its diagonal pivot under the same test and interchanges rows (inside the border: its rows are full) only when the
// diagonal fails it.  Elimination WITHOUT interchanges is what keeps the trace abundances accurate component by
// component (an interchange between the neutron and proton rows costs c11 at Y = 3e-19 two digits on XRB304, where
// the interchange-free elimination is good to 1e-12: tools/structured_lu_accuracy.py, profiles/r02_structured_lu_accuracy.txt).
// A factorisation with a violated heavy column is discarded and redone by a dense partial-pivoting LU in the CTA's
// global scratch (slu_dense_*), so the result is always that of a bounded-growth LU.
#pragma once
#include "xb_device.cuh"
#include "xb_lu.cuh"

namespace xb {

// Default threshold of the column test (DevSlu.tau; XB_SLU_TAU in the environment overrides it at context creation).
// 1e-3 bounds the multipliers by 1000.  Measured on every step of the SN160 flame trajectory: 2 % of the
// factorisations have a multiplier above 10 (always the proton row against ga64 / ge63, whose diagonal has cancelled
// by the time it is the pivot), the largest is 222, and the solutions of those systems are as close to an
// extended-precision solve as LAPACK's partial-pivoting ones: the growth that a large multiplier could cause does not occur.
constexpr double kSluTauDefault = 1.0e-3;
constexpr int kSluBord = 8;     // border rows/columns handled in registers (light particles + T9 <= 7)

// per-zone views of the index tables (shared memory): one 64-bit entry per column j = 0..n (n = the right-hand side as a
// pseudo-column: cb = rhs_off, cf = 0):  x = cb | cf << 16,  y = ncb | nrb << 8 (spans of pivot j, heavy j only)
struct SluIdx {
  const uint2 *col;
  int n, nh, nb, bo, bs;
};
__device__ __forceinline__ SluIdx slu_idx() {
  const DevSlu &D = c_net.slu;
  SluIdx I;
  I.n = D.n; I.nb = D.nbord; I.nh = D.n - D.nbord; I.bo = D.bo; I.bs = D.bs;
  I.col = reinterpret_cast<const uint2 *>(xb_smem + D.off_tab);
  return I;
}
// copies the tables of the context into the CTA's shared memory (kernel prologue)
__device__ __forceinline__ void slu_stage_tables() {
  const DevSlu &D = c_net.slu;
  uint2 *t = reinterpret_cast<uint2 *>(xb_smem + D.off_tab);
  for (int k = threadIdx.x; k <= D.n; k += blockDim.x) t[k] = D.tab[k];
}
__device__ __forceinline__ int slu_cb(const SluIdx &I, const int j) { return (int)(I.col[j].x & 0xffffu); }
__device__ __forceinline__ int slu_cf(const SluIdx &I, const int j) { return (int)(I.col[j].x >> 16); }
__device__ __forceinline__ int slu_nrb(const SluIdx &I, const int k) { return (int)((I.col[k].y >> 8) & 0xffu); }
// last stored heavy row of column j
__device__ __forceinline__ int slu_cbot(const SluIdx &I, const int j) { return (j < I.nh) ? j + slu_nrb(I, j) : I.nh - 1; }
__device__ __forceinline__ int slu_addr(const SluIdx &I, const int i, const int j) {
  return (i >= I.nh) ? I.bo + j * I.bs + (i - I.nh) : slu_cb(I, j) + i;
}

// One chunk of a pivot's update, eight rows of this lane's column:  a[t] -= l[t] * u,  stored for t < nleft, as ONE block
// of PTX: the 16 loads, the multiply-adds, the predicated stores.  What matters is that all sixteen loads are in flight
// before the first multiply-add issues (an SM issues in order: a multiply-add placed between the loads waits for its own
// two operands while the loads behind it are not even issued -- eight exposed shared-memory latencies instead of one).
// ptxas schedules for its own latency model and interleaves them, ordering points it can prove redundant are dropped,
// so the order is enforced through a DATA dependency: the multiplier every multiply-add uses is un with (the OR of the
// low words of all sixteen loaded values) & zero folded into its low word, where zero is a run-time 0 that ptxas cannot
// see through -- nine LOP3, bit-identical arithmetic.
// Every load is unconditional (a conditionally written register would have to be kept alive around the pivot loop:
// ptxas spills it): rows past the span read what follows in the store, idle lanes are handed the multipliers' address
// (a broadcast the warp performs anyway); nothing of that is stored.
// lp / ap: shared addresses of the multipliers (the same for every lane) and of the column segment; un = -u.
// Returns a[0] (the next pivot, for the lane that owns it).
__device__ __forceinline__ double slu_chunk8(const unsigned lp, const unsigned ap, const int nleft, const double un, const unsigned zero) {
  double a0;
  asm volatile(
      "{\n\t"
      ".reg .pred p0, p1, p2, p3, p4, p5, p6, p7;\n\t"
      ".reg .f64 l0, l1, l2, l3, l4, l5, l6, l7, a0, a1, a2, a3, a4, a5, a6, a7, ug;\n\t"
      ".reg .b32 x0, x1, x2, x3, x4, x5, x6, x7, y0, y1, y2, y3, y4, y5, y6, y7, hi, acc, ulo, uhi;\n\t"
      "ld.shared.f64 l0, [%1];\n\tld.shared.f64 l1, [%1+8];\n\tld.shared.f64 l2, [%1+16];\n\tld.shared.f64 l3, [%1+24];\n\t"
      "ld.shared.f64 l4, [%1+32];\n\tld.shared.f64 l5, [%1+40];\n\tld.shared.f64 l6, [%1+48];\n\tld.shared.f64 l7, [%1+56];\n\t"
      "ld.shared.f64 a0, [%2];\n\tld.shared.f64 a1, [%2+8];\n\tld.shared.f64 a2, [%2+16];\n\tld.shared.f64 a3, [%2+24];\n\t"
      "ld.shared.f64 a4, [%2+32];\n\tld.shared.f64 a5, [%2+40];\n\tld.shared.f64 a6, [%2+48];\n\tld.shared.f64 a7, [%2+56];\n\t"
      "setp.gt.s32 p0, %3, 0;\n\tsetp.gt.s32 p1, %3, 1;\n\tsetp.gt.s32 p2, %3, 2;\n\tsetp.gt.s32 p3, %3, 3;\n\t"
      "setp.gt.s32 p4, %3, 4;\n\tsetp.gt.s32 p5, %3, 5;\n\tsetp.gt.s32 p6, %3, 6;\n\tsetp.gt.s32 p7, %3, 7;\n\t"
      // the gate: ug = un with (OR of the low words of all sixteen loads) & zero folded into its low word
      "mov.b64 {x0, hi}, l0;\n\tmov.b64 {x1, hi}, l1;\n\tmov.b64 {x2, hi}, l2;\n\tmov.b64 {x3, hi}, l3;\n\t"
      "mov.b64 {x4, hi}, l4;\n\tmov.b64 {x5, hi}, l5;\n\tmov.b64 {x6, hi}, l6;\n\tmov.b64 {x7, hi}, l7;\n\t"
      "mov.b64 {y0, hi}, a0;\n\tmov.b64 {y1, hi}, a1;\n\tmov.b64 {y2, hi}, a2;\n\tmov.b64 {y3, hi}, a3;\n\t"
      "mov.b64 {y4, hi}, a4;\n\tmov.b64 {y5, hi}, a5;\n\tmov.b64 {y6, hi}, a6;\n\tmov.b64 {y7, hi}, a7;\n\t"
      "lop3.b32 acc, x0, x1, x2, 0xFE;\n\tlop3.b32 acc, acc, x3, x4, 0xFE;\n\tlop3.b32 acc, acc, x5, x6, 0xFE;\n\t"
      "lop3.b32 acc, acc, x7, y0, 0xFE;\n\tlop3.b32 acc, acc, y1, y2, 0xFE;\n\tlop3.b32 acc, acc, y3, y4, 0xFE;\n\t"
      "lop3.b32 acc, acc, y5, y6, 0xFE;\n\tor.b32 acc, acc, y7;\n\t"
      "mov.b64 {ulo, uhi}, %4;\n\t"
      "lop3.b32 ulo, ulo, acc, %5, 0xF8;\n\t"          // ulo | (acc & zero)
      "mov.b64 ug, {ulo, uhi};\n\t"
      "fma.rn.f64 a0, l0, ug, a0;\n\tfma.rn.f64 a1, l1, ug, a1;\n\tfma.rn.f64 a2, l2, ug, a2;\n\tfma.rn.f64 a3, l3, ug, a3;\n\t"
      "fma.rn.f64 a4, l4, ug, a4;\n\tfma.rn.f64 a5, l5, ug, a5;\n\tfma.rn.f64 a6, l6, ug, a6;\n\tfma.rn.f64 a7, l7, ug, a7;\n\t"
      "@p0 st.shared.f64 [%2], a0;\n\t@p1 st.shared.f64 [%2+8], a1;\n\t@p2 st.shared.f64 [%2+16], a2;\n\t@p3 st.shared.f64 [%2+24], a3;\n\t"
      "@p4 st.shared.f64 [%2+32], a4;\n\t@p5 st.shared.f64 [%2+40], a5;\n\t@p6 st.shared.f64 [%2+48], a6;\n\t@p7 st.shared.f64 [%2+56], a7;\n\t"
      "mov.f64 %0, a0;\n\t"
      "}"
      : "=d"(a0)
      : "r"(lp), "r"(ap), "r"(nleft), "d"(un), "r"(zero)
      : "memory");
  return a0;
}

// The rank-2 chunk: eight rows of this lane's column take the updates of the pivots k and k+1 at once,
//   l2[t] = c2[t] + l1[t] * un1k   (the multipliers of pivot k+1 after pivot k's update; every lane forms them itself)
//   a[t]  = (a[t] + l1[t] * un1) + l2[t] * un2
// -- the same operations in the same order as two rank-1 rounds, hence the same bits.  24 loads up front (the gate of
// slu_chunk8 over all of them, applied to the three multipliers), 24 multiply-adds, predicated stores.
// lp1 / lp2: the rows' entries in columns k and k+1; ap: the column segment.
__device__ __forceinline__ void slu_chunk8x2(const unsigned lp1, const unsigned lp2, const unsigned ap, const int nleft, const double un1k,
                                             const double un1, const double un2, const unsigned zero) {
  asm volatile(
      "{\n\t"
      ".reg .pred p0, p1, p2, p3, p4, p5, p6, p7;\n\t"
      ".reg .f64 l0, l1, l2, l3, l4, l5, l6, l7, c0, c1, c2, c3, c4, c5, c6, c7, a0, a1, a2, a3, a4, a5, a6, a7, g0, g1, g2;\n\t"
      ".reg .b32 x0, x1, x2, x3, x4, x5, x6, x7, y0, y1, y2, y3, y4, y5, y6, y7, z0, z1, z2, z3, z4, z5, z6, z7, hi, acc, lo0, hi0, lo1, hi1, lo2, hi2;\n\t"
      "ld.shared.f64 l0, [%0];\n\tld.shared.f64 l1, [%0+8];\n\tld.shared.f64 l2, [%0+16];\n\tld.shared.f64 l3, [%0+24];\n\t"
      "ld.shared.f64 l4, [%0+32];\n\tld.shared.f64 l5, [%0+40];\n\tld.shared.f64 l6, [%0+48];\n\tld.shared.f64 l7, [%0+56];\n\t"
      "ld.shared.f64 c0, [%1];\n\tld.shared.f64 c1, [%1+8];\n\tld.shared.f64 c2, [%1+16];\n\tld.shared.f64 c3, [%1+24];\n\t"
      "ld.shared.f64 c4, [%1+32];\n\tld.shared.f64 c5, [%1+40];\n\tld.shared.f64 c6, [%1+48];\n\tld.shared.f64 c7, [%1+56];\n\t"
      "ld.shared.f64 a0, [%2];\n\tld.shared.f64 a1, [%2+8];\n\tld.shared.f64 a2, [%2+16];\n\tld.shared.f64 a3, [%2+24];\n\t"
      "ld.shared.f64 a4, [%2+32];\n\tld.shared.f64 a5, [%2+40];\n\tld.shared.f64 a6, [%2+48];\n\tld.shared.f64 a7, [%2+56];\n\t"
      "setp.gt.s32 p0, %3, 0;\n\tsetp.gt.s32 p1, %3, 1;\n\tsetp.gt.s32 p2, %3, 2;\n\tsetp.gt.s32 p3, %3, 3;\n\t"
      "setp.gt.s32 p4, %3, 4;\n\tsetp.gt.s32 p5, %3, 5;\n\tsetp.gt.s32 p6, %3, 6;\n\tsetp.gt.s32 p7, %3, 7;\n\t"
      "mov.b64 {x0, hi}, l0;\n\tmov.b64 {x1, hi}, l1;\n\tmov.b64 {x2, hi}, l2;\n\tmov.b64 {x3, hi}, l3;\n\t"
      "mov.b64 {x4, hi}, l4;\n\tmov.b64 {x5, hi}, l5;\n\tmov.b64 {x6, hi}, l6;\n\tmov.b64 {x7, hi}, l7;\n\t"
      "mov.b64 {y0, hi}, c0;\n\tmov.b64 {y1, hi}, c1;\n\tmov.b64 {y2, hi}, c2;\n\tmov.b64 {y3, hi}, c3;\n\t"
      "mov.b64 {y4, hi}, c4;\n\tmov.b64 {y5, hi}, c5;\n\tmov.b64 {y6, hi}, c6;\n\tmov.b64 {y7, hi}, c7;\n\t"
      "mov.b64 {z0, hi}, a0;\n\tmov.b64 {z1, hi}, a1;\n\tmov.b64 {z2, hi}, a2;\n\tmov.b64 {z3, hi}, a3;\n\t"
      "mov.b64 {z4, hi}, a4;\n\tmov.b64 {z5, hi}, a5;\n\tmov.b64 {z6, hi}, a6;\n\tmov.b64 {z7, hi}, a7;\n\t"
      "lop3.b32 acc, x0, x1, x2, 0xFE;\n\tlop3.b32 acc, acc, x3, x4, 0xFE;\n\tlop3.b32 acc, acc, x5, x6, 0xFE;\n\t"
      "lop3.b32 acc, acc, x7, y0, 0xFE;\n\tlop3.b32 acc, acc, y1, y2, 0xFE;\n\tlop3.b32 acc, acc, y3, y4, 0xFE;\n\t"
      "lop3.b32 acc, acc, y5, y6, 0xFE;\n\tlop3.b32 acc, acc, y7, z0, 0xFE;\n\tlop3.b32 acc, acc, z1, z2, 0xFE;\n\t"
      "lop3.b32 acc, acc, z3, z4, 0xFE;\n\tlop3.b32 acc, acc, z5, z6, 0xFE;\n\tor.b32 acc, acc, z7;\n\t"
      "mov.b64 {lo0, hi0}, %4;\n\tmov.b64 {lo1, hi1}, %5;\n\tmov.b64 {lo2, hi2}, %6;\n\t"
      "lop3.b32 lo0, lo0, acc, %7, 0xF8;\n\tlop3.b32 lo1, lo1, acc, %7, 0xF8;\n\tlop3.b32 lo2, lo2, acc, %7, 0xF8;\n\t"
      "mov.b64 g0, {lo0, hi0};\n\tmov.b64 g1, {lo1, hi1};\n\tmov.b64 g2, {lo2, hi2};\n\t"
      "fma.rn.f64 c0, l0, g0, c0;\n\tfma.rn.f64 c1, l1, g0, c1;\n\tfma.rn.f64 c2, l2, g0, c2;\n\tfma.rn.f64 c3, l3, g0, c3;\n\t"
      "fma.rn.f64 c4, l4, g0, c4;\n\tfma.rn.f64 c5, l5, g0, c5;\n\tfma.rn.f64 c6, l6, g0, c6;\n\tfma.rn.f64 c7, l7, g0, c7;\n\t"
      "fma.rn.f64 a0, l0, g1, a0;\n\tfma.rn.f64 a1, l1, g1, a1;\n\tfma.rn.f64 a2, l2, g1, a2;\n\tfma.rn.f64 a3, l3, g1, a3;\n\t"
      "fma.rn.f64 a4, l4, g1, a4;\n\tfma.rn.f64 a5, l5, g1, a5;\n\tfma.rn.f64 a6, l6, g1, a6;\n\tfma.rn.f64 a7, l7, g1, a7;\n\t"
      "fma.rn.f64 a0, c0, g2, a0;\n\tfma.rn.f64 a1, c1, g2, a1;\n\tfma.rn.f64 a2, c2, g2, a2;\n\tfma.rn.f64 a3, c3, g2, a3;\n\t"
      "fma.rn.f64 a4, c4, g2, a4;\n\tfma.rn.f64 a5, c5, g2, a5;\n\tfma.rn.f64 a6, c6, g2, a6;\n\tfma.rn.f64 a7, c7, g2, a7;\n\t"
      "@p0 st.shared.f64 [%2], a0;\n\t@p1 st.shared.f64 [%2+8], a1;\n\t@p2 st.shared.f64 [%2+16], a2;\n\t@p3 st.shared.f64 [%2+24], a3;\n\t"
      "@p4 st.shared.f64 [%2+32], a4;\n\t@p5 st.shared.f64 [%2+40], a5;\n\t@p6 st.shared.f64 [%2+48], a6;\n\t@p7 st.shared.f64 [%2+56], a7;\n\t"
      "}"
      :
      : "r"(lp1), "r"(lp2), "r"(ap), "r"(nleft), "d"(un1k), "d"(un1), "d"(un2), "r"(zero)
      : "memory");
}

// ---- the pivot rounds, one pivot per round (tables without pair alignment: XB_SLU_PAIR=0) -----------------------------
__device__ __forceinline__ void slu_rounds1(double *__restrict__ S, const unsigned sS, const int tid, const int lane, const int w,
                                            int &kfail, bool &viol) {
  const DevSlu &D = c_net.slu;
  const int nwl = D.nwl;
  const double itau = D.itau;
  // loop constants pinned in registers (an opaque move: otherwise they are re-read from the constant bank every round)
  int nround = D.nround, nfl = nwl * 32, nwr = nwl;
  asm volatile("mov.b32 %0, %0;\n\tmov.b32 %1, %1;\n\tmov.b32 %2, %2;" : "+r"(nround), "+r"(nfl), "+r"(nwr));
  const unsigned zero = (unsigned)nfl >> 20;         // 0, but not to ptxas (nfl is pinned above): the load gate of slu_chunk8
  // Schedule entries travel L2 -> registers one loop trip (three rounds) ahead of their use: a rotation through
  // register moves per round would have to read a freshly loaded register at the end of the same round.  The host pads
  // the schedule to a multiple of three rounds plus three idle ones.
  const unsigned *__restrict__ lt = D.sch_lane + tid;
  const uint2 *__restrict__ wt = D.sch_warp + w;
  unsigned eA = __ldg(lt), eB = __ldg(lt + nfl), eC = __ldg(lt + 2 * nfl);
  uint2 gA = __ldg(wt), gB = __ldg(wt + nwr), gC = __ldg(wt + 2 * nwr);
  lt += 3 * nfl; wt += 3 * nwr;
  const int row = lane & 7;
  double rprev = 0.0; unsigned dprev = 0xffffffffu;   // reciprocal of the previous pivot, stored over it once the schedule has left that pivot
#ifdef XB_SLU_PROF
  long long pt[5] = {0, 0, 0, 0, 0}, pc0 = clock64();
#define SLU_PT(i) { const long long now_ = clock64(); pt[i] += now_ - pc0; pc0 = now_; }
#else
#define SLU_PT(i)
#endif
  auto round = [&](const unsigned e, const uint2 g) {
    const unsigned lp = g.x & 0xffffu, pv = g.x >> 16, ap = e & 0xffffu, up = e >> 16;
    const int nrow = (int)(g.y & 15u), ntest = (int)((g.y >> 4) & 15u), kk = (int)(g.y >> 8);
    SLU_PT(4)
    bar_named(1, nfl);
    SLU_PT(0)
    const double piv = S[pv];
    const double u = S[up];
    const double tv = S[lp + row];                      // this lane's row of the chunk for the column test
    if (tid == 0 && dprev != pv && dprev != 0xffffffffu) S[dprev] = rprev;
    // reciprocal of the pivot without a range branch: a pivot outside 1e-280..1e280 (or zero, or NaN) fails the
    // factorisation below, and the dense fallback redoes it with true divisions
    double rk;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rk) : "d"(piv));
    {
      double er = fma(-piv, rk, 1.0);                   // seed good to ~20 bits; r (1 + e + e^2) leaves e^3 < 2^-58
      er = fma(er, er, er);
      rk = fma(rk, er, rk);
    }
    const double un = -(u * rk);
#ifdef XB_SLU_PROF
    if (__double2hiint(un) == 0x7ff12345) viol = true;   // consume u: the time stamp below waits for the operand loads
    SLU_PT(1)
#endif
    slu_chunk8(sS + 8u * lp, sS + 8u * ap, (ap != lp) ? nrow : 0, un, zero);
    SLU_PT(2)
    // column test, lanes = rows of this warp's chunk: |a(i,k)| > thr  <=>  tau |a(i,k)| > |a(k,k)|
    const double apiv = fabs(piv), thr = apiv * itau;
    const bool bad = ((row < ntest) & (fabs(tv) > thr)) | !((apiv > 1e-280) & (apiv < 1e280));
    kfail = min(kfail, bad ? kk : 0x7fffffff);
    rprev = rk; dprev = pv;
    SLU_PT(3)
  };
#pragma unroll 1
  for (int r = 0; r < nround; r += 3) {
    // this trip's entries move to their own registers and all three loads for the next trip are issued at once: whatever
    // register shuffling the compiler does at the loop's back edge then touches values that were requested three
    // rounds earlier (a load issued in the trip's last round would be waited for at the back edge)
    const unsigned fA = eA, fB = eB, fC = eC;
    const uint2 hA = gA, hB = gB, hC = gC;
    eA = __ldg(lt); eB = __ldg(lt + nfl); eC = __ldg(lt + 2 * nfl);
    gA = __ldg(wt); gB = __ldg(wt + nwr); gC = __ldg(wt + 2 * nwr);
    lt += 3 * nfl; wt += 3 * nwr;
    round(fA, hA); round(fB, hB); round(fC, hC);
  }
#ifdef XB_SLU_PROF
  if (lane == 0 && w == 0) {
    long long *ph = reinterpret_cast<long long *>(S) + D.ph_rel;
    for (int i = 0; i < 5; ++i) ph[PH_DBG0 + 4 + i] += pt[i];
  }
#endif
  bar_named(1, nfl);
  if (tid == 0 && dprev != 0xffffffffu) S[dprev] = rprev;
}

// ---- the pivot rounds, two pivots per round (pair-aligned tables, SluSchedule2 in xb_host.h) -----------------------------
// One barrier, one pass over the rows for the pivots k and k+1: the second pivot a(k+1,k+1) - l(k+1,k) u(k,k+1) and its
// reciprocal are formed by every thread from four broadcast loads, every lane forms U(k+1,j) for its own column (and
// from it the second multiplier), every lane forms the multipliers of pivot k+1 for the rows of its chunk.  Row k+1 of U
// is stored one barrier later by the lanes of the item the schedule flags (other warps of the same round still read the
// old row); the lane of column k+1 leaves the updated multipliers behind; thread 0 stores both reciprocals once the
// schedule has left the pair.  Arithmetic identical to two rank-1 rounds, operation by operation.
__device__ __forceinline__ void slu_rounds2(double *__restrict__ S, const unsigned sS, const int tid, const int lane, const int w, int &kfail) {
  const DevSlu &D = c_net.slu;
  const double itau = D.itau;
  int nround = D.nround, nfl = D.nwl * 32, nwr = D.nwl;
  asm volatile("mov.b32 %0, %0;\n\tmov.b32 %1, %1;\n\tmov.b32 %2, %2;" : "+r"(nround), "+r"(nfl), "+r"(nwr));
  const unsigned zero = (unsigned)nfl >> 20;         // 0, but not to ptxas: the load gate of the chunk
  const uint2 *__restrict__ lt = D.sch2_lane + tid;
  const uint4 *__restrict__ wt = D.sch2_warp + w;
  uint2 eA = __ldg(lt), eB = __ldg(lt + nfl), eC = __ldg(lt + 2 * nfl);
  uint4 gA = __ldg(wt), gB = __ldg(wt + nwr), gC = __ldg(wt + 2 * nwr);
  lt += 3 * nfl; wt += 3 * nwr;
  const int row = lane & 7;
  const unsigned none = 0xffffffffu;
  double r1prev = 0.0, r2prev = 0.0, pend_val = 0.0;
  unsigned d1prev = none, d2prev = none, pend_ad = none;
  auto rcp = [](const double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double er = fma(-x, r, 1.0);                       // seed good to ~20 bits; r (1 + e + e^2) leaves e^3 < 2^-58
    er = fma(er, er, er);
    return fma(r, er, r);
  };
  auto round = [&](const uint2 e, const uint4 g) {
    const unsigned lp1 = g.x & 0xffffu, lp2 = g.x >> 16, pv1 = g.z & 0xffffu, u12 = g.z >> 16;
    const unsigned ap = e.x & 0xffffu, up1 = e.x >> 16, up2 = e.y & 0xffffu;
    const bool col2 = (e.y >> 16) & 1u;
    const int nrow = (int)(g.y & 15u), nt1 = (int)((g.y >> 4) & 15u), nt2 = (int)((g.y >> 8) & 15u), kk = (int)g.w;
    const bool commit = (g.y >> 12) & 1u, single = (g.y >> 13) & 1u;
    bar_named(1, nfl);
    const double piv1 = S[pv1], l21 = S[pv1 + 1], u12v = S[u12], piv2 = S[u12 + 1];
    const double u1 = S[up1], a2 = S[up2];
    const double tv1 = S[lp1 + row], tc2 = S[lp2 + row];  // this lane's row of the chunk for the two column tests
    if (pend_ad != none) S[pend_ad] = pend_val;
    const bool st1 = tid == 0 && d1prev != pv1 && d1prev != none, st2 = st1 && d2prev != none;
    if (st1) S[d1prev] = r1prev;
    if (st2) S[d2prev] = r2prev;
    const double rk1 = rcp(piv1);
    const double un1k = -(u12v * rk1);
    const double p2 = fma(l21, un1k, piv2);
    const double rk2f = rcp(p2);                         // a single's second pivot is switched off: integer selects, no branch
    const double rk2 = __hiloint2double(single ? 0 : __double2hiint(rk2f), single ? 0 : __double2loint(rk2f));
    const double un1 = -(u1 * rk1);
    const double upr = fma(l21, un1, a2);               // U(k+1, j)
    const double un2 = col2 ? 0.0 : -(upr * rk2);
    const bool active = ap != lp1;
    slu_chunk8x2(sS + 8u * lp1, sS + 8u * lp2, sS + 8u * ap, active ? nrow : 0, un1k, un1, un2, zero);
    // column tests, lanes = rows of this warp's chunk; a pivot outside 1e-280..1e280 (or zero, or NaN) fails
    const double ap1 = fabs(piv1), thr1 = ap1 * itau, ap2 = fabs(p2), thr2 = ap2 * itau;
    const double l2t = fma(tv1, un1k, tc2);
    const bool bad1 = ((row < nt1) & (fabs(tv1) > thr1)) | (!single & (fabs(l21) > thr1)) | !((ap1 > 1e-280) & (ap1 < 1e280));
    const bool bad2 = ((row < nt2) & (fabs(l2t) > thr2)) | (!single & !((ap2 > 1e-280) & (ap2 < 1e280)));
    kfail = min(kfail, bad1 ? kk : 0x7fffffff);
    kfail = min(kfail, bad2 ? kk + 1 : 0x7fffffff);
    pend_ad = (commit & active & !col2) ? up2 : none;
    pend_val = upr;
    r1prev = rk1; r2prev = rk2; d1prev = pv1; d2prev = single ? none : u12 + 1u;
  };
#pragma unroll 1
  for (int r = 0; r < nround; r += 3) {
    const uint2 fA = eA, fB = eB, fC = eC;
    const uint4 hA = gA, hB = gB, hC = gC;
    eA = __ldg(lt); eB = __ldg(lt + nfl); eC = __ldg(lt + 2 * nfl);
    gA = __ldg(wt); gB = __ldg(wt + nwr); gC = __ldg(wt + 2 * nwr);
    lt += 3 * nfl; wt += 3 * nwr;
    round(fA, hA); round(fB, hB); round(fC, hC);
  }
  bar_named(1, nfl);
  if (pend_ad != none) S[pend_ad] = pend_val;
  if (tid == 0 && d1prev != none) { S[d1prev] = r1prev; if (d2prev != none) S[d2prev] = r2prev; }
}

// ---- factorisation -----------------------------------------------------------------------------------------------
// Executed by the first nwl warps of the CTA (the others return at once; the caller's __syncthreads joins them).
// The right-hand side (S + rhs_off, position order) is carried along as an extra column: it leaves as L^-1 P b.
// ZoneScalars::slu_fail is set when a heavy column violates the threshold test or a pivot is zero/NaN, slu_fail_k
// receives the first such column (caller initialises it to a large value), slu_ipb the interchanges of the border
// block.  On return the diagonal of U holds the RECIPROCALS of the pivots.
//
// The pivot loop is latency bound (a zone has four warps; an SM three zones): what counts is the number of
// instructions and above all of branches on a warp's path through one pivot.  Nothing about WHERE a pivot works
// depends on the data, so all of it -- which column a lane owns, whether that column reaches the pivot row, where its
// segment and its u(k,j) live, how many rows the warp's chunk has and how many of them are tested -- comes from the
// pivot schedule the host derives once per network (structured_lu_schedule, xb_loader.cpp): 4 bytes per lane and 8 per
// warp and round, zone-invariant, streamed from L2 two rounds ahead of their use.  A round is then: barrier, pivot,
// reciprocal, u, one chunk (slu_chunk8), the column test of the chunk's own rows (lanes = rows, folded into every warp:
// no warp is singled out), flat control flow.  A pivot with more chunks than warps or more columns than lanes takes
// further rounds; they recompute the reciprocal from the untouched pivot (thread 0 stores a reciprocal over its
// pivot only once the schedule has moved on).
__device__ __noinline__ void slu_factor(double *__restrict__ S) {
  XB_ASSUME_SHARED(S);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const DevSlu &D = c_net.slu;
  const int nwl = D.nwl;
  if (w >= nwl) return;
  const int n = D.n, nb = D.nbord, nh = n - nb, bo = D.bo, bs = D.bs;
  double *rhs = S + D.rhs_off;
  int *fail = reinterpret_cast<int *>(S) + D.fail_rel, *ipb = fail + 2;      // ZoneScalars::slu_fail, slu_fail_k, slu_ipb[8]
  const unsigned sS = (unsigned)__cvta_generic_to_shared(S);
  // The temperature equation (self-heating: the last row) is exempt from the test: it is the linear functional
  // -sum_i mex_i (row i) / cv of the abundance rows, i.e. a row in different units whose multipliers measure mex/cv and
  // not element growth (elimination without interchanges is invariant under row scaling).
  const double tau = D.tau, itau = D.itau;
  const int nbt = D.ntest - nh;                      // border rows subject to the test
  bool viol = false;
  // the border entries of the right-hand side ride in column n of the border block (same address rule as the matrix columns)
  if (w == 0 && lane < nb) S[bo + n * bs + lane] = rhs[nh + lane];
  int kfail = 0x7fffffff;
  if (D.pair) slu_rounds2(S, sS, tid, lane, w, kfail);
  else slu_rounds1(S, sS, tid, lane, w, kfail, viol);
  if (kfail != 0x7fffffff) { viol = true; atomicMin(fail + 1, kfail); }
  // ---- border block, one warp, in place in shared memory; lane c works on column c of the block (lane nb: the right-hand
  // side).  The diagonal is the pivot while it passes the threshold test against the tested rows below it; otherwise the
  // largest entry of the column (any row) is brought up.  Rolled loops over rows and pivots: the block is 5-7 wide, and
  // what counts for a once-per-factorisation piece of code is its instruction footprint (a register-resident, fully
  // unrolled version was 4400 instructions, 28 KB of them executed per factorisation, against 250 here).
  if (w == 0) {
    const int c = lane;
    const bool mine = c <= nb;
    const int colc = bo + ((c < nb) ? nh + c : n) * bs;
#pragma unroll 1
    for (int jj = 0; jj < nb; ++jj) {
      const int colj = bo + (nh + jj) * bs;
      // every lane scans column jj itself (broadcast loads, the same decision in all lanes): largest entry from the
      // diagonal down (first maximum) and the largest among the tested rows
      double best = -1.0, bestt = -1.0; int p = jj;
#pragma unroll 1
      for (int r = jj; r < nb; ++r) {
        const double av = fabs(S[colj + r]);
        if (av > best) { best = av; p = r; }
        if (r < nbt && av > bestt) bestt = av;
      }
      const double ad = fabs(S[colj + jj]);
      if (ad > 0.0 && !(tau * bestt > ad)) p = jj;
      __syncwarp();
      if (lane == 0) ipb[jj] = p;
      if (p != jj && mine) { const double t = S[colc + jj]; S[colc + jj] = S[colc + p]; S[colc + p] = t; }
      __syncwarp();
      const double pv = S[colj + jj];
      if (!(fabs(pv) > 0.0)) viol = true;
      const double rk = pivot_rcp<false>(pv);
      if (mine && c > jj) {
        const double ujc = S[colc + jj];
#pragma unroll 1
        for (int r = jj + 1; r < nb; ++r) { const double lr = S[colj + r] * rk; S[colc + r] = fma(-lr, ujc, S[colc + r]); }
      }
      __syncwarp();
      if (lane == jj) S[colj + jj] = rk;                 // reciprocal over the pivot
    }
    __syncwarp();
    if (lane < nb) rhs[nh + lane] = S[bo + n * bs + lane];
    // the same interchanges on the heavy-column part of the border rows (their multipliers), in order
#pragma unroll 1
    for (int jj = 0; jj < nb; ++jj) {
      const int pj = ipb[jj];
      if (pj != jj) {
        const int la = bo + jj, lb = bo + pj;
        for (int cc = lane; cc < nh; cc += 32) { const double ta = S[la + cc * bs], tb = S[lb + cc * bs]; S[la + cc * bs] = tb; S[lb + cc * bs] = ta; }
      }
    }
  }
  if (viol) *fail = 1;
}

// ---- triangular solves (all threads of the CTA enter) ------------------------------------------------------------------
// fwd_done: b already holds L^-1 P b (it rode through slu_factor).  Otherwise: forward substitution with the unscaled
// multipliers (warp 0), border interchanges first.  Then U x = y (the diagonal of U holds the reciprocals of the pivots):
//   A  the border unknowns (warp 0, nb dependent steps);
//   B  their columns are full: every heavy row takes its nb updates at once (all threads);
//   C  the band, last column first (warp 0).  The 32 unknowns below the current one live in REGISTERS (lane = row mod
//      32), x(k) travels by shuffle, U's entries and the reciprocal are loaded one step ahead: the dependent chain of a
//      step is shuffle - multiply - multiply-add, no shared-memory round trip.  Rows further than 31 above the diagonal
//      (a few outlier columns) are updated in shared memory before they enter the window.
__device__ __forceinline__ int slu_diag(const SluIdx &I, const int k) {
  return (k < I.nh) ? slu_cb(I, k) + k : I.bo + k * I.bs + (k - I.nh);
}
__device__ __noinline__ void slu_solve(const double *__restrict__ S, double *__restrict__ b, const int *ipb, const bool fwd_done) {
  XB_ASSUME_SHARED(S); XB_ASSUME_SHARED(b); XB_ASSUME_SHARED(ipb);
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  const bool w0 = tid < 32;
  const SluIdx I = slu_idx();
  const int n = I.n, nh = I.nh, nb = I.nb, bo = I.bo, bs = I.bs;
  if (w0) {
    if (!fwd_done) {
      // interchanges of the border block first (DLASWP precedes the forward substitution: the heavy pivots' updates of
      // the border entries use the multipliers of the rows in their final slots)
      if (lane == 0) {
        for (int jj = 0; jj < nb; ++jj) {
          const int p = ipb[jj];
          if (p != jj) { const double t = b[nh + jj]; b[nh + jj] = b[nh + p]; b[nh + p] = t; }
        }
      }
      __syncwarp();
      for (int k = 0; k < n; ++k) {
        const double f = b[k] * S[slu_diag(I, k)];
        if (k < nh) {
          const int nrb = slu_nrb(I, k), cbk = slu_cb(I, k);
          for (int r = lane; r < nrb; r += 32) b[k + 1 + r] = fma(-S[cbk + k + 1 + r], f, b[k + 1 + r]);
          if (lane < nb) b[nh + lane] = fma(-S[bo + k * bs + lane], f, b[nh + lane]);
        } else {
          const int i = k + 1 + lane;
          if (i < n) b[i] = fma(-S[bo + k * bs + (i - nh)], f, b[i]);
        }
        __syncwarp();
      }
    }
    // A: border unknowns
    for (int k = n - 1; k >= nh; --k) {
      const double x = b[k] * S[bo + k * bs + (k - nh)];
      __syncwarp();
      if (lane == 0) b[k] = x;
      const int i = nh + lane;
      if (i < k) b[i] = fma(-S[bo + k * bs + lane], x, b[i]);
      __syncwarp();
    }
  }
  __syncthreads();
  // B: heavy rows minus the border columns times the border unknowns
  for (int i = tid; i < nh; i += nt) {
    double acc = b[i];
#pragma unroll
    for (int cI = 0; cI < kSluBord; ++cI)
      if (cI < nb) acc = fma(-S[slu_cb(I, nh + cI) + i], b[nh + cI], acc);
    b[i] = acc;
  }
  __syncthreads();
  // C: the band, from a per-step, per-lane table (sol_lane: four 16-bit entries per lane and group of four steps, read
  // from L2 two groups ahead): the address of U(i,k) for the window row i = k - ((k - lane) & 31) this lane holds at
  // step k, the address of the reciprocal of the pivot for the lane that owns row k, 0xffff = structurally zero.
  if (w0 && nh > 0) {
    const DevSlu &D = c_net.slu;
    const uint2 *__restrict__ st = D.sol_lane + lane;
    const int *__restrict__ outl = D.sol_out;          // steps with rows above the window, descending, then a negative end mark
    uint2 qA = __ldg(st), qB = __ldg(st + 32);
    st += 64;
    int nxt = __ldg(outl), oi = 1;
    int k = nh - 1;
    double bv, be = 0.0;
    { const int i = k - ((k - lane) & 31); bv = (i >= 0) ? b[i] : 0.0; }
    bool enter = false;                                // this lane owned the previous step: its next window row (loaded then) enters now
    const unsigned sB = (unsigned)__cvta_generic_to_shared(b);
    auto step = [&](const double v) {
      const bool own = lane == (k & 31);
      const double x = __shfl_sync(kFull, bv * v, k & 31);      // owner: b(k) times the reciprocal of the pivot
      if (enter) bv = be;
      if (k == nxt) {                                  // outlier column: rows above the window, in shared memory
        const unsigned t = I.col[k].x;
        const int cfk = (int)(t >> 16), cbk = (int)(t & 0xffffu);
        for (int i = cfk + lane; i <= k - 32; i += 32) b[i] = fma(-S[cbk + i], x, b[i]);
        __syncwarp();
        nxt = __ldg(outl + oi); ++oi;
      }
      bv = fma(-v, x, bv);
      // the owner stores x(k) and fetches the row that takes its place in the window, b(k-32): predicated, no branch
      // (k < 32: a harmless read of b(0); rows above the first never own a step and meet only zero entries)
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "setp.ne.s32 p, %3, 0;\n\t"
          "@p st.shared.f64 [%1], %4;\n\t"
          "@p ld.shared.f64 %0, [%2];\n\t}"
          : "+d"(be)
          : "r"(sB + 8u * (unsigned)max(k, 0)), "r"(sB + 8u * (unsigned)max(k - 32, 0)), "r"((int)(own && k >= 0)), "d"(x)
          : "memory");
      enter = own;
      --k;
    };
    auto group = [&](uint2 &q) {
      const unsigned a0 = q.x & 0xffffu, a1 = q.x >> 16, a2 = q.y & 0xffffu, a3 = q.y >> 16;
      q = __ldg(st); st += 32;                         // two groups ahead (the table ends with two idle groups)
      // U's entries of the four steps before the first of them runs (S is not written here)
      const double v0 = (a0 != 0xffffu) ? S[a0] : 0.0, v1 = (a1 != 0xffffu) ? S[a1] : 0.0;
      const double v2 = (a2 != 0xffffu) ? S[a2] : 0.0, v3 = (a3 != 0xffffu) ? S[a3] : 0.0;
      step(v0); step(v1); step(v2); step(v3);
    };
    const int ngroup = D.sol_ngroup;                   // even
#pragma unroll 1
    for (int g = 0; g < ngroup; g += 2) { group(qA); group(qB); }
  }
}

// ---- dense fallback ------------------------------------------------------------------------------------------------
// The skyline matrix (freshly assembled) is expanded into the CTA's global scratch A[n+1 columns][n] (column n = the
// right-hand side) and factored by a plain right-looking LU with partial pivoting, all threads; multipliers are
// stored scaled, interchanges in ipiv (global, n ints behind the matrix).  Rare path: written for size, not speed.
__device__ __noinline__ void slu_dense_factor(const double *__restrict__ S, const double *__restrict__ rhs, double *__restrict__ A,
                                              int *__restrict__ ipiv, double *red /* shared, >= 66 doubles */, int *info) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wq = tid >> 5, nw = nt >> 5;
  const SluIdx I = slu_idx();
  const int n = I.n;
  for (int e = tid; e < n * n; e += nt) A[e] = 0.0;
  __syncthreads();
  for (int p = wq; p < n; p += nw) {
    const int c0 = slu_cf(I, p), cbp = slu_cb(I, p), nh = I.nh, c1 = slu_cbot(I, p);
    for (int i = c0 + lane; i <= c1; i += 32) A[(size_t)p * n + i] = S[cbp + i];
    if (lane < I.nb) A[(size_t)p * n + nh + lane] = S[I.bo + p * I.bs + lane];
  }
  for (int i = tid; i < n; i += nt) A[(size_t)n * n + i] = rhs[i];
  if (tid == 0) *info = 0;
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    // IDAMAX of column k from the diagonal down (first maximum)
    double bv = -1.0; int bi = 0x7fffffff;
    for (int i = k + tid; i < n; i += nt) { const double v = fabs(A[(size_t)k * n + i]); if (v > bv) { bv = v; bi = i; } }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_xor_sync(kFull, bv, off);
      const int oi = __shfl_xor_sync(kFull, bi, off);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red[wq] = bv; reinterpret_cast<int *>(red + 32)[wq] = bi; }
    __syncthreads();
    double pv = red[0]; int p = reinterpret_cast<int *>(red + 32)[0];
    for (int q = 1; q < nw; ++q) {
      const double ov = red[q]; const int oi = reinterpret_cast<int *>(red + 32)[q];
      if (ov > pv || (ov == pv && oi < p)) { pv = ov; p = oi; }
    }
    if (!(pv > 0.0)) { if (tid == 0) { ipiv[k] = k; if (*info == 0) *info = k + 1; } __syncthreads(); continue; }
    if (tid == 0) ipiv[k] = p;
    if (p != k)
      for (int j = tid; j <= n; j += nt) { double *col = A + (size_t)j * n; const double a = col[k], b2 = col[p]; col[k] = b2; col[p] = a; }
    __syncthreads();
    const double rk = 1.0 / A[(size_t)k * n + k];
    __syncthreads();
    for (int i = k + 1 + tid; i < n; i += nt) A[(size_t)k * n + i] *= rk;
    __syncthreads();
    // four columns per warp pass, two row chunks per iteration: up to 10 loads in flight per lane (L2 latency)
    for (int j0 = k + 1 + 4 * wq; j0 <= n; j0 += 4 * nw) {
      double *cp[4]; double u[4]; bool on[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) { on[t] = j0 + t <= n; cp[t] = A + (size_t)(on[t] ? j0 + t : j0) * n; u[t] = cp[t][k]; }
      for (int i = k + 1 + lane; i < n; i += 64) {
        const int i2 = i + 32; const bool two = i2 < n;
        const double l0 = A[(size_t)k * n + i], l1 = two ? A[(size_t)k * n + i2] : 0.0;
        double c0[4], c1[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { c0[t] = cp[t][i]; c1[t] = two ? cp[t][i2] : 0.0; }
#pragma unroll
        for (int t = 0; t < 4; ++t) { c0[t] = fma(-l0, u[t], c0[t]); c1[t] = fma(-l1, u[t], c1[t]); }
#pragma unroll
        for (int t = 0; t < 4; ++t) if (on[t]) { cp[t][i] = c0[t]; if (two) cp[t][i2] = c1[t]; }
      }
    }
    __syncthreads();
  }
}
// b (shared, position order): fwd_done = the forward-solved column n of A is the right-hand side; else b is permuted
// and forward-solved here.  Result in b.
__device__ __noinline__ void slu_dense_solve(double *__restrict__ A, const int *__restrict__ ipiv, double *__restrict__ b,
                                             const bool fwd_done) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const SluIdx I = slu_idx();
  const int n = I.n;
  double *y = A + (size_t)n * n;
  if (!fwd_done) {
    if (tid == 0)
      for (int k = 0; k < n; ++k) { const int p = ipiv[k]; if (p != k) { const double t = b[k]; b[k] = b[p]; b[p] = t; } }
    __syncthreads();
    for (int i = tid; i < n; i += nt) y[i] = b[i];
    __syncthreads();
    for (int k = 0; k < n; ++k) {
      const double yk = y[k];
      __syncthreads();
      for (int i = k + 1 + tid; i < n; i += nt) y[i] = fma(-A[(size_t)k * n + i], yk, y[i]);
      __syncthreads();
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    const double d = A[(size_t)k * n + k];
    const double x = (d != 0.0) ? y[k] / d : 0.0;
    __syncthreads();
    if (tid == 0) y[k] = x;
    for (int i = tid; i < k; i += nt) y[i] = fma(-A[(size_t)k * n + i], x, y[i]);
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt) b[i] = y[i];
  __syncthreads();
}

}  // namespace xb
