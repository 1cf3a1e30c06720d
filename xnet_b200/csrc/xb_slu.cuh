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
// latency, so the pivot loop is written for instruction count: ~70 per warp and pivot.
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
// (The heavy-column test compares the upper 32 bits of the doubles: a relative slack of 2^-20 on the threshold.)
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

// Shared-memory accesses of the elimination's row loops as VOLATILE PTX with compile-time offsets: volatile accesses keep
// their program order through ptxas, which otherwise sinks each load to its use (load - multiply-add - store, one
// element after the other: eight exposed shared-memory latencies per chunk instead of one).
template <int OFF>
__device__ __forceinline__ double slu_lds(const unsigned addr) {
  double v;
  asm volatile("ld.volatile.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ void slu_sts(const unsigned addr, const double v, const int on) {     // predicated store: no branch
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p st.volatile.shared.f64 [%0+%1], %2;\n\t}" ::"r"(addr), "n"(OFF), "d"(v), "r"(on));
}
template <int T, int N>
struct SluRow {     // rows T..N-1 of a chunk: all loads, then all multiply-adds, then the stores of the first nleft rows
  static __device__ __forceinline__ void load(double (&l)[N], double (&a)[N], const unsigned lp, const unsigned ap) {
    l[T] = slu_lds<8 * T>(lp); a[T] = slu_lds<8 * T>(ap);
    SluRow<T + 1, N>::load(l, a, lp, ap);
  }
  static __device__ __forceinline__ void store(const double (&a)[N], const unsigned ap, const int nleft) {
    slu_sts<8 * T>(ap, a[T], T < nleft);
    SluRow<T + 1, N>::store(a, ap, nleft);
  }
};
template <int N>
struct SluRow<N, N> {
  static __device__ __forceinline__ void load(double (&)[N], double (&)[N], const unsigned, const unsigned) {}
  static __device__ __forceinline__ void store(const double (&)[N], const unsigned, const int) {}
};

__device__ __forceinline__ void slu_sync(const int nwl) {
  if (nwl == 1) __syncwarp(); else bar_named(1, nwl * 32);
}

// ---- factorisation -----------------------------------------------------------------------------------------------
// Executed by the first nwl warps of the CTA (the others return at once; the caller's __syncthreads joins them).
// with_rhs: the right-hand side (S + rhs_off, position order) is carried along as an extra column: it leaves as
// L^-1 P b.  ZoneScalars::slu_fail is set when a heavy column violates the threshold test or a pivot is zero/NaN,
// slu_fail_k receives the first such column (caller initialises it to a large value), slu_ipb the interchanges of the
// border block; the reciprocals of the pivots go to rdiag.
__device__ __noinline__ void slu_factor(double *__restrict__ S, const bool with_rhs) {
  XB_ASSUME_SHARED(S);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const DevSlu &D = c_net.slu;
  const int nwl = D.nwl;
  if (w >= nwl) return;
  const int n = D.n, nb = D.nbord, nh = n - nb, bo = D.bo, bs = D.bs;
  // everything the loop touches is addressed relative to S (one shared-memory base): the index table, the reciprocals
  // of the pivots (rdiag), the right-hand side
  const uint2 *__restrict__ tcol = reinterpret_cast<const uint2 *>(S) + D.tab_rel;
  double *rhs = S + D.rhs_off, *rinv = S + D.rinv_off;
  int *fail = reinterpret_cast<int *>(S) + D.fail_rel, *ipb = fail + 2;      // ZoneScalars::slu_fail, slu_fail_k, slu_ipb[8]
  const unsigned sS = (unsigned)__cvta_generic_to_shared(S);
  // The temperature equation (self-heating: the last row) is exempt from the test: it is the linear functional
  // -sum_i mex_i (row i) / cv of the abundance rows, i.e. a row in different units whose multipliers measure mex/cv and
  // not element growth (elimination without interchanges is invariant under row scaling).
  const double tau = D.tau, itau = D.itau;
  const int nbt = D.ntest - nh;                      // border rows subject to the test
  const int ncx = nb + (with_rhs ? 1 : 0);           // border columns (+ right-hand side) after the band columns
  bool viol = false;
  // the border entries of the right-hand side ride in column n of the border block (same address rule as the matrix columns)
  if (with_rhs && w == 0 && lane < nb) S[bo + n * bs + lane] = rhs[nh + lane];
  for (int k = 0; k < nh; ++k) {
    slu_sync(nwl);
    const uint2 tk = tcol[k];
    const int ncb = (int)(tk.y & 0xffu), nrb = (int)((tk.y >> 8) & 0xffu);
    const int cbk = (int)(tk.x & 0xffffu);
    const int lband = cbk + k + 1, lbord = bo + k * bs;   // multipliers (unscaled): band rows, border rows
    const double piv = S[lband - 1];
    const double rk = pivot_rcp<false>(piv);
    // |a(i,k)| > thr  <=>  tau |a(i,k)| > |a(k,k)|, compared on the upper words; a zero or NaN pivot fails every row
    const double thr = fabs(piv) * itau;
    const int thr_hi = (fabs(piv) > 0.0) ? __double2hiint(thr) : -1;
    int lmax = 0;                                    // largest upper word of |multiplier| this thread has seen
    if (tid == 0) rinv[k] = rk;
    const int ncols = ncb + ncx;
    const int nchunk = (nrb + 7) >> 3;               // chunks of band rows; chunk number nchunk = the border rows
#pragma unroll 1
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      const int c = c0 + lane;
      const bool valid = c < ncols;
      const int j = valid ? ((c < ncb) ? k + 1 + c : nh + (c - ncb)) : k;
      const unsigned tj = tcol[j].x;
      const bool act = valid && (int)(tj >> 16) <= k;
      // idle lanes read the pivot column (valid addresses, results dropped): no predicated loads in the row loops
      const int cbj = act ? (int)(tj & 0xffffu) : cbk;
      const double u = S[cbj + k] * rk;
#pragma unroll 1
      for (int q = w; q <= nchunk; q += nwl) {
        // rows past the span / past the border read (and drop) what follows in the store
        double l[8], a[8];
        const bool band = q < nchunk;
        const int r0 = q << 3;
        const unsigned lp = sS + 8u * (unsigned)(band ? lband + r0 : lbord);
        const unsigned ap = sS + 8u * (unsigned)(band ? cbj + k + 1 + r0 : bo + (act ? j : k) * bs);
        const int nrow = band ? nrb - r0 : nb, ntst = band ? nrb - r0 : nbt;
        SluRow<0, 8>::load(l, a, lp, ap);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          if (t < ntst) lmax = max(lmax, __double2hiint(l[t]) & 0x7fffffff);
          a[t] = fma(-l[t], u, a[t]);
        }
        SluRow<0, 8>::store(a, ap, act ? nrow : 0);
      }
    }
    if (lmax > thr_hi) { viol = true; atomicMin(fail + 1, k); }
  }
  slu_sync(nwl);
  // ---- border block, one warp, lane c holds column c of the block (lane nb: the right-hand side).  The diagonal is
  // the pivot while it passes the threshold test against the tested rows below it; otherwise the largest entry of the
  // column (any row) is brought up.
  if (w == 0) {
    double v[kSluBord];
    const int c = lane;
    const bool mine = c < nb || (c == nb && with_rhs);
#pragma unroll
    for (int r = 0; r < kSluBord; ++r) v[r] = (r < nb && mine) ? S[bo + (nh + c) * bs + r] : 0.0;     // c = nb: column n, the right-hand side
    int sw[kSluBord];
#pragma unroll
    for (int jj = 0; jj < kSluBord; ++jj) {
      sw[jj] = jj;
      if (jj < nb) {
        // lane jj: largest entry of column jj from the diagonal down (first maximum), and the largest among the tested rows
        double best = -1.0, bestt = -1.0; int p = jj;
#pragma unroll
        for (int r = 0; r < kSluBord; ++r) {
          if (r >= jj && r < nb) {
            const double av = fabs(v[r]);
            if (av > best) { best = av; p = r; }
            if (r < nbt && av > bestt) bestt = av;
          }
        }
        const double ad = fabs(v[jj]);
        if (ad > 0.0 && !(tau * bestt > ad)) p = jj;
        p = __shfl_sync(kFull, p, jj);
        sw[jj] = p;
        if (p != jj) {
#pragma unroll
          for (int r = 0; r < kSluBord; ++r) if (r == p) { const double tmp = v[jj]; v[jj] = v[r]; v[r] = tmp; }
        }
        const double pv = __shfl_sync(kFull, v[jj], jj);
        if (!(fabs(pv) > 0.0)) viol = true;
        const double rk = pivot_rcp<false>(pv);
        if (lane == 0) rinv[nh + jj] = rk;
#pragma unroll
        for (int r = 0; r < kSluBord; ++r) {
          if (r > jj && r < nb) {
            const double lr = __shfl_sync(kFull, v[r], jj) * rk;
            if (c > jj) v[r] = fma(-lr, v[jj], v[r]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kSluBord; ++r) {
      if (r < nb) {
        if (c < nb) S[bo + (nh + c) * bs + r] = v[r];
        else if (c == nb && with_rhs) rhs[nh + r] = v[r];
      }
    }
    // the same interchanges on the heavy-column part of the border rows (their multipliers), in order
#pragma unroll
    for (int jj = 0; jj < kSluBord; ++jj) {
      if (jj < nb) {
        if (lane == 0) ipb[jj] = sw[jj];
        if (sw[jj] != jj) {
          const int la = bo + jj, lb = bo + sw[jj];
          for (int cc = lane; cc < nh; cc += 32) { const double ta = S[la + cc * bs], tb = S[lb + cc * bs]; S[la + cc * bs] = tb; S[lb + cc * bs] = ta; }
        }
      }
    }
  }
  if (viol) *fail = 1;
}

// ---- triangular solves (all threads of the CTA enter) ------------------------------------------------------------------
// fwd_done: b already holds L^-1 P b (it rode through slu_factor).  Otherwise: forward substitution with the unscaled
// multipliers (warp 0), border interchanges first.  Then U x = y:
//   A  the border unknowns (warp 0, nb dependent steps);
//   B  their columns are full: every heavy row takes its nb updates at once (all threads);
//   C  the band, last column first (warp 0).  The 32 unknowns below the current one live in REGISTERS (lane = row mod
//      32), x(k) travels by shuffle, U's entries and the reciprocal are loaded one step ahead: the dependent chain of a
//      step is shuffle - multiply - multiply-add, no shared-memory round trip.  Rows further than 31 above the diagonal
//      (a few outlier columns) are updated in shared memory before they enter the window.
__device__ __noinline__ void slu_solve(const double *__restrict__ S, double *__restrict__ b, const double *__restrict__ rinv,
                                       const int *ipb, const bool fwd_done) {
  XB_ASSUME_SHARED(S); XB_ASSUME_SHARED(b); XB_ASSUME_SHARED(rinv); XB_ASSUME_SHARED(ipb);
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
        const double f = b[k] * rinv[k];
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
      const double x = b[k] * rinv[k];
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
  // C: the band
  if (w0 && nh > 0) {
    int k = nh - 1;
    auto row_of = [&](const int kk) { return kk - ((kk - lane) & 31); };   // the window row this lane holds at step kk (kk itself for lane kk & 31)
    auto load_u = [&](const unsigned t, const int kk) {
      const int i = row_of(kk);
      return (i != kk && i >= (int)(t >> 16)) ? S[(int)(t & 0xffffu) + i] : 0.0;     // i >= cf >= 0
    };
    double bv, be = 0.0;
    { const int i = row_of(k); bv = (i >= 0) ? b[i] : 0.0; }
    unsigned t_cur = I.col[k].x, t_next = (k >= 1) ? I.col[k - 1].x : 0u;
    double u_cur = load_u(t_cur, k), r_cur = rinv[k];
    for (; k >= 0; --k) {
      const unsigned t_n2 = (k >= 2) ? I.col[k - 2].x : 0u;
      const double u_next = (k >= 1) ? load_u(t_next, k - 1) : 0.0;
      const double r_next = (k >= 1) ? rinv[k - 1] : 0.0;
      const double x = __shfl_sync(kFull, bv, k & 31) * r_cur;
      if (k + 1 < nh && lane == ((k + 1) & 31)) bv = be;      // row k-31 enters the window (loaded a step ago)
      const int cfk = (int)(t_cur >> 16), cbk = (int)(t_cur & 0xffffu);
      if (cfk <= k - 32) {                                     // outlier column: rows above the window, in shared memory
        for (int i = cfk + lane; i <= k - 32; i += 32) b[i] = fma(-S[cbk + i], x, b[i]);
        __syncwarp();
      }
      bv = fma(-u_cur, x, bv);
      if (lane == (k & 31)) { b[k] = x; be = (k >= 32) ? b[k - 32] : 0.0; }
      t_cur = t_next; t_next = t_n2; u_cur = u_next; r_cur = r_next;
    }
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
