// Structured LU of the backward-Euler matrix: the B200 answer to the reference's sparse plug-ins
// (XNet tree: source/xnet_jacobian_MA48.F90, pattern from sparse_check, source/xnet_preprocess.F90:291-503).
//
// Equations are reordered with the light particles (n, p, d, t, he3, he4) and, with self-heating, T9 LAST.  In that
// order the matrix is a variable band plus a dense border, and an elimination without row interchanges never leaves
// its skyline envelope (host tables: structured_lu_tables, xb_loader.cpp).  Storage in shared memory, S[nstore]:
//     element (i,j), i <= j : S[uo[j] + i]     (column j of U, rows cf[j]..j, contiguous)
//     element (i,j), i >  j : S[lo[i] + j]     (row i of L, columns rf[i]..i-1, contiguous; kept UNSCALED: the
//                                               multiplier is S * rinv[j], rinv = reciprocals of U's diagonal)
// SN160: 5.3k doubles (42 KB) instead of 25.6k, 9e4 flops instead of 2.7e6 per factorisation -- four zones fit on
// one SM where the dense tile allows one.
//
// Numerical safety: every heavy column is verified, tau*|a(i,k)| <= |a(k,k)| for all rows below (the test the dense
// fast path applies to its predicted pivot; tau = 1e-3, see kSluTauDefault below); the border block at the end
// is factored with ordinary partial pivoting (its rows are full, so interchanges stay inside the envelope).  A
// factorisation with a violated column is discarded and redone by a dense partial-pivoting LU in the CTA's global
// scratch (slu_dense_*), so the result is always that of a bounded-growth LU.
#pragma once
#include "xb_device.cuh"
#include "xb_lu.cuh"

namespace xb {

// Default threshold of the column test (DevSlu.tau; XB_SLU_TAU in the environment overrides it at context creation).
// 1e-3 bounds the multipliers by 1000.  Measured on every step of the SN160 flame trajectory
// (tools/structured_lu_stability.py, profiles/r02_structured_lu_stability.txt): 2 % of the factorisations have a
// multiplier above 10 (always the proton row against ga64 / ge63, whose diagonal has cancelled by the time it is the
// pivot), the largest is 222, and the solutions of those systems are as close to an extended-precision solve as
// LAPACK's partial-pivoting ones (error ratio 0.7 - 5.5): the growth that a large multiplier could cause does not occur.
constexpr double kSluTauDefault = 1.0e-3;
constexpr int kSluBord = 8;     // border rows/columns handled in registers (light particles + T9 <= 7)

// per-zone views of the index tables (shared memory, 16-bit)
struct SluIdx {
  const unsigned short *rf, *cf, *lo, *uo, *rhi, *chi;
  int n, nh, nb;   // order, heavy rows (n - nb), border
};
__device__ __forceinline__ SluIdx slu_idx() {
  const DevSlu &D = c_net.slu;
  SluIdx I;
  const unsigned short *t = reinterpret_cast<const unsigned short *>(xb_smem + D.off_tab);
  I.n = D.n; I.nb = D.nbord; I.nh = D.n - D.nbord;
  I.rf = t; I.cf = t + D.n; I.lo = t + 2 * D.n; I.uo = t + 3 * D.n; I.rhi = t + 4 * D.n; I.chi = t + 5 * D.n;
  return I;
}
// copies the tables of the context into the CTA's shared memory (kernel prologue)
__device__ __forceinline__ void slu_stage_tables() {
  const DevSlu &D = c_net.slu;
  unsigned short *t = reinterpret_cast<unsigned short *>(xb_smem + D.off_tab);
  for (int k = threadIdx.x; k < 7 * D.n + 2; k += blockDim.x) t[k] = D.tab16[k];
}
__device__ __forceinline__ int slu_addr(const SluIdx &I, const int i, const int j) {
  return (i <= j) ? (int)I.uo[j] + i : (int)I.lo[i] + j;
}

// candidate number r of the rows below pivot k: the band rows k+1..rhi[k], then the border rows nh..n-1 (those above
// k).  Returns the row, or -1 when r is past the end or the row has no entry in column k.
__device__ __forceinline__ int slu_row_cand(const SluIdx &I, const int k, const int nband, const int r, const int ncand) {
  if (r >= ncand) return -1;
  const int i = (r < nband) ? k + 1 + r : I.nh + (r - nband);
  return ((int)I.rf[i] <= k) ? i : -1;
}
__device__ __forceinline__ int slu_col_cand(const SluIdx &I, const int k, const int nband, const int c, const int ncand) {
  if (c >= ncand) return -1;
  const int j = (c < nband) ? k + 1 + c : I.nh + (c - nband);
  return ((int)I.cf[j] <= k) ? j : -1;
}

__device__ __forceinline__ void slu_sync(const int nwl) {
  if (nwl == 1) __syncwarp(); else bar_named(1, nwl * 32);
}

// ---- factorisation -----------------------------------------------------------------------------------------------
// The elimination of the heavy part is a FIXED sequence of element updates (the structure is the network's, the same
// for every zone and every step), so it is executed from a precomputed schedule instead of being re-derived from the
// index tables by every lane of every warp: entry = (a, b, c) meaning S[a] -= (S[b] * rinv_k) * S[c], with b in the
// pivot column, c in the pivot row.  Entries of one pivot are independent; they are stored in blocks of ntl = 32*nwl
// (one entry per participating thread, coalesced 8-byte loads, padded with no-ops), the first block of a pivot is
// flagged: barrier, new pivot.  Every lane does useful work (the row x column tables above left a third of the lanes
// idle and cost 22 instructions per column and warp).  All blocks of a pivot are processed together -- every load
// issued before the first FMA, every store after the last -- so the latency chain per pivot is barrier, one round of
// shared-memory loads, reciprocal, multiply-add, store.  The schedule streams from L2 (SN160: 0.5 MB per
// factorisation, shared by all zones); the entries of the next two pivots are in flight in registers.
//   x = a | b << 16,  y = c | flags << 16;  flags: 1 = row exempt from the threshold test (T9 row), 4 = right-hand-side
//   entry (skipped when the right-hand side is not carried);  a = 0xffff: no-op.  pbk[k] (16-bit table, shared memory):
//   first block of pivot k.
// Executed by the first nwl warps of the CTA (the others return at once; the caller's __syncthreads joins them).
// with_rhs: the right-hand side (rhs = S + rhs_off, position order) is carried along as an extra column: it leaves as
// L^-1 P b.  fail[0] is set when a heavy column violates the threshold test or a pivot is zero/NaN, fail[1] receives
// the first such column (caller initialises it to a large value).  ipb[kSluBord]: interchanges of the border block.
constexpr int kSluChunk = 6;     // blocks of one pivot processed together (all loads, then all FMAs, then all stores)
__device__ __noinline__ void slu_factor(double *__restrict__ S, double *__restrict__ rhs, double *__restrict__ rinv,
                                        int *ipb, int *fail, const bool with_rhs, const int nwl) {
  XB_ASSUME_SHARED(S); XB_ASSUME_SHARED(rhs); XB_ASSUME_SHARED(rinv); XB_ASSUME_SHARED(fail); XB_ASSUME_SHARED(ipb);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (w >= nwl) return;
  const SluIdx I = slu_idx();
  const int nh = I.nh, nb = I.nb;
  // The temperature equation (self-heating: the last row) is exempt from the test: it is the linear functional
  // -sum_i mex_i (row i) / cv of the abundance rows, i.e. a row in different units whose multipliers measure mex/cv and
  // not element growth (elimination without interchanges is invariant under row scaling); it is pivoted in the border.
  const double tau = c_net.slu.tau;
  bool viol = false;
  {
    const DevSlu &D = c_net.slu;
    const int ntl = nwl * 32;
    const uint2 *__restrict__ sch = D.sched + tid;
    const unsigned short *pbk = I.rf + 6 * I.n;          // [nh+1] first schedule block of every pivot
    const uint2 kNop = make_uint2(0xffffu, 0u);
    // entries of the next two pivots are in flight (registers) while the current one is eliminated: the schedule comes
    // from L2 (~250 cycles), a pivot takes ~150
    uint2 q1[kSluChunk], q2[kSluChunk];
    auto fetch = [&](uint2 (&q)[kSluChunk], const int kk) {
      if (kk < nh) {
        const int g0 = (int)pbk[kk], cnt = (int)pbk[kk + 1] - g0;
#pragma unroll
        for (int m = 0; m < kSluChunk; ++m) q[m] = (m < cnt) ? __ldg(sch + (size_t)(g0 + m) * ntl) : kNop;
      }
    };
    fetch(q1, 0); fetch(q2, 1);
    for (int k = 0; k < nh; ++k) {
      uint2 e[kSluChunk];
#pragma unroll
      for (int m = 0; m < kSluChunk; ++m) { e[m] = q1[m]; q1[m] = q2[m]; }
      fetch(q2, k + 2);
      const int g0 = (int)pbk[k], cnt = (int)pbk[k + 1] - g0;
      if (k > 0) slu_sync(nwl);
      const double piv = S[(int)I.uo[k] + k];
      const double apiv = fabs(piv);
      const double rk = pivot_rcp<false>(piv);
      if (!(apiv > 0.0)) { viol = true; atomicMin(fail + 1, k); }   // zero or NaN pivot
      if (tid == 0) rinv[k] = rk;
      for (int base = 0; base < cnt; base += kSluChunk) {
        if (base > 0) {                                  // wide pivots (the dense head of the band): fetched on demand
#pragma unroll
          for (int m = 0; m < kSluChunk; ++m) e[m] = (base + m < cnt) ? __ldg(sch + (size_t)(g0 + base + m) * ntl) : kNop;
        }
        double sb[kSluChunk], sc[kSluChunk], sa[kSluChunk];
        bool on[kSluChunk];
#pragma unroll
        for (int m = 0; m < kSluChunk; ++m) {
          const unsigned a = e[m].x & 0xffffu, fl = e[m].y >> 16;
          on[m] = a != 0xffffu && (with_rhs || !(fl & 4u));
          sb[m] = on[m] ? S[e[m].x >> 16] : 0.0;
          sc[m] = on[m] ? S[e[m].y & 0xffffu] : 0.0;
          sa[m] = on[m] ? S[a] : 0.0;
        }
#pragma unroll
        for (int m = 0; m < kSluChunk; ++m) {
          if (!((e[m].y >> 16) & 1u) && tau * fabs(sb[m]) > apiv) { viol = true; atomicMin(fail + 1, k); }   // fail[1]: first violated column
          sa[m] = fma(-(sb[m] * rk), sc[m], sa[m]);
        }
#pragma unroll
        for (int m = 0; m < kSluChunk; ++m) if (on[m]) S[e[m].x & 0xffffu] = sa[m];
      }
    }
  }
  slu_sync(nwl);
  // ---- border block: partial pivoting, one warp, lane c holds column c of the block (lane nb: the right-hand side)
  if (w == 0) {
    double v[kSluBord];
    const int c = lane;
#pragma unroll
    for (int r = 0; r < kSluBord; ++r) {
      v[r] = 0.0;
      if (r < nb) {
        if (c < nb) v[r] = S[slu_addr(I, nh + r, nh + c)];
        else if (c == nb && with_rhs) v[r] = rhs[nh + r];
      }
    }
    int sw[kSluBord];
#pragma unroll
    for (int jj = 0; jj < kSluBord; ++jj) {
      sw[jj] = jj;
      if (jj < nb) {
        // IDAMAX over rows jj.. of column jj (lane jj), first maximum
        double best = -1.0; int p = jj;
#pragma unroll
        for (int r = 0; r < kSluBord; ++r) {
          if (r >= jj && r < nb) { const double av = fabs(v[r]); if (av > best) { best = av; p = r; } }
        }
        p = __shfl_sync(kFull, p, jj);
        sw[jj] = p;
        if (p != jj) {
#pragma unroll
          for (int r = 0; r < kSluBord; ++r) if (r == p) { const double tmp = v[jj]; v[jj] = v[r]; v[r] = tmp; }
        }
        const double piv = __shfl_sync(kFull, v[jj], jj);
        if (!(fabs(piv) > 0.0)) viol = true;
        const double rk = pivot_rcp<false>(piv);
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
        if (c < nb) S[slu_addr(I, nh + r, nh + c)] = v[r];
        else if (c == nb && with_rhs) rhs[nh + r] = v[r];
      }
    }
    // the same interchanges on the heavy-column part of the border rows (their multipliers), in order
#pragma unroll
    for (int jj = 0; jj < kSluBord; ++jj) {
      if (jj < nb) {
        if (lane == 0) ipb[jj] = sw[jj];
        if (sw[jj] != jj) {
          const int la = (int)I.lo[nh + jj], lb = (int)I.lo[nh + sw[jj]];
          for (int cc = lane; cc < nh; cc += 32) { const double ta = S[la + cc], tb = S[lb + cc]; S[la + cc] = tb; S[lb + cc] = ta; }
        }
      }
    }
  }
  if (viol) *fail = 1;
}

// ---- triangular solves (one warp) -----------------------------------------------------------------------------------
// fwd_done: b already holds L^-1 P b (it rode through slu_factor).  Otherwise: forward substitution with the unscaled
// multipliers, the border interchanges applied when the border is reached.  Then U x = y from the last column up.
__device__ __noinline__ void slu_solve(const double *__restrict__ S, double *__restrict__ b, const double *__restrict__ rinv,
                                       const int *ipb, const bool fwd_done) {
  XB_ASSUME_SHARED(S); XB_ASSUME_SHARED(b); XB_ASSUME_SHARED(rinv); XB_ASSUME_SHARED(ipb);
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  const SluIdx I = slu_idx();
  const int n = I.n, nh = I.nh, nb = I.nb;
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
        const int nrb = (int)I.rhi[k] - k, nrc = nrb + nb;
        for (int rb = 0; rb < nrc; rb += 32) {
          const int i = slu_row_cand(I, k, nrb, rb + lane, nrc);
          if (i >= 0) b[i] = fma(-S[(int)I.lo[i] + k], f, b[i]);
        }
      } else {
        const int i = k + 1 + lane;
        if (i < n) b[i] = fma(-S[(int)I.lo[i] + k], f, b[i]);
      }
      __syncwarp();
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    const double x = b[k] * rinv[k];
    const int c0 = (int)I.cf[k], uok = (int)I.uo[k];
    __syncwarp();
    if (lane == 0) b[k] = x;
    for (int i = c0 + lane; i < k; i += 32) b[i] = fma(-S[uok + i], x, b[i]);
    __syncwarp();
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
    const int c0 = (int)I.cf[p], r0 = (int)I.rf[p];
    for (int i = c0 + lane; i <= p; i += 32) A[(size_t)p * n + i] = S[(int)I.uo[p] + i];
    for (int j = r0 + lane; j < p; j += 32) A[(size_t)j * n + p] = S[(int)I.lo[p] + j];
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
