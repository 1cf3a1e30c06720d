// Structured-LU variant of the fused evolve kernel (fast arithmetic): its own translation unit, so that its register
// budget (<= 128 per thread: several zones resident per SM) is not dictated by the register-panel dense LU of
// xb_api.cu.  Device code = the same zone routines (xb_zone.cuh) compiled with XB_SLU: skyline Jacobian assembly,
// slu_factor / slu_solve (xb_slu.cuh) in place of lu_blocked / bksub_block, dense partial-pivoting fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/xnet_b200.h"
#include "xb_host.h"
#define XB_SLU 1
#ifndef XB_MAX_THREADS
#define XB_MAX_THREADS 256
#endif
#include "xb_zone.cuh"
#include "xb_ctx.h"

#define CUDA_TRY_S CUDA_TRY

namespace {

__device__ void slu_zone_prologue(const DevScratch &scr) {
  ZoneScalars *s = reinterpret_cast<ZoneScalars *>(xb_smem + c_net.lay.scal);
  if (threadIdx.x == 0) {
    double *g = scr.base + (size_t)blockIdx.x * scr.stride;
    s->g_csect = g + scr.off_csect;
    s->g_rffn = g + scr.off_rffn;
    s->g_qkffn = g + scr.off_qkffn;
    s->g_prod = g + scr.off_prod;
    s->g_jac = g + scr.off_jac;          // dense fallback: (n+1) x n doubles + n ints
    s->rhs_in_jac = 0;
    s->g_yo = g + scr.off_yo;
    s->g_dcsect = g + scr.off_dcsect;
    s->g_dlnrffn = g + scr.off_dlnrffn;
    s->g_bdf = nullptr;                  // the structured kernel integrates with backward Euler only
    reinterpret_cast<double *>(xb_smem + c_net.lay.yt)[c_net.ny] = 1.0;
  }
  slu_stage_tables();
  __syncthreads();
}

// Persistent kernel: every CTA pulls zones from a global counter (as evolve_kernel, xb_api.cu).  256 threads x 2 CTAs
// and 128 threads x 4 CTAs both fit the 128-register budget this bound implies.
__global__ void __launch_bounds__(XB_MAX_THREADS, 2)
evolve_slu_kernel(const __grid_constant__ DevZones zn, const __grid_constant__ DevOut out, const __grid_constant__ DevScratch scr,
                  int *counter) {
  __shared__ int s_zone;
  slu_zone_prologue(scr);
  ZoneScalars *zs = reinterpret_cast<ZoneScalars *>(xb_smem + c_net.lay.scal);
  for (;;) {
    if (threadIdx.x == 0) s_zone = atomicAdd(counter, 1);
    __syncthreads();
    const int iz = s_zone;
    __syncthreads();
    if (iz >= zn.nzone) break;
    full_net_cta<false>(zn, out, iz);
    if (threadIdx.x == 0) {
      for (int k = 0; k < PH_COUNT; ++k)
        if (k != PH_ZONES) atomicAdd(&scr.phase[k], (unsigned long long)zs->ph[k]);
      atomicAdd(&scr.phase[PH_ZONES], 1ULL);
    }
    __syncthreads();
  }
}

// Test kernel: one CTA per system.  a[nz][n*n] column-major in EQUATION order, b[nz][n]; the matrix is scattered into
// the skyline store (flags bit 1: an entry outside the envelope was non-zero), factored and solved.
//   mode 0: right-hand side carried through the factorisation;  mode 1: factor, then forward + backward solve;
//   mode 2: dense fallback (carried);  mode 3: dense fallback, separate solve
// flags bit 0: threshold test failed (the result is then that of the failed structured factorisation unless mode >= 2).
__global__ void __launch_bounds__(XB_MAX_THREADS, 2)
slu_test_kernel(const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ x, int *__restrict__ flags,
                const __grid_constant__ DevScratch scr, const int mode) {
  slu_zone_prologue(scr);
  ZoneMem zm = zone_mem();
  ZoneScalars *s = zm.s;
  const DevSlu &D = c_net.slu;
  const SluIdx I = slu_idx();
  const int n = I.n, tid = threadIdx.x, nt = blockDim.x, iz = blockIdx.x;
  const double *A = a + (size_t)iz * n * n;
  double *S = zm.jac;
  const int *__restrict__ spos = D.pos;
  if (tid == 0) { s->slu_fail = 0; s->slu_fail_k = 0x7fffffff; s->info = 0; s->flag = 0; }
  for (int e = tid; e < D.nstore; e += nt) S[e] = 0.0;
  __syncthreads();
  for (int e = tid; e < n * n; e += nt) {
    const int col = e / n, row = e - col * n;
    const int pi = spos[row], pj = spos[col];
    const bool inside = pi >= I.nh || (slu_cf(I, pj) <= pi && pi <= slu_cbot(I, pj));
    const double v = A[e];
    if (inside) S[slu_addr(I, pi, pj)] = v;
    else if (v != 0.0) s->flag = 2;
  }
  for (int k = tid; k < n; k += nt) zm.rhs[spos[k]] = b[(size_t)iz * n + k];
  __syncthreads();
  int *ipiv = reinterpret_cast<int *>(s->g_jac + (size_t)(n + 1) * n);
  if (mode < 2) {
    if (mode == 0) {
      slu_factor(S);
      __syncthreads();
      slu_solve(S, zm.rhs, s->slu_ipb, true);
    } else {
      // the factorisation always carries the right-hand side: keep a copy for the separate forward + backward solve
      double *keep = s->g_jac;
      for (int k = tid; k < n; k += nt) keep[k] = zm.rhs[k];
      __syncthreads();
      slu_factor(S);
      __syncthreads();
      for (int k = tid; k < n; k += nt) zm.rhs[k] = keep[k];
      __syncthreads();
      slu_solve(S, zm.rhs, s->slu_ipb, false);
    }
  } else {
    slu_dense_factor(S, zm.rhs, s->g_jac, ipiv, zm.part, &s->info);
    __syncthreads();
    slu_dense_solve(s->g_jac, ipiv, zm.rhs, mode == 2);
  }
  __syncthreads();
  for (int k = tid; k < n; k += nt) x[(size_t)iz * n + k] = zm.rhs[spos[k]];
  if (tid == 0) flags[iz] = (s->slu_fail ? 1 : 0) | s->flag | (s->info ? 4 : 0);
}

std::mutex g_slu_mu;

int env_int(const char *name, int dflt) {
  const char *v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

}  // namespace

// Shared-memory carve-up of one zone of the structured kernel
static SmemOff slu_layout(int ny, int n, int nstore, int npart, int nzs, int *off_tab, unsigned rhs_avoid) {
  SmemOff L{};
  size_t o = 0;
  auto al = [&]() { o = (o + 15) / 16 * 16; };
  L.scal = (int)o; o += sizeof(ZoneScalars); al();
  L.jac = (int)o; o += (size_t)nstore * 8; al();
  // the right-hand side is one more column of the elimination: its base (in doubles from S) avoids the bank pairs of the
  // border columns it shares a half-warp with
  while ((rhs_avoid >> (((o - L.jac) / 8) % 16)) & 1u) o += 8;
  L.rhs = (int)o; o += (size_t)n * 8;
  L.y = (int)o; o += (size_t)ny * 8;
  L.yt = (int)o; o += (size_t)(ny + 1) * 8;
  L.ydot = (int)o; o += (size_t)ny * 8;
  al();
  L.part = (int)o; o += (size_t)npart * 8; al();
  L.fhs = (int)o; o += (size_t)2 * nzs * 8; al();
  *off_tab = (int)o; o += ((size_t)n + 1) * 8; al();   // column table [n+1]
  L.rdiag = L.prow = L.indx = L.orig = L.pos = L.scal;    // not used by the structured path (the reciprocals of the pivots replace U's diagonal)
  L.lds2 = 0; L.lu_s = L.lu_u = (int)o;
  // slu_factor's row loops read (and drop) up to 8 rows past a column / past the border block: keep that inside the CTA's memory
  o = std::max(o, (size_t)L.jac + ((size_t)nstore + 16) * 8); al();
  L.cs = (int)o;
  L.total = (int)o;
  return L;
}

int xb_slu_configure(xb_ctx *c) {
  const Network &h = c->net->h;
  const int heat = c->ctl.iheat > 0 ? 1 : 0;
  SluTables &T = c->slu_tab;
  structured_lu_tables(h, heat, T);
  const int n = T.n, ny = h.ny;
  if (n != c->dn.msize) return fail(XB_ERR_ARG, "structured LU: order mismatch");
  if (T.nbord > kSluBord - 1 || T.nstore >= 65536 || n < T.nbord + 1)
    return fail(XB_ERR_UNSUPPORTED, "structured LU: network outside the limits of the kernel (border " + std::to_string(T.nbord) +
                                        ", store " + std::to_string(T.nstore) + " doubles)");
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, c->device));
  const size_t max_smem = prop.sharedMemPerBlockOptin, sm_smem = prop.sharedMemPerMultiprocessor;
  const int nh = n - T.nbord;
  for (int k = 0; k < nh; ++k)
    if (T.nrb[k] > 255 || T.ncb[k] > 255) return fail(XB_ERR_UNSUPPORTED, "structured LU: band wider than 255");
  CUDA_TRY(c->slu_pos.put(T.pos));
  unsigned rhs_avoid = 0;
  for (int j = nh; j < n; ++j) rhs_avoid |= 1u << (((T.cb[j] % 16) + 16) % 16);
  // threads per zone: 128 when three or more zones fit on an SM, else 256 (XB_SLU_THREADS overrides)
  DevNet &d = c->dn_slu;
  d = c->dn;
  int threads = 128;
  for (int pass = 0; pass < 2; ++pass) {
    int rc = build_plans(c, d, c->pby_slu, c->pbj_slu, threads, n, &c->slu_npart, &T); if (rc) return rc;
    int off_tab = 0;
    SmemOff L = slu_layout(ny, n, T.nstore, c->slu_npart, d.nzs, &off_tab, rhs_avoid);
    if ((size_t)L.total + 64 > max_smem) return fail(XB_ERR_UNSUPPORTED, "structured LU: the zone does not fit in shared memory");
    const int per_sm_smem = (int)(sm_smem / ((size_t)L.total + 1024 + 64));
    const int want = env_int("XB_SLU_THREADS", 0);
    const int t2 = want > 0 ? want : (per_sm_smem >= 3 ? 128 : 256);
    d.lay = L; d.slu.off_tab = off_tab;
    if (t2 == threads) break;
    if (t2 != 64 && t2 != 128 && t2 != 256) return fail(XB_ERR_ARG, "XB_SLU_THREADS must be 64, 128 or 256");
    threads = t2;
    if (pass == 1) break;
  }
  {
    int off_tab = 0;
    int rc = build_plans(c, d, c->pby_slu, c->pbj_slu, threads, n, &c->slu_npart, &T); if (rc) return rc;
    SmemOff L = slu_layout(ny, n, T.nstore, c->slu_npart, d.nzs, &off_tab, rhs_avoid);
    d.lay = L; d.slu.off_tab = off_tab;
  }
  c->slu_threads = threads;
  d.tile = 1; d.ld = n;
  d.slu.ntest = n - heat;
  { const char *v = std::getenv("XB_SLU_TAU"); d.slu.tau = (v && *v) ? std::atof(v) : kSluTauDefault; }
  d.slu.on = 1; d.slu.n = n; d.slu.nstore = T.nstore; d.slu.nbord = T.nbord;
  d.slu.nwl = std::max(1, std::min(env_int("XB_SLU_NWL", threads / 32), threads / 32));
  d.slu.force_fb = env_int("XB_SLU_FORCE_FALLBACK", 0);
  d.slu.pos = c->slu_pos.p;
  d.slu.itau = 1.0 / d.slu.tau;
  d.slu.bo = T.bo; d.slu.bs = T.bs;
  {
    // staged tables: per column cb | cf << 16 (entry n = the right-hand side as a pseudo-column), per pivot ncb | nrb << 8
    const int rhs_off = (d.lay.rhs - d.lay.jac) / 8;
    if (rhs_off < T.nstore || rhs_off + n >= 0xffff) return fail(XB_ERR_UNSUPPORTED, "structured LU: store too large for 16-bit addresses");
    d.slu.rhs_off = rhs_off;
    d.slu.tab_rel = (d.slu.off_tab - d.lay.jac) / 8;
    d.slu.ph_rel = (d.lay.scal + (int)offsetof(ZoneScalars, ph) - d.lay.jac) / 8;
    d.slu.fail_rel = (d.lay.scal + (int)offsetof(ZoneScalars, slu_fail) - d.lay.jac) / 4;
    std::vector<uint2> tab((size_t)n + 1, make_uint2(0u, 0u));
    for (int j = 0; j < n; ++j) {
      if (T.cb[j] < 0) return fail(XB_ERR_UNSUPPORTED, "structured LU: negative column base");
      tab[j].x = (unsigned)T.cb[j] | ((unsigned)T.cf[j] << 16);
      if (j < nh) tab[j].y = (unsigned)T.ncb[j] | ((unsigned)T.nrb[j] << 8);
    }
    tab[n].x = (unsigned)rhs_off;
    CUDA_TRY(c->slu_tabd.put(tab));
    d.slu.tab = c->slu_tabd.p;
  }
  {
    // pivot schedule: per round, warp and lane the addresses slu_factor works on (read from L2, three rounds ahead);
    // pair-aligned tables: two pivots per round
    d.slu.pair = T.pair;
    if (T.pair) {
      SluSchedule2 sc;
      structured_lu_schedule2(T, d.slu.nwl, d.slu.rhs_off, d.slu.ntest, sc);
      CUDA_TRY(c->slu_schw.put(sc.warp)); CUDA_TRY(c->slu_schl.put(sc.lane));
      d.slu.nround = sc.nround;
      d.slu.sch2_warp = reinterpret_cast<const uint4 *>(c->slu_schw.p);
      d.slu.sch2_lane = reinterpret_cast<const uint2 *>(c->slu_schl.p);
      d.slu.sch_warp = nullptr; d.slu.sch_lane = nullptr;
    } else {
      SluSchedule sc;
      structured_lu_schedule(T, d.slu.nwl, d.slu.rhs_off, d.slu.ntest, sc);
      CUDA_TRY(c->slu_schw.put(sc.warp)); CUDA_TRY(c->slu_schl.put(sc.lane));
      d.slu.nround = sc.nround;
      d.slu.sch_warp = reinterpret_cast<const uint2 *>(c->slu_schw.p);
      d.slu.sch_lane = c->slu_schl.p;
      d.slu.sch2_warp = nullptr; d.slu.sch2_lane = nullptr;
    }
  }
  {
    // back-substitution over the band (slu_solve, phase C): per step k = nh-1 .. 0 and lane the address of U(i,k) for the
    // window row i = k - ((k - lane) & 31), of the pivot's reciprocal for i = k, 0xffff where the column does not reach
    std::vector<uint32_t> sl; std::vector<int> so; int ngroup = 0;
    structured_lu_solve_table(T, sl, so, &ngroup);
    CUDA_TRY(c->slu_soll.put(sl)); CUDA_TRY(c->slu_solo.put(so));
    d.slu.sol_ngroup = ngroup;
    d.slu.sol_lane = reinterpret_cast<const uint2 *>(c->slu_soll.p);
    d.slu.sol_out = c->slu_solo.p;
  }
  // zones per SM by shared memory, then whatever is left on the SM caches csect (tail of the reaction list first)
  SmemOff &L = d.lay;
  int per_sm = std::max(1, (int)(sm_smem / ((size_t)L.total + 1024 + 64)));
  per_sm = std::min(per_sm, std::max(1, 512 / threads));            // 128 registers per thread: 512 threads per SM
  per_sm = std::min(per_sm, std::max(1, env_int("XB_SLU_PER_SM", per_sm)));
  {
    const size_t budget = std::min(max_smem - 64, sm_smem / per_sm - 1024 - 64);   // 1 KB per CTA reserved by the system, 16 B static
    const size_t room = budget > (size_t)L.total ? (budget - (size_t)L.total) / 8 : 0;
    const int ncache = (int)std::min<size_t>(room, (size_t)d.nrtot);
    d.cs_lo = d.nrtot - ncache;
    L.total = (int)(((size_t)L.total + (size_t)ncache * 8 + 15) / 16 * 16);
  }
  c->slu_smem = (size_t)L.total;
  if (env_int("XB_SLU_VERBOSE", 0))
    std::fprintf(stderr, "structured LU: n %d, store %d doubles, zone without rate cache %zu B (scalars %zu, part %d), rates cached %d of %d, "
                         "%d zones per SM by shared memory, %d threads, schedule %d rounds\n",
                 n, T.nstore, (size_t)L.total - (size_t)(d.nrtot - d.cs_lo) * 8, sizeof(ZoneScalars), c->slu_npart, d.nrtot - d.cs_lo, d.nrtot,
                 per_sm, threads, d.slu.nround);
  CUDA_TRY(cudaFuncSetAttribute(evolve_slu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->slu_smem));
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, evolve_slu_kernel, threads, c->slu_smem));
  if (occ < 1) return fail(XB_ERR_CUDA, "structured evolve kernel does not fit on an SM");
  c->slu_per_sm = std::min(occ, per_sm);
  c->slu_grid = c->nsm * c->slu_per_sm;
  // per-CTA scratch: rates, FFN tables, previous abundances, the dense fallback's matrix (+ rhs column, + pivots)
  DevScratch &s = c->scr_slu;
  size_t o = 0;
  s.off_csect = o; o += d.nrtot;
  s.off_rffn = o; o += std::max(d.nffn, 1);
  s.off_qkffn = o; o += d.nffn + 1;
  s.off_dcsect = o; if (heat) o += d.nrtot;
  s.off_dlnrffn = o; if (heat) o += std::max(d.nffn, 1);
  o = (o + 15) / 16 * 16;
  s.off_prod = o;
  s.off_yo = o; o += (size_t)ny;
  o = (o + 15) / 16 * 16;
  s.off_jac = o; o += (size_t)(n + 1) * n + (size_t)(n + 1) / 2 + 1;
  s.stride = (o + 15) / 16 * 16;
  const size_t nblocks = std::max<size_t>(c->slu_grid, c->nzevolve);
  CUDA_TRY(c->scratch_slu.resize(nblocks * s.stride));
  CUDA_TRY(cudaMemset(c->scratch_slu.p, 0, nblocks * s.stride * sizeof(double)));
  s.base = c->scratch_slu.p; s.phase = c->phase.p;
  c->slu_on = 1;
  return XB_OK;
}

void xb_slu_release(xb_ctx *c) {
  c->slu_tabd.release(); c->slu_pos.release(); c->slu_schw.release(); c->slu_schl.release(); c->slu_soll.release(); c->slu_solo.release(); c->scratch_slu.release();
  xb_ctx::PlanBufs *pb[] = {&c->pby_slu, &c->pbj_slu};
  for (auto *b : pb) {
    b->a.release(); b->nn.release(); b->mc.release(); b->seg_start.release(); b->seg_dest.release();
    b->multi_start.release(); b->multi_dest.release();
  }
}

static int slu_bind(xb_ctx *c) {
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_net, &c->dn_slu, sizeof(DevNet), 0, cudaMemcpyHostToDevice, 0));
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_ctl, &c->dc, sizeof(DevCtl), 0, cudaMemcpyHostToDevice, 0));
  return XB_OK;
}

int xb_slu_evolve(xb_ctx *c, float *kernel_ms) {
  CUDA_TRY(cudaMemsetAsync(c->counter.p, 0, sizeof(int)));
  CUDA_TRY(cudaMemsetAsync(c->phase.p, 0, PH_COUNT * sizeof(unsigned long long)));
  const int grid = std::min(c->slu_grid, c->nzone);
  std::unique_lock<std::mutex> lock(g_slu_mu);
  { int rcb = slu_bind(c); if (rcb) return rcb; }
  CUDA_TRY(cudaFuncSetAttribute(evolve_slu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->slu_smem));
  CUDA_TRY(cudaEventRecord(c->ev0));
  evolve_slu_kernel<<<grid, c->slu_threads, c->slu_smem>>>(c->dz, c->dout, c->scr_slu, c->counter.p);
  xb_count_launch();
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(c->ev1));
  lock.unlock();
  CUDA_TRY(cudaEventSynchronize(c->ev1));
  if (kernel_ms) CUDA_TRY(cudaEventElapsedTime(kernel_ms, c->ev0, c->ev1));
  return XB_OK;
}

// host arrays in / out; see slu_test_kernel
int xb_slu_solve_dense_test(xb_ctx *c, int nz, const double *a, const double *b, double *x, int32_t *flags, int mode) {
  if (!c->slu_on) return fail(XB_ERR_ARG, "the context has no structured solver (ctl.solver = 1, arith = 0)");
  if (nz < 1 || nz > c->nzevolve) return fail(XB_ERR_ARG, "xb_slu_solve: nz outside 1..nzevolve");
  const int n = c->dn_slu.slu.n;
  double *da = nullptr, *db = nullptr, *dx = nullptr; int *df = nullptr;
  CUDA_TRY(cudaMalloc(&da, (size_t)nz * n * n * 8)); CUDA_TRY(cudaMalloc(&db, (size_t)nz * n * 8));
  CUDA_TRY(cudaMalloc(&dx, (size_t)nz * n * 8)); CUDA_TRY(cudaMalloc(&df, (size_t)nz * 4));
  CUDA_TRY(cudaMemcpy(da, a, (size_t)nz * n * n * 8, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(db, b, (size_t)nz * n * 8, cudaMemcpyHostToDevice));
  {
    std::lock_guard<std::mutex> lock(g_slu_mu);
    { int rcb = slu_bind(c); if (rcb) return rcb; }
    CUDA_TRY(cudaFuncSetAttribute(slu_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->slu_smem));
    CUDA_TRY(cudaEventRecord(c->ev0));
    slu_test_kernel<<<nz, c->slu_threads, c->slu_smem>>>(da, db, dx, df, c->scr_slu, mode);
    xb_count_launch();
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev1));
  }
  CUDA_TRY(cudaMemcpy(x, dx, (size_t)nz * n * 8, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(flags, df, (size_t)nz * 4, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(db); cudaFree(dx); cudaFree(df);
  return XB_OK;
}
