#!/usr/bin/env python
"""bench.py -- zone-Newton-iterations/s of the batched per-zone implicit step (BASELINE.json metric).

Workload (default): BASELINE.json configs[2] -- SN160, backward Euler (the integrator the reference builds), screening
on, th_sn1aflame + Data_SN160/ab_co replicated to 16384 zones per GPU.

A "step" is one launch of the fused evolve kernel over this rank's zones, each zone advancing a fixed WINDOW of the
trajectory: zone z restarts from accepted step idx[(z div K) mod M] of fixture z mod K (M restart points spread evenly
over the complete trajectory, taken from a complete run of one zone per fixture made during set-up) and takes up to S
accepted steps.  With M*S >= the length of the trajectory the M windows tile it, so one step does the complete
trajectory's mix of work (cold start, flame passage, freeze-out) in a bounded time; the complete-trajectory figure of
the same zones is measured once and reported as `also.complete_trajectory`.  Zones are independent: N GPUs evolve N
blocks of zones (weak scaling, no data-path collective); value = sum over all zones of the reference's own counter
ktot(2) / max-over-ranks device time.

  value : inputs resident in HBM, CUDA-event time of the evolve kernel (xb_evolve_resident)
  e2e   : the public C-ABI call xb_evolve with pinned HOST buffers (H2D of the zone state + kernel + D2H of the results
          inside the timed region, every step)
  --impl reference : the reference's CPU algorithm (the C oracle port -- no Fortran compiler exists here or on the GPU
          box) with OpenMP over zones on all host threads, same windows, bounded zone sample per step.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
DATA = os.path.join(ROOT, 'data')

WORKLOADS = {   # settings fragment, problem setup, description[, screening switch]
    'sn160_scrn': ('test_settings', 'setup_tnsn_sn160', 'SN160 (160 species, 2048 reactions), BE, screening on, th_sn1aflame', 1),
    'sn160': ('test_settings', 'setup_tnsn_sn160', 'SN160 (160 species, 2048 reactions), BE, th_sn1aflame'),
    'sn160_heat': ('test_settings_heat', 'setup_heat_sn160', 'SN160, BE, self-heating + screening, C/O burn (test_settings_heat)'),
    'sn231_logft': ('test_settings_logft', 'setup_logft_sn231', 'SN231 logft weak rates, BE, self-heating + screening, C/O burn (test_settings_logft)'),
    'torch47': ('test_settings', 'setup_tnsn_torch47', 'torch47 (47 species), BE, th_sn1aflame'),
    'torch47_batch': ('test_settings_batch', 'setup_batch_torch47', 'torch47, BE, 16 distinct th_batch/ab_batch fixtures, self-heating + screening (test_settings_batch)'),
    'alpha': ('test_settings_small', 'setup_tnsn_alpha', 'alpha (14 species), BE, th_sn1aflame'),
    'xrb304': ('test_settings', 'setup_xrb_fisker', 'XRB304 (304 species), BE, th_fisker_aboveign_b1'),
}
HIST_CAP = 8000          # records of the set-up run that supplies the restart points (kstmx of the shipped controls <= 5000)


def control_text(w):
    st, su = WORKLOADS[w][:2]
    txt = open(os.path.join(DATA, st)).read() + open(os.path.join(DATA, 'Test_Problems', su)).read()
    if len(WORKLOADS[w]) > 3 and WORKLOADS[w][3]:
        off = '0         Include Screening (yes=1)'
        assert off in txt
        txt = txt.replace(off, '1         Include Screening (yes=1)')
    return txt


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        import statistics
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = [float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': float(self.rows[0][1]), 'reasons': reasons,
                'samples': len(self.rows)}


def dist_setup(ngpus):
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    import torch
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend)
    return rank, world, local


def reduce_max_sum(world, tmax, vsum):
    """max over ranks of a time, sum over ranks of a count (only the metric bookkeeping is reduced;
    the hot loop has no communication)."""
    if world == 1:
        return tmax, vsum
    import torch
    import torch.distributed as dist
    dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    a = torch.tensor([tmax], dtype=torch.float64, device=dev); b = torch.tensor([vsum], dtype=torch.float64, device=dev)
    dist.all_reduce(a, op=dist.ReduceOp.MAX); dist.all_reduce(b, op=dist.ReduceOp.SUM)
    return float(a.item()), float(b.item())


def zone_block(rank, world, zones_per_gpu, nfix):
    """Block partition of the job's zones: rank r owns the global zones [r*Z, (r+1)*Z); global zone z
    evolves fixture z mod K.  Returns (global zone numbers, fixture index per zone)."""
    import numpy as np
    assert 0 <= rank < world and zones_per_gpu > 0 and nfix > 0
    gz = np.arange(rank * zones_per_gpu, (rank + 1) * zones_per_gpu)
    return gz, gz % nfix


def window_points(hist_n, points):
    """Restart points of one fixture: `points` record indices spread evenly over the accepted steps 1 .. n-2 of its
    complete run (record 0 is the initial state; the last record is tstop)."""
    import numpy as np
    n = int(hist_n)
    if n < 4:
        raise SystemExit('bench.py: the set-up run recorded only %d steps' % n)
    return np.unique(np.linspace(1, n - 2, min(points, n - 2)).astype(int))


def window_restart(hist, gz, nfix, points):
    """Restart state of the global zones gz: zone z -> fixture f = z mod K, point (z div K) mod M_f of that fixture.
    `hist` is the history of a complete run with one zone per fixture (hist_* arrays indexed [fixture, record])."""
    import numpy as np
    f = gz % nfix
    idx = np.zeros(len(gz), dtype=np.int64)
    for k in range(nfix):
        pts = window_points(hist['hist_n'][k], points)
        m = f == k
        idx[m] = pts[(gz[m] // nfix) % len(pts)]
    return dict(t=hist['hist_t'][f, idx], tdel=hist['hist_tdel'][f, idx], t9=hist['hist_t9'][f, idx],
                rho=hist['hist_rho'][f, idx], t9o=hist['hist_t9'][f, idx - 1], y=hist['hist_y'][f, idx],
                yo=hist['hist_y'][f, idx - 1], kstep0=hist['hist_kstep'][f, idx]), f, idx


def dist_teardown(world):
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def host_threads():
    """Threads the CPU arm uses: every hardware thread this process may run on (cgroup/affinity aware)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class CpuArm:
    """The reference's CPU algorithm for the same windows: C oracle (-O3 -march=native), OpenMP over zones."""
    def __init__(self, workload, points, window):
        os.environ.setdefault('OMP_PROC_BIND', 'true')      # one thread per hardware thread, no migration
        os.environ.setdefault('OMP_PLACES', 'threads')
        from oracle import xnet_oracle as O
        self.p = O.Problem.from_control(control_text(workload), DATA)
        self.threads = host_threads()
        self.nfix = len(self.p.fixtures)
        self.points, self.window = points, window
        self.hist = self.p.run(nzone=self.nfix, hist_cap=HIST_CAP, nthreads=self.threads, fast=True)

    def step(self, nzone):
        import numpy as np
        gz = np.arange(nzone)
        rs, f, _ = window_restart(self.hist, gz, self.nfix, self.points)
        t0 = time.perf_counter()
        r = self.p.run(nzone=nzone, nthreads=self.threads, fast=True, restart=rs, zfix=f, ctl_override=dict(kstmx=self.window))
        dt = time.perf_counter() - t0
        return float(r['ktot'][:, 1].sum()), dt


def cpu_sample_zones(points, nfix, threads, target_iters, window):
    """zones of one CPU step: a whole number of passes over the restart points, about target_iters Newton iterations"""
    per_pass = max(points * nfix, 1)
    passes = max(1, int(round(target_iters / float(per_pass * window))))
    return per_pass * passes


_REAL_STDOUT = None


def quiet_stdout():
    """Everything any library writes to stdout from here on (NCCL prints its version there) goes to stderr; the one
    JSON line of the contract is written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def run_reference(args):
    # no process group: the reference arm has nothing to exchange; under torchrun rank 0 alone works and prints
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    arm = CpuArm(args.workload, args.points, args.window)
    nz = args.cpu_zones or cpu_sample_zones(args.points, arm.nfix, arm.threads, 1.5e5, args.window)
    iters = []; times = []
    for i in range(args.warmup + args.steps):
        n, dt = arm.step(nz)
        if i >= args.warmup:
            iters.append(n); times.append(dt)
    tot_t = sum(times); value = sum(iters) / tot_t
    sample = ('%d zones per step (restart points x fixtures x passes), each up to %d accepted steps from its restart point: '
              'the GPU arm\'s windows on a bounded zone count; OpenMP over zones, C oracle -O3 -march=native (port: the Fortran '
              'reference cannot be built, no Fortran compiler)' % (nz, args.window))
    line = {'impl': 'reference', 'metric': 'zone-Newton-iters/sec', 'value': value, 'unit': 'zone-Newton-iterations/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * tot_t / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': workload_name(args, nz) + ' [CPU sample]', 'zones_per_step': nz,
                       'window_steps': args.window, 'restart_points': args.points},
            'cpu_baseline': {'value': value, 'unit': 'zone-Newton-iterations/s', 'cores': arm.threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'zone-Newton-iterations/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)


def workload_name(args, nz):
    return ('%s, %d zones per GPU; step = every zone advances up to %d accepted steps from one of %d restart points spread '
            'evenly over the complete trajectory (windows tile the trajectory)' % (WORKLOADS[args.workload][2], nz, args.window, args.points))


# ----------------------------------------------------------------------------------------------------------------
class GpuArm:
    """Product path for one workload: network, fixtures, the set-up run that supplies the restart points, contexts."""
    def __init__(self, workload, device, arith, solver, points, window):
        import xnet_b200 as X
        self.X = X
        self.workload, self.device, self.points, self.window = workload, device, points, window
        self.ctl = X.read_controls(control_text(workload))
        self.net = X.Network(os.path.join(DATA, self.ctl['data_dir']), self.ctl['iweak0'])
        self.fx = [X.load_fixture(self.net, os.path.join(DATA, a), os.path.join(DATA, t))
                   for a, t in zip(self.ctl['inab_file'], self.ctl['thermo_file'])]
        self.nfix = len(self.fx)
        self.arith, self.solver = arith, solver
        self.ctx_full = self.context()                    # complete trajectories (kstmx of the control file)
        self.hist = self.ctx_full.full_net(X.build_zone_set(self.fx, nzone=self.nfix), hist_cap=HIST_CAP)
        if (self.hist['hist_n'] > HIST_CAP).any() or (self.hist['status'] != -1).any():
            raise SystemExit('bench.py: the set-up run did not finish: status %s' % self.hist['status'])
        self.ctx = self.context(kstmx=window)             # the timed windows

    def context(self, arith=None, **ov):
        a = self.arith if arith is None else arith
        return self.X.Context(self.net, self.ctl, nzevolve=1, device=self.device, arith=a, solver=self.solver if a == 0 else 0, **ov)

    def zones(self, gz, pinned=False):
        rs, f, idx = window_restart(self.hist, gz, self.nfix, self.points)
        return self.X.build_zone_set(self.fx, nzone=len(gz), zfix=f, restart=rs, pinned=pinned)

    def close(self):
        self.ctx.close(); self.ctx_full.close()


def timed_resident(ctx, steps, flush, world, torch):
    ms = []
    for _ in range(steps):
        flush.zero_(); torch.cuda.synchronize()
        ms.append(ctx.run_resident())      # CUDA events on the launching stream, inside the library
    return ms


def lu_isolated(arm, batch, device):
    """The LU and the triangular solves alone on `batch` Jacobians sampled from the trajectory: xb_jacobian_decomp /
    xb_jacobian_bksub (module-level drop-ins, dense partial-pivoting LU) next to cublasDgetrfBatched / cublasDgetrsBatched
    -- what the reference's GPU build calls (source/xnet_linalg.F90:1847,1938) -- on the same matrices."""
    import numpy as np
    X = arm.X
    h = arm.hist
    pts = window_points(h['hist_n'][0], arm.points)
    idx = pts[np.arange(batch) % len(pts)]
    y = np.ascontiguousarray(h['hist_y'][0, idx]); t9 = h['hist_t9'][0, idx].copy(); rho = h['hist_rho'][0, idx].copy()
    rdt = 1.0 / h['hist_tdel'][0, idx + 1]
    ctx = X.Context(arm.net, arm.ctl, nzevolve=batch, device=device, arith=0)
    out = {'batch': batch, 'n': ctx.msize, 'matrices': 'assembled by xb_jacobian_build at %d states along the trajectory (diag = 1/dt of '
           'the step taken there, mult = -1), replicated to the batch' % len(pts)}
    try:
        ctx.cross_sect(y, t9, rho)
        ydot = ctx.yderiv(y)
        heat = ctx.msize > arm.net.ny
        t9dot = ctx.eos(batch)[3] if heat else None
        ctx.jacobian_build(batch, diag=rdt, mult=-np.ones(batch))
        jac = ctx.jac(batch)
        N = ctx.msize
        flops = (2.0 / 3.0) * N ** 3 * batch
        best_d, best_b = 1e30, 1e30
        for rep in range(3):
            ctx.set_jac(jac)
            ctx.jacobian_decomp(batch); md = ctx.last_kernel_ms()
            dy = ctx.jacobian_bksub(ydot, t9rhs=t9dot); mb = ctx.last_kernel_ms()
            if rep > 0:
                best_d = min(best_d, md); best_b = min(best_b, mb)
        x_ours = dy[0] if heat else dy
        out['xb'] = {'decomp_ms': best_d, 'bksub_ms': best_b, 'factorizations_per_s': batch / (best_d * 1e-3),
                     'lu_tflops': flops / (best_d * 1e-3) / 1e12,
                     'note': 'xb_jacobian_decomp reads and writes jac(msize,msize,nz) in HBM like the library call it replaces; '
                             'fast arithmetic (threshold partial pivoting)'}
        so = os.path.join(ROOT, 'tools', 'libxb_cublas_probe.so')
        if os.path.exists(so):
            pl = C.CDLL(so)
            pl.xbp_last_error.restype = C.c_char_p
            cdp = C.POINTER(C.c_double)
            pl.xbp_cublas_getrf_getrs.argtypes = [C.c_int, C.c_int, C.c_int, cdp, cdp, C.c_int, C.POINTER(C.c_float),
                                                  C.POINTER(C.c_float), cdp, C.POINTER(C.c_int), C.POINTER(C.c_int), cdp]
            rhs = np.zeros((batch, N)); rhs[:, :arm.net.ny] = ydot
            if heat:
                rhs[:, N - 1] = t9dot
            xs = np.zeros((batch, N)); mf = C.c_float(); msv = C.c_float()
            rc = pl.xbp_cublas_getrf_getrs(device, N, batch, jac.ctypes.data_as(cdp), rhs.ctypes.data_as(cdp), 2, C.byref(mf),
                                           C.byref(msv), None, None, None, xs.ctypes.data_as(cdp))
            if rc == 0:
                scale = np.abs(xs[:, :arm.net.ny]).max(axis=1, keepdims=True)
                out['cublas'] = {'getrf_ms': mf.value, 'getrs_ms': msv.value, 'factorizations_per_s': batch / (mf.value * 1e-3),
                                 'lu_tflops': flops / (mf.value * 1e-3) / 1e12,
                                 'max_abs_diff_of_solutions_over_max_abs': float((np.abs(xs[:, :arm.net.ny] - x_ours) / scale).max())}
                out['xb_over_cublas'] = {'getrf': mf.value / best_d, 'getrf_plus_getrs': (mf.value + msv.value) / (best_d + best_b)}
            else:
                out['cublas'] = {'error': pl.xbp_last_error().decode()}
        else:
            out['cublas'] = {'error': 'tools/libxb_cublas_probe.so not built (make -C tools)'}
    finally:
        ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--workload', default='sn160_scrn', choices=list(WORKLOADS),
                    help="default: BASELINE.json configs[2] (SN160 with screening, 16384 zones; BE, the integrator the reference builds)")
    ap.add_argument('--zones', type=int, default=16384, help='zones per GPU')
    ap.add_argument('--arith', default='fast', choices=['fast', 'strict'])
    ap.add_argument('--solver', default=os.environ.get('XB_BENCH_SOLVER', 'structured'), choices=['dense', 'structured'],
                    help='linear solver of the fused kernel in fast arithmetic: dense partial-pivoting LU or the structured '
                         '(skyline, light particles last) LU with dense fallback')
    ap.add_argument('--window', type=int, default=96, help='accepted steps per zone per bench step')
    ap.add_argument('--points', type=int, default=128, help='restart points per fixture')
    ap.add_argument('--cpu-zones', type=int, default=None)
    ap.add_argument('--lu-batch', type=int, default=16384, help='matrices of the isolated-LU / cuBLAS comparison (0 = skip)')
    ap.add_argument('--second', default='sn160_heat', help="second workload reported under also.second_workload ('' = none)")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='only the headline measurement (value, e2e, roofline)')
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == 'reference':
        return run_reference(args)

    import numpy as np
    import torch
    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback')
    torch.cuda.set_device(local)
    import xnet_b200 as X
    from xnet_b200 import api

    arith = 1 if args.arith == 'strict' else 0
    solver = 1 if args.solver == 'structured' else 0
    arm = GpuArm(args.workload, local, arith, solver, args.points, args.window)
    nz = args.zones
    gz, _ = zone_block(rank, world, nz, arm.nfix)
    zones = arm.zones(gz, pinned=True)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')   # > 126 MB L2

    fp64_peak = api.measure_fp64_peak(local)
    X.lib().xb_launch_count(1)
    ctx = arm.ctx

    # ---- device-resident timing (value) ----
    ctx.upload(zones)
    for _ in range(args.warmup):
        ctx.run_resident()
    sampler = ClockSampler(local); sampler.start()
    barrier(world); torch.cuda.synchronize()
    kern_ms = timed_resident(ctx, args.steps, flush, world, torch)
    torch.cuda.synchronize(); barrier(world)
    res = ctx.download()
    ph = ctx.phase_cycles()
    newton = float(res['ktot'][:, 1].sum()); nlu = float(res['ktot'][:, 2].sum())
    nts = float(res['ktot'][:, 0].sum())       # time-step attempts (ktot(1))
    nacc = float(np.where(res['status'] == -1, res['kstep'] - zones._keep['rs_kstep0'], args.window).sum())   # accepted steps
    t_dev = sum(kern_ms) * 1e-3
    t_dev_max, newton_all = reduce_max_sum(world, t_dev, newton * args.steps)
    value = newton_all / t_dev_max

    # ---- end-to-end through the C ABI with pinned host buffers (e2e): H2D + kernel + D2H every step ----
    ctx.full_net(zones, pinned=True, check=False)     # untimed: allocates the pinned result buffers' device side
    barrier(world); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r2 = ctx.full_net(zones, pinned=True, check=False)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier(world)
    sampler.stop_flag = True; sampler.join(timeout=2)
    h2d, d2h = ctx.copy_bytes()
    t_e2e_max, newton_e2e = reduce_max_sum(world, t_e2e, float(r2['ktot'][:, 1].sum()) * args.steps)
    launches_timed = X.lib().xb_launch_count(0)
    ok = bool(((res['status'] == -1) | (res['status'] == 0)).all())

    # ---- secondary measurements (all ranks take part where zones are evolved; rank 0 reports) ----
    also = {}
    if not args.no_extras:
        # (1) the same zones over their COMPLETE trajectories, once
        zfull = X.build_zone_set(arm.fx, nzone=nz, zfix=gz % arm.nfix)
        arm.ctx_full.upload(zfull)
        barrier(world); torch.cuda.synchronize()
        ms_full = arm.ctx_full.run_resident()
        rf = arm.ctx_full.download()
        tf, nf = reduce_max_sum(world, ms_full * 1e-3, float(rf['ktot'][:, 1].sum()))
        also['complete_trajectory'] = {'value': nf / tf, 'unit': 'zone-Newton-iterations/s', 'seconds': tf, 'launches': 1,
                                       'all_zones_finished': bool((rf['status'] == -1).all()),
                                       'accepted_steps_per_zone': float(rf['kstep'].mean())}
        # (2) strict arithmetic (the bit-exact mode: no FMA contraction, portable elementary functions, DGETRF's pivots) on the
        #     timed windows, once after one warm-up
        if arith == 0:
            cs = arm.context(arith=1, kstmx=args.window)
            cs.upload(zones); cs.run_resident()
            barrier(world); torch.cuda.synchronize()
            ms_s = cs.run_resident(); rs_ = cs.download(); cs.close()
            ts, ns = reduce_max_sum(world, ms_s * 1e-3, float(rs_['ktot'][:, 1].sum()))
            also['strict_value'] = ns / ts
            also['strict'] = {'value': ns / ts, 'unit': 'zone-Newton-iterations/s', 'ms_per_step': 1e3 * ts, 'steps': 1,
                              'note': 'arith=strict: bit-identical to the portable-math oracle (tests/test_gpu_parity.py)'}
            if solver == 1:     # dense fast-arithmetic LU on the same windows, for the A/B of the structured solver
                cd = arm.X.Context(arm.net, arm.ctl, nzevolve=1, device=local, arith=0, solver=0, kstmx=args.window)
                cd.upload(zones); cd.run_resident()
                barrier(world); torch.cuda.synchronize()
                ms_d = cd.run_resident(); rd_ = cd.download(); cd.close()
                td, nd = reduce_max_sum(world, ms_d * 1e-3, float(rd_['ktot'][:, 1].sum()))
                also['dense_solver'] = {'value': nd / td, 'unit': 'zone-Newton-iterations/s', 'ms_per_step': 1e3 * td, 'steps': 1}
        # (3) end to end with the per-step ts records AND reaction fluxes written out (what the reference's evolve timer
        #     includes: ts_output), on a subset of the zones (the records of 16384 zones would be 13 GB)
        nsub = min(nz, 2048)
        zsub = arm.zones(gz[:nsub], pinned=True)
        ctx.full_net(zsub, hist_cap=args.window + 1, flux=True, pinned=True, check=False)
        barrier(world); torch.cuda.synchronize()
        t0 = time.perf_counter()
        r3 = ctx.full_net(zsub, hist_cap=args.window + 1, flux=True, pinned=True, check=False)
        torch.cuda.synchronize(); dt3 = time.perf_counter() - t0
        b3 = ctx.copy_bytes()
        t3, n3 = reduce_max_sum(world, dt3, float(r3['ktot'][:, 1].sum()))
        also['e2e_with_ts_output'] = {'value': n3 / t3, 'unit': 'zone-Newton-iterations/s', 'zones_per_gpu': nsub,
                                      'h2d_bytes': b3[0], 'd2h_bytes': b3[1], 'seconds': t3,
                                      'note': 'xb_evolve with hist_cap = window+1 and flux output: every accepted step leaves a ts record '
                                              '(t, T9, rho, dt, edot, sqnu, y[ny], flx[mflx]) that is copied to the host'}
        # (5) the BDF integrator (isolv = 3) on the reference's own BDF problem (test 54: SN160, self-heating, screening,
        #     six C/O-burn trajectories), complete trajectories, dense solver
        if args.workload == 'sn160_scrn':
            try:
                txt_b = open(os.path.join(DATA, 'test_settings_bdf')).read() + open(os.path.join(DATA, 'Test_Problems', 'setup_bdf_sn160')).read()
                ctl_b = X.read_controls(txt_b)
                net_b = X.Network(os.path.join(DATA, ctl_b['data_dir']), ctl_b['iweak0'])
                fx_b = [X.load_fixture(net_b, os.path.join(DATA, a), os.path.join(DATA, t)) for a, t in zip(ctl_b['inab_file'], ctl_b['thermo_file'])]
                cb = X.Context(net_b, ctl_b, nzevolve=1, device=local, arith=0, solver=0)
                nzb = min(nz, 4096)
                gb, _ = zone_block(rank, world, nzb, len(fx_b))
                zb = X.build_zone_set(fx_b, nzone=nzb, zfix=gb % len(fx_b))
                cb.upload(zb); cb.run_resident()
                barrier(world); torch.cuda.synchronize()
                ms_b = cb.run_resident(); rb = cb.download(); cb.close()
                tb, nb_ = reduce_max_sum(world, ms_b * 1e-3, float(rb['ktot'][:, 1].sum()))
                also['bdf'] = {'workload': 'SN160, BDF (isolv = 3, variable order 1-5), self-heating + screening, six C/O-burn trajectories '
                                           '(test_settings_bdf + setup_bdf_sn160), %d zones per GPU, complete trajectories, dense LU' % nzb,
                               'value': nb_ / tb, 'unit': 'zone-Newton-iterations/s', 'seconds': tb,
                               'accepted_steps_per_zone': float(rb['kstep'].mean()),
                               'newton_iterations_per_attempt': float(rb['ktot'][:, 1].sum()) / max(float(rb['ktot'][:, 0].sum()), 1.0),
                               'all_zones_finished': bool((rb['status'] == -1).all())}
            except Exception as e:   # secondary measurement: never takes the headline line down
                also['bdf'] = {'error': str(e)[:200]}
        # (4) a workload that exercises retries, multi-iteration Newton and Jacobian reuse
        if args.second and args.second != args.workload:
            a2 = GpuArm(args.second, local, 0, solver, args.points, args.window)
            g2, _ = zone_block(rank, world, nz, a2.nfix)
            z2 = a2.zones(g2)
            a2.ctx.upload(z2); a2.ctx.run_resident()
            barrier(world); torch.cuda.synchronize()
            ms2 = timed_resident(a2.ctx, 3, flush, world, torch)
            torch.cuda.synchronize(); barrier(world)
            r_2 = a2.ctx.download()
            t2, n2 = reduce_max_sum(world, sum(ms2) * 1e-3, float(r_2['ktot'][:, 1].sum()) * 3)
            also['second_workload'] = {'workload': WORKLOADS[args.second][2] + ', %d zones per GPU, same windowed step' % nz,
                                       'value': n2 / t2, 'unit': 'zone-Newton-iterations/s', 'steps': 3, 'ms_per_step': 1e3 * t2 / 3,
                                       'newton_iterations_per_timestep_attempt': float(r_2['ktot'][:, 1].sum()) / max(float(r_2['ktot'][:, 0].sum()), 1.0),
                                       'lu_per_newton_iteration': float(r_2['ktot'][:, 2].sum()) / max(float(r_2['ktot'][:, 1].sum()), 1.0),
                                       'launch': a2.ctx.launch_config}
            a2.close()
    launches = X.lib().xb_launch_count(0)

    if rank != 0:
        arm.close()
        dist_teardown(world)
        return
    N = ctx.msize
    lu_flops = (2.0 / 3.0) * N ** 3 * nlu               # per launch, this rank
    ms_launch = sum(kern_ms) / len(kern_ms)
    achieved = lu_flops / (ms_launch * 1e-3) / 1e12
    phase_tot = max(ph['total'], 1.0)
    if os.environ.get('XB_DUMP_PHASES'):
        sys.stderr.write('phases per Newton iteration: %s\n' % {k: round(v / max(newton, 1.0), 1) for k, v in ph.items()})
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    slu = ctx.solver_info()
    # roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel at this zone count, from
    # the ncu --set full capture of the same command kept under profiles/ (tools/ncu_summary.py writes the file)
    traffic, traffic_src = None, 'profiles/r02d_evolve_slu_ncu_%dzones.txt' % nz
    try:
        rd = wr = None
        for ln in open(os.path.join(ROOT, traffic_src)):
            f = ln.split()
            if len(f) >= 3 and f[0] == 'dram__bytes_read.sum' and rd is None:
                rd = float(f[-1]) * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[f[1]]
            if len(f) >= 3 and f[0] == 'dram__bytes_write.sum' and wr is None:
                wr = float(f[-1]) * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[f[1]]
        if rd is not None and wr is not None and arith == 0 and args.solver == 'structured' and args.workload == 'sn160_scrn':
            traffic = rd + wr
    except Exception:
        pass
    roof = {'bound': 'fp64', 'achieved': achieved, 'peak': fp64_peak / 1e12, 'unit': 'TFLOP/s',
            'frac': achieved / (fp64_peak / 1e12), 'traffic': traffic,
            'kernel': ctx.kernel_name + ' (fused rates+ydot+Jacobian+LU+bksub+Newton/timestep control)',
            'algorithmic': '2/3*N^3 flops x ktot(3) LU factorizations per launch (N=%d, %d factorizations)' % (N, int(nlu)),
            'bound_note': 'FP64 compute (the LU): neither HBM nor the bf16 tensor peak bounds this path; the FP64 tensor '
                          'pipe (DMMA.8x8x4) has the same 64 FMA/clk/SM peak as DFMA on B200 (profiles/r02_dmma_probe.txt)',
            'peak_source': 'FP64 FMA peak measured live on this GPU (DFMA microbenchmark); MEASURED_PEAKS.json has only bf16/HBM entries',
            'phase_share': {k: round(ph[k] / phase_tot, 4) for k in ('rates', 'yderiv', 'jac', 'lu', 'bksub', 'other', 'lu_panel', 'lu_swap_trsm', 'lu_gemm')},
            'cycles_per_newton_iter': phase_tot / max(newton, 1.0),
            'traffic_note': 'bytes; dram__bytes_read + dram__bytes_write of one launch from the ncu capture ' + traffic_src +
                            ' (the Jacobian, csect and the zone state never touch HBM; HBM sees the restart state in, the '
                            'results out, the zone-invariant tables and the pivot schedule on their first touch)',
            'hbm_io_gbs': (h2d + d2h) / (ms_launch * 1e-3) / 1e9, 'hbm_peak_gbs': peaks.get('hbm_gbs', 6650.0)}
    if slu:
        # structured solver: the flops actually executed are a few per cent of the dense count; `achieved` above stays the
        # DENSE-EQUIVALENT rate (what a dense LU would have to sustain for the same factorisations per second)
        roof['achieved_is'] = 'dense-equivalent (2/3 N^3 per factorisation); the structured LU executes %.3g flops per factorisation' % slu['flops']
        roof['actual_lu_tflops'] = slu['flops'] * nlu / (ms_launch * 1e-3) / 1e12
        roof['structured'] = slu
        roof['dense_fallbacks_per_lu'] = ph['spare'] / max(nlu, 1.0)
        nfb = max(ph['spare'], 1.0)
        roof['fallback_stats'] = {'count': ph['spare'], 'mean_first_violated_column': ph['dbg0'] / nfb,
                                  'mean_lu_number_in_zone': ph['dbg1'] / nfb, 'mean_newton_iteration': ph['dbg2'] / nfb,
                                  'cycles_per_fallback': ph['lu_swap_trsm'] / nfb,
                                  'phase_note': 'lu_panel = failed structured attempts, lu_swap_trsm = attempts + dense fallback'}
    else:
        roof['pivot_steps_per_lu'] = ph['spare'] / max(nlu, 1.0)
    line = {
        'metric': 'zone-Newton-iters/sec', 'value': value, 'unit': 'zone-Newton-iterations/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t_dev_max / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(args, nz), 'zones_per_gpu': nz, 'arith': args.arith,
                   'solver': args.solver if arith == 0 else 'dense', 'window_steps': args.window, 'restart_points': args.points,
                   'l2': 'L2 flushed (256 MiB write) between timed iterations; tables are re-read from HBM each step',
                   'launch': ctx.launch_config, 'all_zones_ok': ok},
        'e2e': {'value': newton_e2e / t_e2e_max, 'unit': 'zone-Newton-iterations/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'steps': args.steps},
        'gpu_launches': int(launches_timed), 'gpu_launches_total': int(launches),
        # the same run in the reference's other counters (this rank's zones, device-resident timing)
        'also': dict({'accepted_steps_per_s': nacc * args.steps / t_dev, 'timestep_attempts_per_s': nts * args.steps / t_dev,
                      'lu_factorizations_per_s': nlu * args.steps / t_dev, 'scope': 'rank 0'}, **also),
        'clocks': sampler.summary(),
        'roofline': roof,
    }
    if world == 1 and not args.no_extras and args.lu_batch > 0:
        try:
            iso = lu_isolated(arm, args.lu_batch, local)
            line['roofline']['lu_isolated'] = dict(iso.get('xb', {}), batch=iso['batch'], n=iso['n'], matrices=iso['matrices'],
                                                   frac_of_fp64_peak=iso.get('xb', {}).get('lu_tflops', 0.0) / (fp64_peak / 1e12))
            line['roofline']['cublas_batched'] = dict(iso.get('cublas', {}), what='cublasDgetrfBatched + cublasDgetrsBatched (pointer-array API, '
                                                      'the reference GPU build: source/xnet_linalg.F90:1847,1938), same matrices, same GPU')
            if 'xb_over_cublas' in iso:
                line['roofline']['xb_over_cublas_speedup'] = iso['xb_over_cublas']
        except Exception as e:      # a probe must never take the headline down
            line['roofline']['lu_isolated'] = {'error': repr(e)}
    if world == 1 and not args.no_cpu_baseline:
        cpu = CpuArm(args.workload, args.points, args.window)
        cz = args.cpu_zones or cpu_sample_zones(args.points, cpu.nfix, cpu.threads, 6e5, args.window)
        n_cpu, dt_cpu = cpu.step(cz)
        line['cpu_baseline'] = {'value': n_cpu / dt_cpu, 'unit': 'zone-Newton-iterations/s', 'cores': cpu.threads, 'kind': 'port',
                                'sample': '%d zones, the same windows (up to %d accepted steps from %d restart points), OpenMP over zones with '
                                          'OMP_PROC_BIND=true on %d threads, C oracle -O3 -march=native (Fortran reference not buildable: '
                                          'no Fortran compiler); %.1f s' % (cz, args.window, args.points, cpu.threads, dt_cpu)}
    emit(line)
    arm.close()
    dist_teardown(world)


if __name__ == '__main__':
    main()
