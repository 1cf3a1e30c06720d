#!/usr/bin/env python
"""bench.py -- zone-Newton-iterations/s of the batched per-zone implicit step (BASELINE.json metric).

A "step" is one complete evolve (full_net) of this rank's zones: SN160 network, backward Euler,
th_sn1aflame trajectory + Data_SN160/ab_co replicated to the zone count (SURVEY 8d config 3 with
BE, the only integrator the reference builds).  Zones are independent, so N GPUs each evolve
their own block of zones (weak scaling, no data-path collective); whole-job value = sum of the
reference's own counter ktot(2) over all zones / max-over-ranks time.

  value : inputs resident in HBM, CUDA-event time of the evolve kernel (xb_evolve_resident)
  e2e   : the public C-ABI call xb_evolve with HOST buffers (H2D + kernel + D2H inside the timing)
  --impl reference : the reference's CPU algorithm (the C oracle port; the Fortran reference cannot
          be compiled here, no Fortran compiler) on all host threads, bounded zone sample.
"""
import argparse
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
    'alpha': ('test_settings_small', 'setup_tnsn_alpha', 'alpha (14 species), BE, th_sn1aflame'),
    'xrb304': ('test_settings', 'setup_xrb_fisker', 'XRB304 (304 species), BE, th_fisker_aboveign_b1'),
}


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
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': float(self.rows[0][1]), 'reasons': reasons}


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


def dist_teardown(world):
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def cpu_run(workload, nzone, threads, kstmx=None):
    """Times the CPU oracle (-O3 -march=native build, OpenMP over zones) on `nzone` zones."""
    from oracle import xnet_oracle as O
    p = O.Problem.from_control(control_text(workload), DATA)
    ov = dict(kstmx=kstmx) if kstmx else None
    t0 = time.perf_counter()
    r = p.run(nzone=nzone, nthreads=threads, fast=True, ctl_override=ov)
    dt = time.perf_counter() - t0
    return float(r['ktot'][:, 1].sum()), dt


def run_reference(args):
    # no process group: the reference arm has nothing to exchange; under torchrun rank 0 alone works and prints
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nz = args.cpu_zones or max(cores, 8)
    iters = []; times = []
    for i in range(args.warmup + args.steps):
        n, dt = cpu_run(args.workload, nz, cores, args.kstmx)
        if i >= args.warmup:
            iters.append(n); times.append(dt)
    tot_t = sum(times); value = sum(iters) / tot_t
    line = {'impl': 'reference', 'metric': 'zone-Newton-iters/sec', 'value': value, 'unit': 'zone-Newton-iterations/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * tot_t / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOADS[args.workload][2] + ', %d-zone sample per step' % nz, 'zones_per_step': nz,
                       'kstmx': args.kstmx},
            'cpu_baseline': {'value': value, 'unit': 'zone-Newton-iterations/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d zones x complete trajectory per step, OpenMP over zones, oracle -O3 -march=native '
                                       '(the Fortran reference cannot be built: no Fortran compiler)' % nz},
            'e2e': {'value': value, 'unit': 'zone-Newton-iterations/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--workload', default='sn160_scrn', choices=list(WORKLOADS),
                    help="default: BASELINE.json configs[2] (SN160 with screening, 16384 zones; BE, the integrator the reference builds)")
    ap.add_argument('--zones', type=int, default=16384, help='zones per GPU')
    ap.add_argument('--arith', default='fast', choices=['fast', 'strict'])
    ap.add_argument('--kstmx', type=int, default=None, help='cap on accepted steps per zone (default: complete trajectory)')
    ap.add_argument('--cpu-zones', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
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

    ctl = X.read_controls(control_text(args.workload))
    net = X.Network(os.path.join(DATA, ctl['data_dir']), ctl['iweak0'])
    fx = [X.load_fixture(net, os.path.join(DATA, a), os.path.join(DATA, t))
          for a, t in zip(ctl['inab_file'], ctl['thermo_file'])]
    ov = dict(kstmx=args.kstmx) if args.kstmx else {}
    ctx = X.Context(net, ctl, nzevolve=1, device=local, arith=1 if args.arith == 'strict' else 0, **ov)
    nz = args.zones
    # zone z of the whole job uses fixture (global z) mod K; rank r owns the block [r*nz, (r+1)*nz)
    gz, zfix = zone_block(rank, world, nz, len(fx))
    zones = X.build_zone_set(fx, nzone=nz, zfix=zfix)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')   # > 126 MB L2

    fp64_peak = api.measure_fp64_peak(local)
    X.lib().xb_launch_count(1)

    # ---- device-resident timing (value) ----
    ctx.upload(zones)
    for _ in range(args.warmup):
        ctx.run_resident()
    sampler = ClockSampler(local); sampler.start()
    barrier(world); torch.cuda.synchronize()
    kern_ms = []
    for _ in range(args.steps):
        flush.zero_(); torch.cuda.synchronize()
        kern_ms.append(ctx.run_resident())      # CUDA events on the launching stream, inside the library
    torch.cuda.synchronize(); barrier(world)
    res = ctx.download()
    ph = ctx.phase_cycles()
    newton = float(res['ktot'][:, 1].sum()); nlu = float(res['ktot'][:, 2].sum())
    nts = float(res['ktot'][:, 0].sum())       # time-step attempts (ktot(1)); accepted steps = kstep per zone
    nacc = float(np.where(res['status'] == -1, res['kstep'], args.kstmx or 0).sum())
    t_dev = sum(kern_ms) * 1e-3
    t_dev_max, newton_all = reduce_max_sum(world, t_dev, newton * args.steps)
    value = newton_all / t_dev_max

    # ---- end-to-end through the C ABI with host buffers (e2e) ----
    # the kernel is warm from the resident runs: one untimed call touches the pinned host buffers, then at most two
    # timed calls (a step is tens of seconds at the full zone count; the e2e figure does not need K of them)
    e2e_steps = max(1, min(args.steps, 2))
    ctx.full_net(zones, pinned=True)
    barrier(world); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r2 = ctx.full_net(zones, pinned=True)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier(world)
    sampler.stop_flag = True; sampler.join(timeout=2)
    h2d, d2h = ctx.copy_bytes()
    t_e2e_max, newton_e2e = reduce_max_sum(world, t_e2e, float(r2['ktot'][:, 1].sum()) * e2e_steps)
    launches = X.lib().xb_launch_count(0)
    ok = bool((res['status'] == -1).all())

    if rank != 0:
        ctx.close()
        dist_teardown(world)
        return
    N = ctx.msize
    lu_flops = (2.0 / 3.0) * N ** 3 * nlu               # per launch, this rank
    ms_launch = sum(kern_ms) / len(kern_ms)
    achieved = lu_flops / (ms_launch * 1e-3) / 1e12
    phase_tot = max(ph['total'], 1.0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    line = {
        'metric': 'zone-Newton-iters/sec', 'value': value, 'unit': 'zone-Newton-iterations/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t_dev_max / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOADS[args.workload][2] + ', %d zones per GPU, complete trajectory%s' % (
                       nz, '' if not args.kstmx else ' capped at %d accepted steps' % args.kstmx),
                   'zones_per_gpu': nz, 'arith': args.arith, 'kstmx': args.kstmx,
                   'l2': 'L2 flushed (256 MiB write) between timed iterations; tables are re-read from HBM each step',
                   'launch': ctx.launch_config, 'all_zones_finished': ok},
        'e2e': {'value': newton_e2e / t_e2e_max, 'unit': 'zone-Newton-iterations/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'steps': e2e_steps},
        'gpu_launches': int(launches),
        # the same run in the reference's other counters (this rank's zones, device-resident timing)
        'also': {'accepted_steps_per_s': nacc * args.steps / t_dev, 'timestep_attempts_per_s': nts * args.steps / t_dev,
                 'lu_factorizations_per_s': nlu * args.steps / t_dev, 'scope': 'rank 0'},
        'clocks': sampler.summary(),
        'roofline': {'bound': 'fp64', 'achieved': achieved, 'peak': fp64_peak / 1e12, 'unit': 'TFLOP/s',
                     'frac': achieved / (fp64_peak / 1e12), 'traffic': None,
                     'kernel': 'evolve_kernel (fused rates+ydot+Jacobian+LU+bksub+Newton/timestep control)',
                     'algorithmic': '2/3*N^3 flops x ktot(3) LU factorizations per launch (N=%d, %d factorizations)' % (N, int(nlu)),
                     'bound_note': 'FP64 compute (the LU): neither HBM nor the bf16 tensor peak bounds this path; the FP64 tensor '
                                   'pipe (DMMA.8x8x4) has the same 64 FMA/clk/SM peak as DFMA on B200 (tools/dmma_probe.cu)',
                     'peak_source': 'FP64 FMA peak measured live on this GPU (DFMA microbenchmark); MEASURED_PEAKS.json has only bf16/HBM entries',
                     'phase_share': {k: round(ph[k] / phase_tot, 4) for k in ('rates', 'yderiv', 'jac', 'lu', 'bksub', 'other', 'lu_panel', 'lu_swap_trsm', 'lu_gemm')},
                     'cycles_per_newton_iter': phase_tot / max(newton, 1.0),
                     'pivot_steps_per_lu': ph['spare'] / max(nlu, 1.0),
                     'partially_accepted_panels_per_lu': ph['dbg8'] / max(nlu, 1.0),
                     'dbg_cycles_per_newton_iter': [round(ph['dbg%d' % k] / max(newton, 1.0)) for k in range(12) if k != 8],
                     'hbm_note': 'Jacobian/csect/state never touch HBM (shared memory / L2 scratch); HBM traffic per launch is '
                                 'inputs+outputs only: %.3g GB/s of measured %.0f GB/s' % (
                                     (h2d + d2h) / (ms_launch * 1e-3) / 1e9, peaks.get('hbm_gbs', 6650.0))},
    }
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cz = args.cpu_zones or max(cores, 8)
        n_cpu, dt_cpu = cpu_run(args.workload, cz, cores, args.kstmx)
        line['cpu_baseline'] = {'value': n_cpu / dt_cpu, 'unit': 'zone-Newton-iterations/s', 'cores': cores, 'kind': 'port',
                                'sample': '%d zones x complete trajectory, OpenMP over zones, C oracle -O3 -march=native '
                                          '(Fortran reference not buildable: no Fortran compiler)' % cz}
    print(json.dumps(line))
    ctx.close()
    dist_teardown(world)


if __name__ == '__main__':
    main()
