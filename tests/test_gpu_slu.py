"""GPU tests of the structured LU (ctl.solver = 1), through the C ABI.

1. The solver alone (xb_slu_solve): Jacobians the oracle assembles along real trajectories, and random matrices with
   the network's structural pattern, against LAPACK dense solves -- right-hand side carried through the factorisation,
   separate forward/backward solves, and the dense fallback, each way.
2. The fused evolve with the structured solver against the dense solver and against the oracle: single-step restart
   parity over the whole trajectory (no accumulated divergence), accepted-step counts and end states of complete runs,
   bit-identical replicated zones, and the forced dense fallback.
Tolerances: a structured and a dense LU of the same matrix are two different, equally valid, rounding sequences, so the
bound is the conditioning of the systems (the oracle built with and without FMA differs by as much, test_gpu_parity)."""
import os

import numpy as np
import pytest

from conftest import DATA, ROOT
from test_slu_cpu import Slu, _systems

pytestmark = pytest.mark.gpu


def _ctx(gpu_problem, name, nz=1, solver=1, **ov):
    import xnet_b200 as X
    ctl, net, fx = gpu_problem(name)
    return X, ctl, net, fx, X.Context(net, ctl, nzevolve=nz, arith=0, solver=solver, **ov)


@pytest.mark.parametrize('name', ['alpha', 'torch47', 'sn160_scrn', 'torch47_heat', 'sn160_heat', 'sn231_logft', 'xrb304'])
def test_slu_solve_trajectory_jacobians(gpu_problem, oracle_problem, name):
    op = oracle_problem(name)
    ctl0, net0, _ = gpu_problem(name)
    heat = 1 if ctl0['iheat'] > 0 else 0
    sys_ = _systems(op, name, heat, 4)
    nz = len(sys_)
    X, ctl, net, fx, ctx = _ctx(gpu_problem, name, nz)
    info = ctx.solver_info()
    assert info is not None and info['msize'] == net.ny + heat
    a = np.stack([s[0].T for s in sys_])          # [zone, col, row]
    b = np.stack([s[1] for s in sys_])
    ref = np.stack([np.linalg.solve(s[0], s[1]) for s in sys_])
    # conditioning after row equilibration (the T9 row of a self-heating system is 1e5..1e9 times larger than the
    # abundance rows: that scaling is harmless to an elimination and must not loosen the bound)
    cond = np.array([np.linalg.cond(s[0] / np.abs(s[0]).max(axis=1, keepdims=True)) for s in sys_])
    scale = np.abs(ref).max(axis=1)
    tol = np.maximum(1e-13 * cond, 1e-12) * scale
    xs = {}
    for mode in (0, 1, 2, 3):
        x, fl = ctx.slu_solve(a, b, mode)
        assert (fl == 0).all(), (mode, fl)
        err = np.abs(x - ref).max(axis=1)
        assert (err <= tol).all(), (name, mode, (err / scale).max(), cond.max())
        xs[mode] = x
    # carried and separate right-hand sides: the same operations up to the order of two multiplications
    assert np.abs(xs[0] - xs[1]).max() <= 1e-11 * scale.max()
    ctx.close()


@pytest.mark.parametrize('name', ['torch47', 'sn160_scrn', 'sn160_heat'])
def test_slu_solve_random_pattern(gpu_problem, name):
    """Random diagonally dominant matrices with exactly the envelope's pattern (every stored entry non-zero: the
    elimination must then touch every slot of the store), plus one system whose first column violates the threshold."""
    ctl0, net0, _ = gpu_problem(name)
    heat = 1 if ctl0['iheat'] > 0 else 0
    slu = Slu(net0, heat)
    n = slu.n
    rng = np.random.default_rng(20261018)
    nz = 6
    a = np.zeros((nz, n, n))                       # [zone, row, col] equation order
    for z in range(nz):
        for pi in range(n):
            for pj in range(n):
                if slu.inside(pi, pj):
                    a[z, slu.perm[pi], slu.perm[pj]] = rng.uniform(-1, 1)
        a[z] += np.diag(np.sign(np.diag(a[z])) * (np.abs(a[z]).sum(axis=1) + 10.0))
    b = rng.uniform(-1, 1, (nz, n))
    # zone nz-1: a heavy sub-diagonal entry 1e4 times the pivot -> the threshold test must fire (first heavy column)
    r1 = slu.rows(0)[0]
    a[nz - 1, slu.perm[r1], slu.perm[0]] = 1.0e4 * a[nz - 1, slu.perm[0], slu.perm[0]]     # multiplier 1e4 > 1/tau = 1e3
    X, ctl, net, fx, ctx = _ctx(gpu_problem, name, nz)
    at = np.ascontiguousarray(a.transpose(0, 2, 1))
    ref = np.stack([np.linalg.solve(a[z], b[z]) for z in range(nz)])
    for mode in (0, 1, 2, 3):
        x, fl = ctx.slu_solve(at, b, mode)
        good = slice(0, nz - 1)
        assert (fl[good] == 0).all()
        assert np.abs(x[good] - ref[good]).max() <= 1e-11 * np.abs(ref[good]).max()
        if mode < 2:
            assert fl[nz - 1] & 1
        else:    # the dense fallback pivots: correct answer for the violated system too
            assert np.abs(x[nz - 1] - ref[nz - 1]).max() <= 1e-10 * np.abs(ref[nz - 1]).max()
    # an entry outside the envelope is reported
    bad = at.copy()
    pi, pj = next((i, j) for i in range(n) for j in range(n) if not slu.inside(i, j))
    bad[0, slu.perm[pj], slu.perm[pi]] = 1.0
    x, fl = ctx.slu_solve(bad, b, 0)
    assert fl[0] & 2
    ctx.close()


@pytest.mark.parametrize('name,stride', [('torch47', 7), ('sn160_scrn', 9), ('sn160_heat', 5), ('sn231_logft', 7), ('xrb304', 23)])
def test_slu_single_step_restart(gpu_problem, oracle_problem, name, stride):
    """One accepted step from recorded oracle states over the WHOLE trajectory, structured solver vs the libm oracle and vs
    the dense solver: same step taken; abundances within 1e-9*max(Y) of the oracle and 1e-6 relative on Y > 1e-12 (the
    bounds of the dense fast path, test_gpu_parity.test_fast_single_step_restart)."""
    op = oracle_problem(name)
    h = op.run(nzone=1, hist_cap=6000)
    n = int(h['hist_n'][0])
    idx = np.arange(1, n - 1, stride)
    rs = dict(t=h['hist_t'][0, idx], tdel=h['hist_tdel'][0, idx], t9=h['hist_t9'][0, idx], rho=h['hist_rho'][0, idx],
              t9o=h['hist_t9'][0, idx - 1], y=h['hist_y'][0, idx], yo=h['hist_y'][0, idx - 1], kstep0=h['hist_kstep'][0, idx])
    res = {}
    for solver in (1, 0):
        X, ctl, net, fx, ctx = _ctx(gpu_problem, name, 1, solver=solver, kstmx=1)
        zs = X.build_zone_set(fx, nzone=len(idx), zfix=np.zeros(len(idx), int), restart=rs)
        res[solver] = ctx.full_net(zs, check=False)
        ctx.close()
    ref = h['hist_y'][0, idx + 1]
    heat = ctl['iheat'] > 0
    for solver in (1, 0):
        r = res[solver]
        assert np.abs(r['tdel'] / h['hist_tdel'][0, idx + 1] - 1).max() < 1e-10, solver
        e = np.abs(r['y'] - ref).max(axis=1) / ref.max(axis=1)
        m = ref > 1e-12
        e12 = (np.abs(r['y'] - ref)[m] / ref[m]).max()
        print('\n%s solver %d: %d restarts, |dY|/max(Y): median %.2g  90%% %.2g  max %.2g;  rel(Y>1e-12) max %.2g' % (
            name, solver, len(idx), np.median(e), np.percentile(e, 90), e.max(), e12))
        if not heat:
            assert e.max() < 1e-9 and e12 < 1e-6, solver
        else:
            # Self-heating runs end in NSE, where the Newton matrix is ill-conditioned (LAPACK's own solution of those
            # systems is only good to 1e-8 of max|dy| against an extended-precision solve): any two roundings of the same
            # step differ at that level.  Measured on B200, both solvers alike: 90 % of the restarts within 3e-10, the
            # worst (last steps before NSE) 5.4e-8 (SN160) / 3.1e-8 (SN231).
            assert np.percentile(e, 90) < 1e-9 and e.max() < 1e-6 and e12 < 1e-4, solver
    assert np.array_equal(res[1]['ktot'][:, :2], res[0]['ktot'][:, :2])     # same attempts and Newton iterations
    if heat:
        assert np.abs(res[1]['t9'] / h['hist_t9'][0, idx + 1] - 1).max() < 1e-7


@pytest.mark.parametrize('name', ['alpha', 'torch47', 'sn160_scrn', 'sn160_heat'])
def test_slu_complete_trajectory(gpu_problem, oracle_problem, name):
    """Complete runs: structured vs dense solver (both fast arithmetic) and vs the oracle: finished, accepted-step counts
    within 1 %, mass fractions X > 1e-6 within 1e-2 (the bound of test_fast_end_state), mass conserved, replicas identical."""
    out = {}
    for solver in (1, 0):
        X, ctl, net, fx, ctx = _ctx(gpu_problem, name, 3, solver=solver)
        out[solver] = ctx.full_net(X.build_zone_set(fx, nzone=3 * len(fx)))
        ctx.close()
    o = oracle_problem(name).run(nzone=len(fx))
    aa = net.dtable('aa')
    K = len(fx)
    for f in range(K):
        xs, xd, xo = out[1]['y'][f] * aa, out[0]['y'][f] * aa, o['y'][f] * aa
        assert out[1]['status'][f] == -1 and out[1]['t'][f] == o['t'][f]
        assert abs(int(out[1]['kstep'][f]) - int(o['kstep'][f])) <= max(2, 0.01 * o['kstep'][f])
        assert abs(xs.sum() - 1.0) < (1e-7 if ctl['iheat'] > 0 else 1e-8)
        m = xo > 1e-6
        assert np.abs(xs[m] / xo[m] - 1).max() < 1e-2 and np.abs(xs[m] / xd[m] - 1).max() < 1e-2
        for rep in (1, 2):
            assert np.array_equal(out[1]['y'][f + rep * K], out[1]['y'][f])
            assert np.array_equal(out[1]['ktot'][f + rep * K], out[1]['ktot'][f])


def test_slu_forced_fallback(gpu_problem, monkeypatch):
    """Every third factorisation is handed to the dense partial-pivoting fallback (XB_SLU_FORCE_FALLBACK): the run must be
    the same run to solver accuracy -- same accepted steps, abundances within the restart-parity bound."""
    name = 'torch47'
    import xnet_b200 as X
    ctl, net, fx = gpu_problem(name)
    zs = X.build_zone_set(fx, nzone=2)
    ctx = X.Context(net, ctl, nzevolve=1, arith=0, solver=1, kstmx=400)
    a = ctx.full_net(zs, hist_cap=401, check=False); ctx.close()
    monkeypatch.setenv('XB_SLU_FORCE_FALLBACK', '3')
    ctx = X.Context(net, ctl, nzevolve=1, arith=0, solver=1, kstmx=400)
    b = ctx.full_net(zs, hist_cap=401, check=False)
    ph = ctx.phase_cycles(); ctx.close()
    assert ph['spare'] >= 2 * 100          # fallbacks were taken
    n = 401
    assert int(a['hist_n'][0]) == n == int(b['hist_n'][0])
    assert np.array_equal(a['hist_kmon'][0, :n], b['hist_kmon'][0, :n])
    assert np.abs(b['hist_tdel'][0, 1:n] / a['hist_tdel'][0, 1:n] - 1).max() < 1e-8
    ya, yb = a['hist_y'][0, :n], b['hist_y'][0, :n]
    assert (np.abs(ya - yb).max(axis=1) / ya.max(axis=1)).max() < 1e-9
