"""BDF integrator on the device (xnet_b200/csrc/xb_bdf.cuh from source/xnet_integrate_bdf.F90), through the
C ABI: test/test_xnet.sh test 54 (test_settings_bdf + setup_bdf_sn160: SN160, self-heating, screening, six C/O-burn
trajectories, isolv = 3) and the same integrator on the small networks.

strict arithmetic: accepted-step sequence (dt, T9), abundances, order/iteration counters bit-identical to the
portable-math oracle over the complete runs of every trajectory; fast arithmetic: same run lengths to a few per cent and the
same end state to the integrator's own tolerance (tolc = 1e-4 relative per step)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bdf_controls(gpu_problem_all, name):
    import xnet_b200 as X
    ctl, net, fx = gpu_problem_all(name)
    return X, ctl, net, fx


def test_bdf_strict_bit_exact_sn160(gpu_problem_all, oracle_problem):
    X, ctl, net, fx = _bdf_controls(gpu_problem_all, 'sn160_bdf')
    K = len(fx)
    assert K == 6 and ctl['isolv'] == 3
    cap = 600
    ctx = X.Context(net, ctl, nzevolve=1, arith=1)
    r = ctx.full_net(X.build_zone_set(fx, nzone=2 * K), hist_cap=cap)
    ctx.close()
    o = oracle_problem('sn160_bdf').run(nzone=K, hist_cap=cap, fast='pm', nthreads=0)
    for z in range(2 * K):
        f = z % K
        n = int(o['hist_n'][f])
        assert 100 < n <= cap and int(r['hist_n'][z]) == n, (z, n, int(r['hist_n'][z]))
        assert np.array_equal(r['hist_tdel'][z, :n], o['hist_tdel'][f, :n])
        assert np.array_equal(r['hist_t9'][z, :n], o['hist_t9'][f, :n])
        assert np.array_equal(r['hist_y'][z, :n], o['hist_y'][f, :n])
        assert np.array_equal(r['hist_kmon'][z, :n], o['hist_kmon'][f, :n])
        assert np.array_equal(r['ktot'][z], o['ktot'][f]) and r['status'][z] == -1


@pytest.mark.parametrize('name,kstmx', [('alpha_heat', None), ('torch47_heat', None), ('alpha', 500), ('torch47', 500), ('sn160_scrn', 300)])
def test_bdf_strict_bit_exact_other_networks(gpu_problem, oracle_problem, name, kstmx):
    """isolv = 3 on problems the reference ships for backward Euler: the self-heating burns complete; the
    trajectory-driven flame problems (their controls are tuned for backward Euler: max norm, ymin = 1e-30; BDF resolves
    the flame with steps of 1e-5 s and does not finish inside their step limit, on either side) are compared over their
    first kstmx steps, which exercise the trajectory lookup of retried steps and the order changes."""
    import xnet_b200 as X
    from oracle import xnet_oracle as O
    from conftest import DATA, PROBLEMS, control_text
    txt = control_text(*PROBLEMS[name]).replace('1         Choice of integration Scheme', '3         Choice of integration Scheme')
    assert '3         Choice of integration Scheme' in txt
    ctl = X.read_controls(txt)
    assert ctl['isolv'] == 3 and ctl['changemx'] == 1.0e10
    _, net, fx = gpu_problem(name)
    ov = {} if kstmx is None else {'kstmx': kstmx}
    ctx = X.Context(net, ctl, nzevolve=1, arith=1, **ov)
    r = ctx.full_net(X.build_zone_set(fx, nzone=2), hist_cap=4000, check=False)
    ctx.close()
    o = O.Problem.from_control(txt, DATA).run(nzone=1, hist_cap=4000, fast='pm', ctl_override=ov or None)
    n = int(o['hist_n'][0])
    assert o['status'][0] == (-1 if kstmx is None else 0) and n > 100
    for z in range(2):
        assert int(r['hist_n'][z]) == n and r['status'][z] == o['status'][0]
        assert np.array_equal(r['hist_tdel'][z, :n], o['hist_tdel'][0, :n])
        assert np.array_equal(r['hist_t9'][z, :n], o['hist_t9'][0, :n])
        assert np.array_equal(r['hist_y'][z, :n], o['hist_y'][0, :n])
        assert np.array_equal(r['hist_kmon'][z, :n], o['hist_kmon'][0, :n])
        assert np.array_equal(r['ktot'][z], o['ktot'][0])


def test_bdf_fast_sn160(gpu_problem_all, oracle_problem):
    X, ctl, net, fx = _bdf_controls(gpu_problem_all, 'sn160_bdf')
    K = len(fx)
    ctx = X.Context(net, ctl, nzevolve=1, arith=0)
    r = ctx.full_net(X.build_zone_set(fx, nzone=K), hist_cap=600)
    ctx.close()
    o = oracle_problem('sn160_bdf').run(nzone=K, hist_cap=600, fast=False)
    aa = net.dtable('aa')
    for z in range(K):
        assert r['status'][z] == -1
        assert abs(int(r['kstep'][z]) - int(o['kstep'][z])) <= max(3, 0.05 * o['kstep'][z]), (z, r['kstep'][z], o['kstep'][z])
        assert abs(r['t9'][z] / o['t9'][z] - 1) < 1e-5
        m = o['y'][z] > 1e-8
        assert np.abs(r['y'][z][m] / o['y'][z][m] - 1).max() < 2e-3      # step sequences differ in the last digits: O(tolc) apart
        # mass is conserved to the integrator's tolerance (tolc = 1e-4 per step), as in the oracle's run
        mg, mo = (r['y'][z] * aa).sum(), (o['y'][z] * aa).sum()
        assert abs(mg - mo) < 1e-6 and abs(mg - 1.0) < 1e-3
