"""CPU tests of the oracle (no GPU): reference golden vectors, LAPACK pin, self-consistency."""
import os

import numpy as np
import pytest

from conftest import DATA, ROOT, control_text, PROBLEMS

GOLD = os.path.join(ROOT, 'tests', 'golden')
NETS = ['alpha', 'torch47', 'SN160', 'SN231', 'XRB304']


@pytest.mark.parametrize('net', NETS)
def test_net_desc_golden(net):
    """The reference's own known-answer files: species count, la/le of the 11 REACLIB chapters,
    nreac(1:4), nan(1:4), nffn/nnu (written at xnet_preprocess.F90:971-980) are reproduced exactly."""
    from oracle import xnet_loader as L
    nuc, onet, _ = L.load_network(os.path.join(DATA, 'Data_' + net))
    got = L.net_desc_text(nuc, onet)
    ref = open(os.path.join(GOLD, 'net_desc_%s.txt' % net)).read().split('\n')[1:1 + len(got)]
    assert [g.rstrip() for g in got] == [r.rstrip() for r in ref]


@pytest.mark.parametrize('net', NETS)
def test_map_invariants(net):
    """Structural invariants: nan(n)=le(n,ny); extents tile 1..nan; every reaction
    conserves baryon number, so sum_i aa_i*a*... = 0 per reaction."""
    from oracle import xnet_loader as L
    nuc, onet, _ = L.load_network(os.path.join(DATA, 'Data_' + net))
    ny = nuc['ny']
    for c in range(4):
        la, le = onet['la'][c], onet['le'][c]
        assert le[ny - 1] == onet['nan'][c]
        assert la[0] == 1 and np.all(la[1:] == le[:-1] + 1)
        # baryon conservation: sum over registered terms of aa(species)*a, grouped by reaction
        tot = np.zeros(onet['nreac'][c])
        for i in range(ny):
            for j in range(la[i] - 1, le[i]):
                tot[onet['mu%d' % (c + 1)][j] - 1] += nuc['aa'][i] * onet['a%d' % (c + 1)][j]
        fac = [1, 1, 2, 6, 24]
        assert np.allclose(tot * fac[c + 1], np.round(tot * fac[c + 1]), atol=1e-9)
        iw = onet['iwk%d' % (c + 1)]
        assert np.all(np.abs(tot) < 1e-9), (net, c, np.abs(tot).max())
        assert set(np.unique(iw)) <= {0, 1, 2, 3, 4, 7, 8}


@pytest.mark.parametrize('n', [1, 2, 5, 14, 47, 64, 65, 130, 160, 161, 232, 304])
def test_lu_against_lapack(built, n):
    """ox_dgetrf/ox_dgetrs (restated DGETRF NB=64 + recursive DGETRF2, DGETRS) against LAPACK as
    shipped by scipy: identical pivot sequence, factors equal to rounding, small residual."""
    from scipy.linalg import lapack
    from oracle import xnet_oracle as O
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n)) + np.diag(rng.standard_normal(n) * 3)
    lu, piv, info = O.dgetrf(a)
    lu2, piv2, info2 = lapack.dgetrf(a)
    assert info == 0 and info2 == 0
    assert np.array_equal(piv, piv2 + 1)
    assert np.abs(lu - lu2).max() <= 1e-11 * max(1.0, np.abs(lu2).max())
    b = rng.standard_normal(n)
    x = O.dgetrs(lu, piv, b)
    assert np.abs(a @ x - b).max() < 1e-9


def test_lu_singular_info(built):
    from oracle import xnet_oracle as O
    a = np.eye(6); a[3, 3] = 0.0
    _, _, info = O.dgetrf(a)
    assert info == 4
    a = np.zeros((3, 3))
    _, piv, info = O.dgetrf(a)
    assert info == 1 and list(piv) == [1, 2, 3]


def test_pmath_accuracy(built):
    """Portable elementary functions stay within a few ulp of libm (exp/log <=1, log10 <=2, 10^x <=6).
    Checked through the rates: csect = exp(h) with |h| up to a few hundred, so a few-ulp difference in
    the T9 powers shows up as <= |h|*few*eps ~ 1e-12 relative between the two oracle builds."""
    from oracle import xnet_oracle as O
    import ctypes as C
    p = O.Problem.from_control(control_text(*PROBLEMS['torch47']), DATA)
    ny = p.nuc['ny']
    y = p.fixtures[0]['ystart']
    for t9, rho in ((0.05, 1e9), (0.9, 3e8), (3.7, 2e7)):
        out = []
        for key in ('strict', 'pm'):
            lib = O.lib(False if key == 'strict' else 'pm')
            nrt = sum(p.net['nreac']); cs = np.zeros(nrt); yd = np.zeros(ny); jac = np.zeros(ny * ny); ye = np.zeros(1)
            yy = np.ascontiguousarray(y)
            rc = lib.ox_eval_point(C.byref(p.cnet), C.byref(p.cctl), t9, rho, O._pd(yy), 0.0, 1.0, 0.0, 1e3, -1.0,
                                   O._pd(cs), O._pd(yd), O._pd(jac), O._pd(ye))
            assert rc == 0
            out.append(cs)
        m = out[0] > 0
        assert np.abs(out[1][m] / out[0][m] - 1).max() < 2e-12


def test_oracle_matches_committed_golden(oracle_problem):
    """The oracle reproduces the committed golden trajectories bit for bit (guards the fixtures and
    the generating script)."""
    for name, cap in (('alpha', 3000), ('torch47', 400), ('torch47_scrn', 400), ('torch47_heat', 300), ('sn160_heat', 120)):
        for math, fast in (('libm', False), ('pm', 'pm')):
            g = np.load(os.path.join(GOLD, 'ts_%s_%s.npz' % (name, math)))
            r = oracle_problem(name).run(nzone=1, hist_cap=cap, fast=fast,
                                         ctl_override=dict(kstmx=int(g['kstmx'])))
            n = len(g['kstep'])
            assert np.array_equal(r['hist_tdel'][0, :n], g['tdel'])
            assert np.array_equal(r['hist_y'][0, :n], g['y'])


def test_alpha_physics(oracle_problem):
    """Weak physical pins: mass conservation and runaway of C/O into the Si/Fe group."""
    p = oracle_problem('alpha')
    r = p.run(nzone=1)
    x = r['y'][0] * p.nuc['aa']
    assert r['status'][0] == -1 and abs(r['t'][0] - 2.0) < 1e-12
    assert abs(x.sum() - 1.0) < 1e-9
    names = [n.strip() for n in p.nuc['names']]
    assert x[names.index('c12')] < 1e-3 and x[names.index('si28')] + x[names.index('s32')] > 0.5


def test_restart_reproduces_history(oracle_problem):
    """Restarting from any recorded accepted step reproduces the next record bit for bit."""
    p = oracle_problem('torch47')
    h = p.run(nzone=1, hist_cap=700, ctl_override=dict(kstmx=650))
    idx = np.arange(1, 640, 7)
    rs = dict(t=h['hist_t'][0, idx], tdel=h['hist_tdel'][0, idx], t9=h['hist_t9'][0, idx], rho=h['hist_rho'][0, idx],
              t9o=h['hist_t9'][0, idx - 1], y=h['hist_y'][0, idx], yo=h['hist_y'][0, idx - 1], kstep0=h['hist_kstep'][0, idx])
    r = p.run(nzone=len(idx), restart=rs, zfix=np.zeros(len(idx), int), ctl_override=dict(kstmx=1))
    assert np.array_equal(r['y'], h['hist_y'][0, idx + 1])
    assert np.array_equal(r['tdel'], h['hist_tdel'][0, idx + 1])


def test_zone_replication_and_threads(oracle_problem):
    """Zones sharing a fixture give bitwise identical results, independent of the thread count."""
    p = oracle_problem('alpha')
    a = p.run(nzone=5, nthreads=1)
    b = p.run(nzone=5, nthreads=4)
    assert np.array_equal(a['y'], b['y']) and np.array_equal(a['ktot'], b['ktot'])
    assert all(np.array_equal(a['y'][0], a['y'][k]) for k in range(5))


# ---------------------------------------------------------------- EOS / Fermi-Dirac / screening pins
def _olib():
    import ctypes as C
    from oracle import xnet_oracle as O
    l = O.lib('strict')
    for f in ('ox_fdm3h', 'ox_fdm1h', 'ox_fd1h'):
        getattr(l, f).restype = C.c_double
        getattr(l, f).argtypes = [C.c_double]
    return l


@pytest.mark.parametrize('k,name', [(0, 'ox_fd0h'), (1, 'ox_fd2h'), (2, 'ox_fd4h'), (3, 'ox_fd6h'), (4, 'ox_fd8h')])
def test_integer_order_fermi_dirac_against_quadrature(k, name):
    """Integer orders k = 0..4 used by the logft phase space (both intervals and the reflection branch): 1e-13."""
    import ctypes as C
    from scipy import integrate
    from oracle import xnet_oracle as O
    fn = getattr(O.lib('strict'), name)
    fn.restype = C.c_double; fn.argtypes = [C.c_double]
    for x in [-40, -10, -2.5, -2, -1, -0.1, 0, 0.3, 1, 2, 2.5, 5, 10, 30, 100]:
        h = lambda t: t ** k / (np.exp(min(t - x, 700.0)) + 1.0)
        q = integrate.quad(h, 0, max(x, 0) + 80, limit=500, epsabs=0, epsrel=1e-13, points=[x] if x > 0 else None)[0]
        assert abs(fn(x) / q - 1) < 1e-13


@pytest.mark.parametrize('x', [-30, -5, -2.0001, -2, -1, -1e-3, 0, 0.5, 1.999, 2, 3.3, 5, 7, 10, 13, 20, 31, 40, 55, 200])
def test_fermi_dirac_against_quadrature(x):
    """Pins the restated Fukushima approximations (every interval of every order) to the DEFINITION of the
    integrals, F_k(x) = int_0^inf t^k/(exp(t-x)+1) dt, by adaptive quadrature: 1e-13 relative for k=+-1/2;
    F_{-3/2} through d/dx F_{-1/2} = -1/2 F_{-3/2} (central difference of the quadrature, 1e-6)."""
    from scipy import integrate
    l = _olib()

    def fd(k, xx):   # t = u^2 removes the square-root singularity
        h = lambda u: 2 * u ** (2 * k + 1) / (np.exp(min(u * u - xx, 700.0)) + 1.0)
        return integrate.quad(h, 0, np.sqrt(max(xx, 0) + 60), limit=400, epsabs=0, epsrel=1e-13)[0]
    assert abs(l.ox_fd1h(x) / fd(0.5, x) - 1) < 1e-13
    assert abs(l.ox_fdm1h(x) / fd(-0.5, x) - 1) < 1e-13
    h = 1e-4 * max(1.0, abs(x))
    d = (fd(-0.5, x + h) - fd(-0.5, x - h)) / (2 * h)
    assert abs(l.ox_fdm3h(x) / (-2 * d) - 1) < 1e-5


def test_helmholtz_eta_and_cv_consistency():
    """Helmholtz pieces: d(eta)/dT9 returned by the bicubic interpolation equals the finite difference of
    eta itself (pins the derivative weights), and cv reproduces the ideal-gas + radiation limit at low
    density (ions 3/2 k/(A m_u), fully ionised electrons 3/2 k Ye/m_u, radiation 4aT^3/rho) to 1 %."""
    import ctypes as C
    from oracle import xnet_oracle as O
    p = O.Problem.from_control(control_text('test_settings_small', 'setup_tnsn_alpha', 1), DATA)
    l = O.lib('strict')
    l.ox_eos_interface.restype = C.c_int
    c_dp = C.POINTER(C.c_double)
    ny = p.nuc['ny']
    y = np.zeros(ny); y[0] = 1.0 / p.nuc['aa'][0]   # pure 4He

    def eos(t9, rho):
        out = [C.c_double() for _ in range(4)]
        rc = l.ox_eos_interface(C.byref(p.cnet), C.byref(p.cctl), C.c_double(t9), C.c_double(rho), y.ctypes.data_as(c_dp),
                                C.c_double(0.0), C.c_double(1.0), C.c_double(0.0), *[C.byref(o) for o in out])
        assert rc == 0
        return [o.value for o in out]   # ye, cv [MeV/GK/nucleon], eta, detadt9
    for t9, rho in [(0.01, 1.0), (0.5, 1e5), (1.0, 1e7), (3.0, 1e9), (4.7, 2e9)]:
        ye, cv, eta, deta = eos(t9, rho)
        h = 1e-6 * t9
        fd = (eos(t9 + h, rho)[2] - eos(t9 - h, rho)[2]) / (2 * h)
        assert ye == 0.5 and abs(deta / fd - 1) < 1e-6
    amu = 1.0 / (6.022140857e+23 * 1.6021766208e-06)
    cv_cgs = eos(0.01, 1.0)[1] / (amu * 1e9)
    ideal = 1.5 * 1.380650424e-16 * 6.0221417930e23 * (0.25 + 0.5) + 4 * 7.5657e-15 * 1e21
    assert abs(cv_cgs / ideal - 1) < 1e-2


def test_screening_enhances_charged_particle_rates_only():
    """ox_screening: factors are zero for reactions with a neutral reactant pair (zeta_w = 0), positive
    otherwise, grow with density, and the weak-screening limit h = ztilde*lambda0*Z1*Z2 holds at low density."""
    import ctypes as C
    from oracle import xnet_oracle as O
    p = O.Problem.from_control(control_text('test_settings', 'setup_tnsn_torch47', 1), DATA)
    l = O.lib('strict')
    ny = p.nuc['ny']; nr = p.net['nreac']
    y = np.zeros(ny); y[:] = 1.0 / (ny * np.asarray(p.nuc['aa']))
    zz = np.asarray(p.nuc['zz'])
    n2 = np.asarray(p.net['n2']).reshape(-1, 6)
    zw = zz[n2[:, 0] - 1] * zz[n2[:, 1] - 1]

    def h2(t9, rho):
        nrt = sum(nr); ocs = np.zeros(nrt); oyd = np.zeros(ny); oj = np.zeros(ny * ny); oye = np.zeros(1)
        assert l.ox_eval_point(C.byref(p.cnet), C.byref(p.cctl), t9, rho, O._pd(y), 0.0, 1.0, 0.0, 1.0, -1.0,
                               O._pd(ocs), O._pd(oyd), O._pd(oj), O._pd(oye)) == 0
        return ocs[nr[0]:nr[0] + nr[1]]
    p0 = O.Problem.from_control(control_text('test_settings', 'setup_tnsn_torch47'), DATA)

    def h2_off(t9, rho):
        nrt = sum(nr); ocs = np.zeros(nrt); oyd = np.zeros(ny); oj = np.zeros(ny * ny); oye = np.zeros(1)
        assert l.ox_eval_point(C.byref(p0.cnet), C.byref(p0.cctl), t9, rho, O._pd(y), 0.0, 1.0, 0.0, 1.0, -1.0,
                               O._pd(ocs), O._pd(oyd), O._pd(oj), O._pd(oye)) == 0
        return ocs[nr[0]:nr[0] + nr[1]]
    for t9, rho in [(1.0, 1e6), (2.0, 1e8), (0.5, 1e9)]:
        on, off = h2(t9, rho), h2_off(t9, rho)
        live = off > 0
        assert np.array_equal(on[live & (zw == 0)], off[live & (zw == 0)])
        assert (on[live & (zw > 0)] > off[live & (zw > 0)]).all()
    f_lo = (h2(1.0, 1e4) / np.where(h2_off(1.0, 1e4) > 0, h2_off(1.0, 1e4), 1))[(zw > 0) & (h2_off(1.0, 1e4) > 0)]
    f_hi = (h2(1.0, 1e8) / np.where(h2_off(1.0, 1e8) > 0, h2_off(1.0, 1e8), 1))[(zw > 0) & (h2_off(1.0, 1e8) > 0)]
    assert (f_hi > f_lo).all() and f_lo.max() < 1.25
    # weak-screening limit (lambda <= 0.1): ln(enhancement) = ztilde*lambda0 * Z1*Z2, one constant for all reactions
    sel = (zw > 0) & (h2_off(1.0, 1e4) > 0)
    k = np.log(f_lo) / zw[sel]
    k = k[np.log(f_lo) < 0.09]
    assert k.size > 10 and np.ptp(k) / k.mean() < 1e-9
