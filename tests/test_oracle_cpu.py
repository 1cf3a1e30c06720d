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
    for name, cap in (('alpha', 3000), ('torch47', 400)):
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
