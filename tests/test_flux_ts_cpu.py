"""CPU tests of the ts/flux row: reaction-pair matching (match_react + flux_init) of the product's
C++ loader against the oracle's independent restatement, structural invariants that pin both against the species
<-> reaction maps (which are pinned by the reference's net_desc goldens), the ts writer against an independent
byte-level restatement of the reference's record layout, and the reader / comparator.  No GPU calls."""
import os
import struct

import numpy as np
import pytest

from conftest import DATA, ROOT, control_text, PROBLEMS

NETS = [('alpha', 1), ('torch47', 1), ('SN160', 1), ('SN231', 1), ('SN231_logft', 2), ('XRB304', 1)]


@pytest.mark.parametrize('net,iweak0', NETS)
def test_match_react_bit_exact_against_oracle(built, net, iweak0):
    import xnet_b200 as X
    from oracle import xnet_loader as L
    d = os.path.join(DATA, 'Data_' + net)
    n = X.Network(d, iweak0)
    nuc, on, ffn = L.load_network(d, iweak0)
    assert n.mflx == on['mflx'] > 0
    assert np.array_equal(n.itable('nflx').reshape(-1, 8), on['nflx'])
    for c in range(1, 5):
        assert np.array_equal(n.itable('ifl%d' % c), on['ifl%d' % c])
    assert np.array_equal(n.itable('iwflx'), on['iwflx']) and np.array_equal(n.dtable('qflx'), on['qflx'])
    assert np.array_equal(n.dtable('dcflx').reshape(-1, 2), on['dcflx'])
    assert np.array_equal(n.itable('ifl_orig'), on['ifl_orig']) and np.array_equal(n.itable('ifl_term'), on['ifl_term'])


@pytest.mark.parametrize('net,iweak0', NETS)
def test_match_react_structure(built, net, iweak0):
    """Every reaction belongs to exactly one pair; a pair written in its exoergic direction holds the reaction's own
    nuclei (reactants | products for ifl > 0, swapped for ifl < 0); pairs opened by non-weak reactions are unique; baryon number and
    charge balance across every pair; origin/terminus are the last reactant/product."""
    import xnet_b200 as X
    d = os.path.join(DATA, 'Data_' + net)
    n = X.Network(d, iweak0)
    nflx = n.itable('nflx').reshape(-1, 8); iw = n.itable('iwflx')
    aa = n.dtable('aa'); zz = n.dtable('zz')
    nreactant = [1, 2, 3, 4]
    seen = np.zeros(n.mflx, int)
    for c in range(4):
        ni = n.itable('n%di' % (c + 1)).reshape(-1, 6); ifl = n.itable('ifl%d' % (c + 1)); q = n.dtable('q%d' % (c + 1))
        assert (ifl != 0).all() and (np.sign(ifl) == np.where(q > 0, 1, -1)).all()
        for k in range(len(ifl)):
            m = abs(ifl[k]) - 1; seen[m] += 1
            reac = [v for v in ni[k, :nreactant[c]]]; prod = [v for v in ni[k, nreactant[c]:] if v != 0]
            lhs = [v for v in nflx[m, :4] if v]; rhs = [v for v in nflx[m, 4:] if v]
            assert (lhs, rhs) == ((reac, prod) if ifl[k] > 0 else (prod, reac))
    assert (seen >= 1).all()    # (a weak pair may also collect a later non-weak reaction of the same nuclei: the search is over all sets)
    keys = [tuple(r) for r in nflx[iw == 0]]
    assert len(set(keys)) == len(keys)
    for m in range(n.mflx):
        lhs = [v - 1 for v in nflx[m, :4] if v]; rhs = [v - 1 for v in nflx[m, 4:] if v]
        assert abs(aa[lhs].sum() - aa[rhs].sum()) < 1e-9
        if iw[m] == 0:
            assert abs(zz[lhs].sum() - zz[rhs].sum()) < 1e-9
        assert n.itable('ifl_orig')[m] == nflx[m, len(lhs) - 1] and n.itable('ifl_term')[m] == nflx[m, 4 + len(rhs) - 1]


@pytest.mark.parametrize('name', ['alpha', 'torch47', 'sn160_scrn', 'xrb304'])
def test_flux_reproduces_ydot(oracle_problem, name):
    """flux() and yderiv() are two summations of the same reaction rates: ydot_i = sum over pairs of
    flx_m * (multiplicity of i among the products - among the reactants).  Holds to round-off and ties the flux
    restatement (matching, double-counting factors, signs) to the ydot maps pinned by net_desc."""
    op = oracle_problem(name)
    net = op.net; ny = op.nuc['ny']
    h = op.run(nzone=1, hist_cap=4000, flux=True)
    n = int(h['hist_n'][0])
    assert np.abs(h['hist_flx'][0, 0]).max() == 0.0            # initial record: flx = 0 (xnet_output.F90:60-66)
    S = np.zeros((ny, net['mflx']))
    for m in range(net['mflx']):
        for s in range(4):
            if net['nflx'][m, s]: S[net['nflx'][m, s] - 1, m] -= 1
            if net['nflx'][m, 4 + s]: S[net['nflx'][m, 4 + s] - 1, m] += 1
    # Reference quirk: for a direction with two PAIRS of identical particles (n n he4 he4, p p he4 he4) flux_init
    # uses 1/(2!2!) = 1/4 (fact(5), xnet_flux.F90:40-46) while the ydot coefficients a_n carry 1/2! (largest
    # multiplicity only, xnet_preprocess.F90:752-765).  Both are restated as they are; species touched by such a
    # pair are left out of the ydot identity and checked through the per-pair sums below instead.
    odd = np.zeros(ny, bool)
    for m in np.nonzero((net['dcflx'] == 0.25).any(axis=1))[0]:
        odd[[v - 1 for v in net['nflx'][m] if v]] = True
    assert not odd.all()
    for k in np.linspace(1, n - 1, 6).astype(int):
        y = h['hist_y'][0, k]; t9 = h['hist_t9'][0, k]; rho = h['hist_rho'][0, k]
        flx = op.flux_point(t9, rho, y)
        cs, yd, jac, ye = op.eval_point(t9, rho, y, 0.0, 1.0)
        # per-pair sums from the loader's arrays in numpy (independent of the C flux loop)
        want = np.zeros(net['mflx']); off0 = 0
        for c in range(4):
            ni = net['n%d' % (c + 1)]; ifl = net['ifl%d' % (c + 1)]
            for kk in range(len(ifl)):
                m = abs(ifl[kk]) - 1
                want[m] += np.sign(ifl[kk]) * net['dcflx'][m, 0 if ifl[kk] > 0 else 1] * cs[off0 + kk] * np.prod(y[ni[kk, :c + 1] - 1])
            off0 += len(ifl)
        # a net flux is the difference of a forward and a reverse term: the rounding bound uses the gross flux
        gross = np.zeros(net['mflx']); off = 0
        for c in range(4):
            ni = net['n%d' % (c + 1)]; ifl = net['ifl%d' % (c + 1)]
            for kk in range(len(ifl)):
                gross[abs(ifl[kk]) - 1] += cs[off + kk] * np.prod(y[ni[kk, :c + 1] - 1])
            off += len(ifl)
        assert (np.abs(flx) <= gross * (1 + 1e-12)).all()
        assert (np.abs(flx - want) <= 1e-14 * gross).all()
        ok = ~odd
        assert (np.abs(S @ flx - yd)[ok] <= 1e-13 * (np.abs(S) @ gross)[ok] + 1e-300).all()
        # the recorded flux of step k uses the rates of the step's own rate evaluation (start-of-step Ye in the
        # electron-capture rates, start-of-step composition in the screening factors), so it equals the flux re-evaluated at the accepted state only approximately
        assert (np.abs(h["hist_flx"][0, k] - flx) <= 1e-2 * gross + 1e-300).all()


def _golden_bytes(hdr, rec):
    """Independent restatement of the reference's Write statements with struct.pack (record markers included)."""
    def frec(payload):
        return struct.pack('<i', len(payload)) + payload + struct.pack('<i', len(payload))
    c80 = lambda s: s.encode()[:80].ljust(80)
    out = frec(b''.join(c80(s) for s in hdr['descript']) + c80(hdr['data_desc']))
    c = hdr['ctl']
    out += frec(struct.pack('<5i6d2i', c['kstmx'], c['kitmx'], c['iweak0'], c['iscrn'], c['iconvc'], c['changemx'], c['tolm'],
                            c['tolc'], c['yacc'], c['ymin'], c['tdel_maxmult'], c['iheat'], c['isolv']))
    out += frec(c80(hdr['inab_file']) + c80(hdr['abund_desc']))
    out += frec(c80(hdr['thermo_file']) + c80(hdr['thermo_desc']))
    ny = len(hdr['zz'])
    out += frec(struct.pack('<i', ny) + hdr['zz'].astype('<f8').tobytes() + hdr['aa'].astype('<f8').tobytes())
    mflx = len(hdr['ifl_orig'])
    out += frec(struct.pack('<i', mflx) + hdr['ifl_orig'].astype('<i4').tobytes() + hdr['ifl_term'].astype('<i4').tobytes())
    for k in range(len(rec['kstep'])):
        out += frec(struct.pack('<i6d', rec['kstep'][k], rec['t'][k], rec['t9'][k], rec['rho'][k], rec['sqnu'][k], rec['tdel'][k],
                                rec['edot'][k]) + rec['y'][k].astype('<f8').tobytes() + rec['flx'][k].astype('<f8').tobytes())
    return out


@pytest.mark.parametrize('name', ['alpha', 'torch47'])
def test_ts_writer_reader_round_trip(built, oracle_problem, tmp_path, name):
    """xb_ts_write (C++) produces byte for byte what the reference's Write statements produce for the same data
    (header records of net.F90:288-301, step records of xnet_output.F90:78-79), and read_ts returns it unchanged."""
    import xnet_b200 as X
    op = oracle_problem(name)
    ctl = X.read_controls(control_text(*PROBLEMS[name]))
    net = X.Network(os.path.join(DATA, ctl['data_dir']), ctl['iweak0'])
    h = op.run(nzone=1, hist_cap=4000, flux=True)      # the history to be written (any source of records will do)
    n = int(h['hist_n'][0])
    path = str(tmp_path / X.name_ordered('ts_' + name + '_', 1, 10))
    assert path.endswith('_01')
    desc = ('first line', 'second', 'x' * 100)
    nw = X.write_ts(path, net, ctl, None, h, zone=0, descript=desc, data_desc='net', inab_file=ctl['inab_file'][0],
                    abund_desc='ab', thermo_file=ctl['thermo_file'][0], thermo_desc='th')
    assert nw == n
    st = ctl.to_struct()
    hdr = dict(descript=desc, data_desc='net', ctl={k: getattr(st, k) for k in ('kstmx', 'kitmx', 'iweak0', 'iscrn', 'iconvc',
               'changemx', 'tolm', 'tolc', 'yacc', 'ymin', 'tdel_maxmult', 'iheat', 'isolv')},
               inab_file=ctl['inab_file'][0], abund_desc='ab', thermo_file=ctl['thermo_file'][0], thermo_desc='th',
               zz=net.dtable('zz'), aa=net.dtable('aa'), ifl_orig=net.itable('ifl_orig'), ifl_term=net.itable('ifl_term'))
    rec = {k: h['hist_' + k][0, :n] for k in ('kstep', 't', 't9', 'rho', 'sqnu', 'tdel', 'edot', 'y', 'flx')}
    assert open(path, 'rb').read() == _golden_bytes(hdr, rec)
    r = X.read_ts(path)
    assert r['descript'] == ['first line', 'second', 'x' * 80] and r['ny'] == net.ny and r['mflx'] == net.mflx
    assert r['controls']['kstmx'] == ctl['kstmx'] and r['controls']['tolc'] == ctl['tolc'] and r['controls']['isolv'] == 1
    for k in ('kstep', 't', 't9', 'rho', 'sqnu', 'tdel', 'edot', 'y', 'flx'):
        assert np.array_equal(r[k], rec[k]), k
    assert np.array_equal(r['xmf'], rec['y'] * net.dtable('aa'))
    assert abs(r['xmf'][-1].sum() - 1.0) < 1e-6                      # mass fractions still sum to one at the end
    # the reference's own reader slices five scalars per record: its view of the same file
    lg = X.read_ts(path, legacy=True)
    assert np.array_equal(lg['tdel'], rec['sqnu']) and np.array_equal(lg['edot'], rec['tdel'])


def test_compare_ts_semantics(built, oracle_problem, tmp_path):
    """compare_ts: identical files pass; a 1e-9 perturbation of one abundance fails the per-step 1e-10 test and is
    reported; a dropped step fails the accepted-step-sequence test; truncated files are rejected by the reader."""
    import xnet_b200 as X
    name = 'alpha'
    op = oracle_problem(name)
    ctl = X.read_controls(control_text(*PROBLEMS[name]))
    net = X.Network(os.path.join(DATA, ctl['data_dir']), ctl['iweak0'])
    h = op.run(nzone=1, hist_cap=4000, flux=True)
    a = str(tmp_path / 'a'); b = str(tmp_path / 'b'); c = str(tmp_path / 'c')
    X.write_ts(a, net, ctl, None, h)
    X.write_ts(b, net, ctl, None, h)
    r = X.compare_ts(a, b)
    assert r['ok'] and r['y_rel_max'] == 0.0 and r['final_rerr_norm'] == 0.0
    h2 = {k: v.copy() for k, v in h.items() if hasattr(v, 'copy')}
    big = int(np.argmax(h2['hist_y'][0, 100]))
    h2['hist_y'][0, 100, big] *= 1 + 1e-9
    X.write_ts(b, net, ctl, None, h2)
    r = X.compare_ts(a, b)
    assert not r['ok'] and 5e-10 < r['y_rel_max'] < 2e-9 and r['kstep_equal']
    assert X.compare_ts(a, b, rtol=1e-8)['ok']
    h3 = {k: v.copy() for k, v in h.items() if hasattr(v, 'copy')}
    n = int(h3['hist_n'][0])
    for k in h3:
        if k.startswith('hist_') and k != 'hist_n':
            h3[k][0, 50:n - 1] = h3[k][0, 51:n]
    h3['hist_n'][0] = n - 1
    X.write_ts(c, net, ctl, None, h3)
    r = X.compare_ts(a, c)
    assert not r['ok'] and not r['same_step_count']
    assert X.compare_ts(a, c, steps=False)['ok']                      # final-state comparison only (compare_final)
    raw = open(a, 'rb').read()
    open(c, 'wb').write(raw[:-5])
    with pytest.raises(ValueError):
        X.read_ts(c)


def test_ev_writer_fortran_formats(built, oracle_problem, tmp_path):
    """xb_ev_write: header (a4,a15,5a10,<n>a9,a4) and lines (i4,1es15.8,2es10.3,3es10.2,<n>es9.2,2i2) as Fortran prints
    them: ES editing with one digit before the point, three-digit exponents without the letter E, asterisks on
    overflow of the field, right-justified character fields."""
    import xnet_b200 as X
    name = 'alpha'
    op = oracle_problem(name)
    ctl = X.read_controls(control_text(*PROBLEMS[name]))
    net = X.Network(os.path.join(DATA, ctl['data_dir']), ctl['iweak0'])
    h = op.run(nzone=1, hist_cap=4000)
    # hand-made values in the first records to pin the formats
    h['hist_kstep'][0, :4] = [0, 7, 9999, 10000]
    h['hist_t'][0, 1] = 1.23456789123e-5; h['hist_t9'][0, 1] = 9.9996; h['hist_rho'][0, 1] = 1.0e7
    h['hist_edot'][0, 1] = -3.14159e17; h['hist_sqnu'][0, 1] = 0.0; h['hist_tdel'][0, 1] = 2.5e-112
    h['hist_kmon'][0, 1] = [1, 3]
    h['hist_y'][0, 1, :] = 0.0
    names = net.names()
    i_he4 = [s.strip() for s in names].index('he4'); i_c12 = [s.strip() for s in names].index('c12')
    h['hist_y'][0, 1, i_he4] = 0.25 / 4.0; h['hist_y'][0, 1, i_c12] = 1.0e-120 / 12.0
    path = str(tmp_path / 'ev_1')
    n = X.write_ev(path, net, h, zone=0, names=['he4', 'c12'])
    lines = open(path).read().split('\n')
    assert len(lines) == n + 2 and lines[-1] == ''
    assert lines[0] == '  k ' + '          Time ' + '    T(GK) ' + ' Density '.rjust(10) + '    dE/dt ' + '     sqnu ' + ' Timestep ' + \
        '      he4' + '      c12' + ' It '
    assert lines[2] == '   7' + ' 1.23456789E-05' + ' 1.000E+01' + ' 1.000E+07' + ' -3.14E+17' + '  0.00E+00' + ' 2.50-112' .rjust(10) + \
        ' 2.50E-01' + ' 1.00-120' .rjust(9) + ' 1' + ' 3'
    assert lines[3].startswith('9999') and lines[4].startswith('****')
    w = 4 + 15 + 2 * 10 + 3 * 10 + 2 * 9 + 2 * 2
    assert all(len(ln) == w for ln in lines[:-1])
