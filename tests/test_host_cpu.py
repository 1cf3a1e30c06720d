"""CPU tests of the product's host side: the C ABI library loads and exports every declared symbol,
its C++ loader/pre-processor agrees bit for bit with the oracle's independent parser and with the
reference's net_desc goldens.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import DATA, ROOT, control_text, PROBLEMS

NETS = [('alpha', 1), ('torch47', 1), ('SN160', 1), ('SN231', 1), ('SN231_logft', 2), ('XRB304', 1)]


def test_library_exports_every_declared_symbol(built):
    import xnet_b200 as X
    from xnet_b200 import api
    hdr = open(os.path.join(ROOT, 'include', 'xnet_b200.h')).read()
    declared = set(re.findall(r'\b(xb_[a-z0-9_]+)\s*\(', hdr))
    declared -= {'xb_step_be'} if 'int xb_step_be' not in hdr else set()
    lib = C.CDLL(X.lib_path())
    for name in sorted(declared):
        assert hasattr(lib, name), 'missing export ' + name
    assert declared == set(api.EXPORTS)
    assert X.lib().xb_version() == 1


@pytest.mark.parametrize('net,iweak0', NETS)
def test_loader_bit_exact_against_oracle(built, net, iweak0):
    import xnet_b200 as X
    from oracle import xnet_loader as L
    d = os.path.join(DATA, 'Data_' + net)
    n = X.Network(d, iweak0)
    nuc, on, ffn = L.load_network(d, iweak0)
    assert (n.ny, n.nreac, n.nan, n.nffn, n.nnnu) == (nuc['ny'], on['nreac'], on['nan'], on['nffn'], on['nnnu'])
    assert np.array_equal(n.itable('la'), on['la'].ravel()) and np.array_equal(n.itable('le'), on['le'].ravel())
    assert np.array_equal(n.itable('chap_la'), on['chap_la']) and np.array_equal(n.itable('chap_le'), on['chap_le'])
    for c in range(1, 5):
        for key, ref in (('mu%d', 'mu%d'), ('iwk%d', 'iwk%d'), ('irev%d', 'irev%d'), ('ires%d', 'ires%d')):
            assert np.array_equal(n.itable(key % c), on[ref % c]), key % c
        assert np.array_equal(n.itable('n%di' % c), on['n%d' % c].ravel())
        assert np.array_equal(n.dtable('a%d' % c), on['a%d' % c])
        assert np.array_equal(n.dtable('rc%d' % c), on['rc%d' % c].ravel())
        assert np.array_equal(n.dtable('q%d' % c), on['q%d' % c])
    assert np.array_equal(n.itable('iffn'), on['iffn'])
    for nm in ('aa', 'zz', 'mex', 'be', 't9i'):
        assert np.array_equal(n.dtable(nm), nuc[nm])
    assert np.array_equal(n.dtable('g'), nuc['g'].ravel())
    for nm in ('ffnsum', 'ffnenu', 'ffn_ec', 'ffn_beta'):
        assert np.array_equal(n.dtable(nm), ffn[nm].ravel())


@pytest.mark.parametrize('net', ['alpha', 'torch47', 'SN160', 'SN231', 'XRB304'])
def test_loader_against_net_desc(built, net):
    import xnet_b200 as X
    n = X.Network(os.path.join(DATA, 'Data_' + net))
    ref = open(os.path.join(ROOT, 'tests', 'golden', 'net_desc_%s.txt' % net)).read().split('\n')
    assert int(ref[1].split('=')[1]) == n.ny
    la, le = n.itable('chap_la'), n.itable('chap_le')
    for j in range(11):
        assert [int(v) for v in ref[3 + j].split()] == [j + 1, la[j], le[j]]
    assert [int(v) for v in ref[15].split('=')[1].split()] == n.nreac
    assert [int(v) for v in ref[16].split('=')[1].split()] == n.nan


@pytest.mark.parametrize('name', list(PROBLEMS))
def test_fixture_and_controls_against_oracle(built, name):
    import xnet_b200 as X
    from oracle import xnet_loader as L
    txt = control_text(*PROBLEMS[name])
    c, oc = X.read_controls(txt), L.read_controls(txt)
    for k in c:
        if k in oc:
            assert c[k] == oc[k], k
    net = X.Network(os.path.join(DATA, c['data_dir']), c['iweak0'])
    nuc = L.read_nuclear_data(os.path.join(DATA, c['data_dir']))
    f = X.load_fixture(net, os.path.join(DATA, c['inab_file'][0]), os.path.join(DATA, c['thermo_file'][0]))
    z = L.load_zone(os.path.join(DATA, c['inab_file'][0]), os.path.join(DATA, c['thermo_file'][0]), nuc, c['t9nse'])
    for k in ('nh', 'tstart', 'tstop', 'tdelstart', 'nstart', 't9start', 'rhostart', 'yestart', 'xext', 'aext', 'zext'):
        assert f[k] == z[k], k
    for k in ('th', 't9h', 'rhoh', 'ystart'):
        assert np.array_equal(f[k], z[k]), k


def test_missing_input_fails_loudly(built):
    import xnet_b200 as X
    with pytest.raises(X.XnetB200Error):
        X.Network(os.path.join(DATA, 'Data_does_not_exist'))


def test_no_cpu_fallback_without_gpu(built):
    """On a box without a CUDA device every compute entry point must fail, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    import xnet_b200 as X
    c = X.read_controls(control_text(*PROBLEMS['alpha']))
    net = X.Network(os.path.join(DATA, c['data_dir']))
    with pytest.raises(X.XnetB200Error):
        X.Context(net, c, nzevolve=1)


def test_product_never_imports_oracle():
    for dp, _, files in os.walk(os.path.join(ROOT, 'xnet_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.cpp', '.h')):
                s = open(os.path.join(dp, f)).read()
                assert 'oracle' not in s.replace('the oracle', '').replace("oracle's", '').replace('CPU oracle', ''), f


@pytest.mark.parametrize('net,iweak0', [('torch47', 1), ('SN160', 1), ('SN231_logft', 2)])
def test_network_from_fortran_arrays_round_trip(built, net, iweak0):
    """xb_net_from_arrays (the entry the Fortran shim uses: tables in Fortran layout, including the reaction pairs of
    Module xnet_match) reproduces the network that xb_net_load builds from the files, table by table; the
    double-counting factors are formed by the library when the caller passes none."""
    import xnet_b200 as X
    d = os.path.join(DATA, 'Data_' + net)
    from conftest import fortran_tables
    n = X.Network(d, iweak0)
    t = fortran_tables(n)
    m = X.Network.from_fortran_arrays(t, iweak0)
    assert (m.ny, m.nreac, m.nan, m.nffn, m.mflx) == (n.ny, n.nreac, n.nan, n.nffn, n.mflx)
    for key in ['la', 'le', 'iffn', 'has_logft', 'nflx', 'iwflx', 'ifl_orig', 'ifl_term'] + \
            ['%s%d' % (k, c) for k in ('mu', 'iwk', 'irev', 'ifl') for c in range(1, 5)] + ['n%di' % c for c in range(1, 5)]:
        assert np.array_equal(m.itable(key), n.itable(key)), key
    for key in ['aa', 'zz', 'mex', 'be', 'g', 't9i', 'ffnsum', 'ffnenu', 'ffn_ec', 'ffn_beta', 'ffn_qval', 'qflx', 'dcflx'] + \
            ['%s%d' % (k, c) for k in ('a', 'rc') for c in range(1, 5)]:
        assert np.array_equal(m.dtable(key), n.dtable(key)), key
    bad = dict(t); bad['ifl2'] = t['ifl2'].copy(); bad['ifl2'][0] = n.mflx + 5
    with pytest.raises(Exception):
        X.Network.from_fortran_arrays(bad, iweak0)


@pytest.mark.parametrize('name,net,frac', [('torch47', 'torch47', 0.60), ('sn160', 'SN160', 0.25), ('xrb304', 'XRB304', 0.20)])
def test_structured_lu_plan_envelope(built, oracle_problem, name, net, frac):
    """Groundwork for the structured LU: with the light particles ordered last, the skyline envelope the
    loader derives from the species <-> reaction maps (a) contains every non-zero of the assembled Jacobian, (b) is not
    left by an LU without interchanges (the numeric factors stay inside it), (c) is a small fraction of N^2."""
    import xnet_b200 as X
    n = X.Network(os.path.join(DATA, 'Data_' + net), 1)
    ny = n.ny
    perm = n.itable('slu_perm') - 1; rf = n.itable('slu_rowfirst') - 1; cf = n.itable('slu_colfirst') - 1
    assert np.array_equal(np.sort(perm), np.arange(ny))
    zz = n.dtable('zz'); aa = n.dtable('aa')
    light = (zz[perm] <= 2) & (aa[perm] <= 4)
    nl = int(light.sum())
    assert 3 <= nl <= 6 and light[-nl:].all() and not light[:-nl].any()
    assert (rf <= np.arange(ny)).all() and (cf <= np.arange(ny)).all()
    assert n.int('slu_doubles') == ny + int((np.arange(ny) - rf).sum() + (np.arange(ny) - cf).sum())
    assert n.int('slu_doubles') <= frac * ny * ny
    inside = np.zeros((ny, ny), bool)
    for i in range(ny):
        inside[i, rf[i]:i + 1] = True; inside[cf[i]:i + 1, i] = True
    op = oracle_problem(name)
    h = op.run(nzone=1, hist_cap=4000)
    k = int(h['hist_n'][0]) - 2
    cs, yd, jac, ye = op.eval_point(h['hist_t9'][0, k], h['hist_rho'][0, k], h['hist_y'][0, k], 1.0 / h['hist_tdel'][0, k], -1.0)
    a = jac[np.ix_(perm, perm)]                       # [row, col] in the structured order
    assert not (a != 0)[~inside].any()
    lu = a.copy()
    for c in range(ny - 1):                            # LU without interchanges
        assert lu[c, c] != 0.0
        lu[c + 1:, c] /= lu[c, c]
        lu[c + 1:, c + 1:] -= np.outer(lu[c + 1:, c], lu[c, c + 1:])
    assert not (lu != 0)[~inside].any()


def test_committed_bench_line_has_the_contract_keys():
    """The bench line committed under profiles/ (measured on a B200 this round) carries every key of the bench contract:
    headline metric, e2e with per-step copy bytes, launch count, clocks, the roofline object and the CPU baseline."""
    import glob
    import json
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_bench_default_*.json')))
    assert files
    d = json.load(open(files[-1]))
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline', 'cpu_baseline'):
        assert k in d, k
    assert d['metric'] == 'zone-Newton-iters/sec' and d['dtype'] == 'f64' and d['scaling'] == 'weak' and d['warmup'] >= 3
    assert 'workload' in d['config'] and d['config']['zones_per_gpu'] == 16384 and d['config'].get('all_zones_ok', d['config'].get('all_zones_finished'))
    assert set(('value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step')) <= set(d['e2e']) and d['e2e']['h2d_bytes_per_step'] > 0
    assert d['gpu_launches'] > 0 and not d['clocks']['reasons']
    r = d['roofline']
    assert set(('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic')) <= set(r) and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-12
    c = d['cpu_baseline']
    assert set(('value', 'unit', 'cores', 'kind', 'sample')) <= set(c) and c['kind'] in ('port', 'reference')
    assert d['value'] > 10 * c['value']
