import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DATA = os.path.join(ROOT, 'data')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


SCRN_OFF = '0         Include Screening (yes=1)'
SCRN_ON = '1         Include Screening (yes=1)'


def control_text(settings, setup, screening=None):
    """control file = settings fragment + problem setup (test/test_xnet.sh:86-128); screening=1 flips the
    'Include Screening' switch of the settings fragment (the reference ships it switched on only in
    test_settings_heat)."""
    txt = open(os.path.join(DATA, settings)).read() + open(os.path.join(DATA, 'Test_Problems', setup)).read()
    if screening:
        assert SCRN_OFF in txt
        txt = txt.replace(SCRN_OFF, SCRN_ON)
    return txt


PROBLEMS = {
    'alpha': ('test_settings_small', 'setup_tnsn_alpha'),
    'torch47': ('test_settings', 'setup_tnsn_torch47'),
    'sn160': ('test_settings', 'setup_tnsn_sn160'),
    'xrb304': ('test_settings', 'setup_xrb_fisker'),
    'alpha_scrn': ('test_settings_small', 'setup_tnsn_alpha', 1),
    'torch47_scrn': ('test_settings', 'setup_tnsn_torch47', 1),
    'sn160_scrn': ('test_settings', 'setup_tnsn_sn160', 1),
    # self-heating + screening (test_settings_heat; test/test_xnet.sh tests 50-52)
    'alpha_heat': ('test_settings_heat', 'setup_heat_alpha'),
    'torch47_heat': ('test_settings_heat', 'setup_heat_torch47'),
    'sn160_heat': ('test_settings_heat', 'setup_heat_sn160'),
    # logft weak rates + self-heating + screening on SN231 (test/test_xnet.sh test 53; BASELINE configs[3])
    'sn231_logft': ('test_settings_logft', 'setup_logft_sn231'),
    # 16 zones with their own trajectory / abundance files, self-heating + screening, zone batching
    # (test/test_xnet.sh tests 61/62: test_settings_batch + setup_batch_torch47; BASELINE configs[1] batch variant)
    'torch47_batch': ('test_settings_batch', 'setup_batch_torch47'),
    # variable-order BDF integrator, self-heating + screening, six C/O-burn trajectories (test/test_xnet.sh test 54)
    'sn160_bdf': ('test_settings_bdf', 'setup_bdf_sn160'),
}


@pytest.fixture(scope='session')
def built():
    import __graft_entry__
    __graft_entry__.build()
    return True


@pytest.fixture(scope='session')
def oracle_problem(built):
    from oracle import xnet_oracle as O
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = O.Problem.from_control(control_text(*PROBLEMS[name]), DATA)
        return cache[name]
    return get


@pytest.fixture(scope='session')
def gpu_problem(built):
    """(controls, Network, fixtures) of the product side, loaded through the C ABI."""
    import xnet_b200 as X
    cache = {}

    def get(name):
        if name not in cache:
            ctl = X.read_controls(control_text(*PROBLEMS[name]))
            net = X.Network(os.path.join(DATA, ctl['data_dir']), ctl['iweak0'])
            fx = [X.load_fixture(net, os.path.join(DATA, a), os.path.join(DATA, t))
                  for a, t in list(zip(ctl['inab_file'], ctl['thermo_file']))[:1]]
            cache[name] = (ctl, net, fx)
        return cache[name]
    return get


@pytest.fixture(scope='session')
def gpu_problem_all(built):
    """As gpu_problem, with EVERY fixture of the control file (multi-zone set-ups)."""
    import xnet_b200 as X
    cache = {}

    def get(name):
        if name not in cache:
            ctl = X.read_controls(control_text(*PROBLEMS[name]))
            net = X.Network(os.path.join(DATA, ctl['data_dir']), ctl['iweak0'])
            fx = [X.load_fixture(net, os.path.join(DATA, a), os.path.join(DATA, t))
                  for a, t in zip(ctl['inab_file'], ctl['thermo_file'])]
            cache[name] = (ctl, net, fx)
        return cache[name]
    return get


def fortran_tables(n):
    """The tables of a loaded Network re-laid out as XNet's Fortran modules hold them (what the Fortran shim hands to
    xb_net_from_arrays): n1i(5,nr1), la/le(4,ny) class index fastest, FFN tables (nffn,143) rate index fastest, and the
    reaction pairs of Module xnet_match."""
    import numpy as np
    ny = n.ny
    t = dict(ny=ny, nffn=n.nffn, nnnu=n.nnnu, nreac=n.nreac, nan=n.nan, mex_n=n.dtable('mex_np')[0], mex_p=n.dtable('mex_np')[1])
    for c in range(1, 5):
        ni = n.itable('n%di' % c).reshape(-1, 6)
        t['n%di' % c] = np.ascontiguousarray(ni[:, :5]) if c == 1 else ni
        for k in ('rc', 'iwk', 'irev', 'mu', 'a', 'ifl'):
            t['%s%d' % (k, c)] = n.itable('%s%d' % (k, c)) if k in ('iwk', 'irev', 'mu', 'ifl') else n.dtable('%s%d' % (k, c))
    t['la'] = n.itable('la').reshape(4, ny).T.copy(); t['le'] = n.itable('le').reshape(4, ny).T.copy()
    for k in ('aa', 'zz', 'mex', 'be', 'g', 't9i', 'ffn_qval'):
        t[k] = n.dtable(k)
    nf = max(n.nffn, 1)
    for k in ('ffnsum', 'ffnenu', 'ffn_ec', 'ffn_beta'):
        t[k] = n.dtable(k).reshape(nf, 143)[:n.nffn].T.copy()
    t['has_logft'] = n.itable('has_logft')
    t.update(mflx=n.mflx, nflx=n.itable('nflx'), iwflx=n.itable('iwflx'), qflx=n.dtable('qflx'))
    return t
