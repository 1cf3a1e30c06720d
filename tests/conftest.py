import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DATA = os.path.join(ROOT, 'data')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def control_text(settings, setup):
    return open(os.path.join(DATA, settings)).read() + open(os.path.join(DATA, 'Test_Problems', setup)).read()


PROBLEMS = {
    'alpha': ('test_settings_small', 'setup_tnsn_alpha'),
    'torch47': ('test_settings', 'setup_tnsn_torch47'),
    'sn160': ('test_settings', 'setup_tnsn_sn160'),
    'xrb304': ('test_settings', 'setup_xrb_fisker'),
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
