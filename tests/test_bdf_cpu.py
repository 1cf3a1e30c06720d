"""BDF integrator, CPU side: the oracle's restatement of solve_bdf (oracle/xnet_oracle.c, from
source/xnet_integrate_bdf.F90) on the reference's own BDF test problem (test/test_xnet.sh test 54: test_settings_bdf +
setup_bdf_sn160: SN160, self-heating, screening, six C/O-burn trajectories, iconvc = 3).

Parity status: UNPINNED like every floating-point trajectory of this repository (no Fortran compiler, no shipped
results).  What can be checked without the reference binary: the run completes on every trajectory inside the step
limit, conserves mass, needs fewer steps than first-order backward Euler, ends in the same physical state as the
backward-Euler run of the same problem (test_settings_heat: same trajectories, same physics) to the accuracy of the
first-order method, and the three differently compiled oracle builds agree."""
import numpy as np
import pytest

from conftest import DATA, PROBLEMS, control_text


@pytest.fixture(scope='module')
def runs(oracle_problem):
    bdf = oracle_problem('sn160_bdf').run(nzone=6, hist_cap=3000)
    be = oracle_problem('sn160_heat').run(nzone=6, hist_cap=3000)
    return bdf, be


def test_bdf_controls(oracle_problem):
    c = oracle_problem('sn160_bdf').ctl
    # xnet_controls.F90:270-273: a higher-order solver does not need XNet's own limits on the step
    assert c['isolv'] == 3 and c['changemx'] == 1.0e10 and c['changemxt'] == 1.0e10 and c['iconvc'] == 3 and c['iheat'] == 1


def test_bdf_completes_and_conserves_mass(oracle_problem, runs):
    bdf, be = runs
    aa = np.asarray(oracle_problem('sn160_bdf').net_tables['aa']) if hasattr(oracle_problem('sn160_bdf'), 'net_tables') else None
    assert (bdf['status'] == -1).all() and (be['status'] == -1).all()
    assert (bdf['kstep'] < 0.6 * be['kstep']).all(), (bdf['kstep'], be['kstep'])       # higher order: far fewer steps
    assert (bdf['kstep'] > 50).all()
    # Newton iterations per step: the corrector converges in 1-3 iterations with the step's own Jacobian
    it = bdf['ktot'][:, 1] / bdf['ktot'][:, 0]
    assert (it > 1.0).all() and (it < 3.0).all()
    assert np.array_equal(bdf['ktot'][:, 2], bdf['ktot'][:, 0])     # one Jacobian build + factorisation per attempt (max_p_age = -1)


def test_bdf_agrees_with_backward_euler(runs):
    bdf, be = runs
    # final temperature of the self-heating burn and the abundant species
    assert np.abs(bdf['t9'] / be['t9'] - 1).max() < 2e-3
    for z in range(6):
        m = be['y'][z] > 1e-4
        assert np.abs(bdf['y'][z][m] / be['y'][z][m] - 1).max() < 0.1
    assert np.abs(bdf['t'] / be['t'] - 1).max() < 1e-12        # both stop at tstop


def test_bdf_oracle_builds_agree(oracle_problem):
    op = oracle_problem('sn160_bdf')
    a = op.run(nzone=6, hist_cap=3000, fast=False)
    for build in ('pm', True):
        b = op.run(nzone=6, hist_cap=3000, fast=build)
        assert (b['status'] == -1).all()
        assert np.abs(b['kstep'] - a['kstep']).max() <= 0.05 * a['kstep'].max()
        assert np.abs(b['t9'] / a['t9'] - 1).max() < 1e-4
        for z in range(6):
            m = a['y'][z] > 1e-6
            assert np.abs(b['y'][z][m] / a['y'][z][m] - 1).max() < 1e-2
