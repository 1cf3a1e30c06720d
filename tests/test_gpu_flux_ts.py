"""GPU parity tests of the ts/flux row: reaction fluxes computed on the device (fused history and
module-level xb_flux) against the oracle's flux(), and the ts file written from a GPU run against the ts file written
from the oracle's run of the same problem.  Strict arithmetic: bit for bit (files byte-identical).  Fast arithmetic:
fluxes to 1e-10 relative of the pair's gross flux (a net flux is a difference of a forward and a reverse term)."""
import os

import numpy as np
import pytest

from conftest import DATA, PROBLEMS, control_text

pytestmark = pytest.mark.gpu


def _ctx(gpu_problem, name, nz=1, arith=0, **ov):
    import xnet_b200 as X
    ctl, net, fx = gpu_problem(name)
    return X, ctl, net, fx, X.Context(net, ctl, nzevolve=nz, arith=arith, **ov)


@pytest.mark.parametrize('name,cap', [('alpha', 3000), ('torch47', 3100), ('sn160_scrn', 3200), ('sn160_heat', 1200)])
def test_strict_flux_history_and_ts_file_bit_exact(gpu_problem, oracle_problem, tmp_path, name, cap):
    X, ctl, net, fx, ctx = _ctx(gpu_problem, name, 2, arith=1)
    res = ctx.full_net(X.build_zone_set(fx, nzone=2), hist_cap=cap, flux=True)
    ref = oracle_problem(name).run(nzone=1, hist_cap=cap, fast='pm', flux=True)
    n = int(ref['hist_n'][0])
    assert n <= cap and int(res['hist_n'][0]) == n and int(res['hist_n'][1]) == n
    assert np.array_equal(res['hist_y'][0, :n], ref['hist_y'][0, :n])
    assert np.array_equal(res['hist_flx'][0, :n], ref['hist_flx'][0, :n])
    assert np.array_equal(res['hist_flx'][1, :n], ref['hist_flx'][0, :n])       # replicated zone: identical
    assert np.abs(res['hist_flx'][0, 1:n]).max() > 0 and np.abs(res['hist_flx'][0, 0]).max() == 0
    a = str(tmp_path / 'ts_gpu'); b = str(tmp_path / 'ts_oracle')
    X.write_ts(a, net, ctl, None, res, zone=1, thermo_file=ctl['thermo_file'][0], inab_file=ctl['inab_file'][0])
    X.write_ts(b, net, ctl, None, ref, zone=0, thermo_file=ctl['thermo_file'][0], inab_file=ctl['inab_file'][0])
    assert open(a, 'rb').read() == open(b, 'rb').read()
    r = X.compare_ts(a, b, rtol=0.0)
    assert r['ok'] and r['nstep1'] == n
    ctx.close()


@pytest.mark.parametrize('name', ['torch47', 'sn160_scrn', 'xrb304'])
@pytest.mark.parametrize('arith', [1, 0])
def test_module_level_flux(gpu_problem, oracle_problem, name, arith):
    """xb_flux after xb_cross_sect at states sampled along the trajectory against the oracle's flux() of the same
    states; masked-out zones keep the caller's buffer."""
    op = oracle_problem(name)
    h = op.run(nzone=1, hist_cap=4000)
    n = int(h['hist_n'][0])
    idx = np.linspace(1, n - 2, 4).astype(int)
    y = h['hist_y'][0, idx]; t9 = h['hist_t9'][0, idx]; rho = h['hist_rho'][0, idx]
    X, ctl, net, fx, ctx = _ctx(gpu_problem, name, 4, arith=arith)
    ctx.cross_sect(y, t9, rho)
    flx = ctx.flux(y)
    nflx = net.itable('nflx').reshape(-1, 8)
    for z in range(4):
        o = op.flux_point(t9[z], rho[z], y[z], which='pm' if arith else 'strict')
        if arith:
            assert np.array_equal(flx[z], o)
        else:
            # gross flux of each pair from the oracle's rates: bound for the rounding of the net flux
            cs, yd, jac, ye = op.eval_point(t9[z], rho[z], y[z], 0.0, 1.0)
            gross = np.zeros(net.mflx); off = 0
            for c in range(4):
                ni = op.net['n%d' % (c + 1)]; ifl = op.net['ifl%d' % (c + 1)]
                for k in range(len(ifl)):
                    gross[abs(ifl[k]) - 1] += cs[off + k] * np.prod(y[z][ni[k, :c + 1] - 1])
                off += len(ifl)
            assert (np.abs(flx[z] - o) <= 1e-10 * gross + 1e-300).all()
    keep = np.full((4, net.mflx), -7.0)
    import ctypes as C
    from xnet_b200 import api
    m = np.array([1, 0, 1, 0], np.int32)
    api._chk(api.lib().xb_flux(ctx.h, 1, 4, api._pi(m), None, api._pd(keep)))
    assert np.array_equal(keep[[0, 2]], flx[[0, 2]]) and (keep[[1, 3]] == -7.0).all()
    ctx.close()


def test_fast_ts_file_against_oracle_prefix(gpu_problem, oracle_problem, tmp_path):
    """The north-star criterion through the ts files themselves: fast arithmetic on the GPU against the libm oracle,
    same accepted-step sequence and Y > 1e-20 within 1e-10 over the first 1000 records (torch47)."""
    name = 'torch47'
    X, ctl, net, fx, ctx = _ctx(gpu_problem, name, 1, arith=0)
    res = ctx.full_net(X.build_zone_set(fx, nzone=1), hist_cap=3100, flux=True)
    ref = oracle_problem(name).run(nzone=1, hist_cap=3100, flux=True)
    for h in (res, ref):
        h['hist_n'] = np.minimum(h['hist_n'], 1000)
    a = str(tmp_path / 'a'); b = str(tmp_path / 'b')
    X.write_ts(a, net, ctl, None, ref); X.write_ts(b, net, ctl, None, res)
    r = X.compare_ts(a, b, rtol=1e-10, floor=1e-20)
    assert r['ok'], {k: v for k, v in r.items() if not hasattr(v, 'shape')}
    ctx.close()


@pytest.mark.parametrize('name', ['torch47', 'sn160_heat', 'sn231_logft'])
def test_network_from_fortran_arrays_runs_identically(gpu_problem, name):
    """The Fortran shim's route: a network handed over as module arrays (xb_net_from_arrays) gives bit-identical
    trajectories and fluxes to the network the library loads from the files itself."""
    import xnet_b200 as X
    from conftest import fortran_tables
    ctl, net, fx = gpu_problem(name)
    net2 = X.Network.from_fortran_arrays(fortran_tables(net), ctl['iweak0'])
    net2.data_dir = net.data_dir            # where load_helm looks for the Helmholtz table
    out = []
    for n in (net, net2):
        ctx = X.Context(n, ctl, nzevolve=2, arith=1, kstmx=300)
        out.append(ctx.full_net(X.build_zone_set(fx, nzone=2), hist_cap=320, flux=True))
        ctx.close()
    a, b = out
    k = int(a['hist_n'][0])
    assert k == int(b['hist_n'][0]) and 10 < k <= 301
    for key in ('hist_y', 'hist_t', 'hist_tdel', 'hist_t9', 'hist_flx', 'hist_edot'):
        assert np.array_equal(a[key][:, :k], b[key][:, :k]), key          # (records beyond hist_n are not written)
    for key in ('y', 'ktot', 'hist_n'):
        assert np.array_equal(a[key], b[key]), key
