"""Ad-hoc GPU parity probe (development aid): evolves small problems on the GPU and compares with the oracle."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import xnet_b200 as X
from oracle import xnet_oracle as O

ROOT = os.path.join(os.path.dirname(__file__), '..', 'data')

def problem(settings, setup):
    txt = open(os.path.join(ROOT, settings)).read() + open(os.path.join(ROOT, 'Test_Problems', setup)).read()
    ctl = X.read_controls(txt)
    net = X.Network(os.path.join(ROOT, ctl['data_dir']), ctl['iweak0'])
    fx = [X.load_fixture(net, os.path.join(ROOT, a), os.path.join(ROOT, t)) for a, t in zip(ctl['inab_file'], ctl['thermo_file'])]
    op = O.Problem.from_control(txt, ROOT)
    return ctl, net, fx, op

def compare(name, settings, setup, nzone, arith, hist_cap=3200):
    ctl, net, fx, op = problem(settings, setup)
    fx = fx[:1] if setup.startswith('setup_tnsn_alpha') else fx
    ctx = X.Context(net, ctl, nzevolve=1, arith=arith)
    print(name, 'arith', arith, ctx.launch_config, flush=True)
    zs = X.build_zone_set(fx, nzone=nzone)
    t0 = time.time(); r = ctx.full_net(zs, hist_cap=hist_cap); dt = time.time() - t0
    o = op.run(nzone=1, hist_cap=hist_cap, fast='pm' if arith else False)
    n = min(r['hist_n'][0], o['hist_n'][0])
    print('  gpu %.3fs status %s kstep %s ktot %s | oracle kstep %s ktot %s' % (dt, r['status'][:2], r['kstep'][:2], r['ktot'][0], o['kstep'], o['ktot'][0]))
    td = np.abs(r['hist_tdel'][0, :n] - o['hist_tdel'][0, :n]) / np.maximum(o['hist_tdel'][0, :n], 1e-300)
    yo = o['hist_y'][0, :n]; yg = r['hist_y'][0, :n]; m = yo > 1e-20
    rel = np.where(m, np.abs(yg - yo) / np.where(m, yo, 1), 0).max(axis=1)
    first_bad = np.argmax(rel > 1e-10) if (rel > 1e-10).any() else -1
    print('  tdel rel max %.3g; y rel max %.3g; first step with rel>1e-10: %d of %d; bit-identical y: %s, tdel: %s' % (
        td.max(), rel.max(), first_bad, n, np.array_equal(yg, yo), np.array_equal(r['hist_tdel'][0, :n], o['hist_tdel'][0, :n])))
    if nzone > 1:
        print('  all zones identical:', all(np.array_equal(r['y'][0], r['y'][k]) for k in range(nzone)))
    print('  phases', {k: '%.3g' % v for k, v in ctx.phase_cycles().items()})
    ctx.close()

if __name__ == '__main__':
    which = sys.argv[1:] or ['alpha', 'torch47']
    for arith in (1, 0):
        if 'alpha' in which: compare('alpha', 'test_settings_small', 'setup_tnsn_alpha', 4, arith)
        if 'torch47' in which: compare('torch47', 'test_settings', 'setup_tnsn_torch47', 2, arith)
        if 'sn160' in which: compare('sn160', 'test_settings', 'setup_tnsn_sn160', 2, arith)
        if 'xrb' in which: compare('xrb304', 'test_settings', 'setup_xrb_fisker', 1, arith)
