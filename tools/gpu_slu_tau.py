#!/usr/bin/env python
"""Diagnostic (GPU box): deviation of the fast-mode complete trajectory from the libm oracle for several thresholds of the
structured LU's column test.  usage: python tools/gpu_slu_tau.py xrb304 1e-3 1e-2 0.1 dense"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import __graft_entry__
__graft_entry__.build()
import xnet_b200 as X
from oracle import xnet_oracle as O
from conftest import DATA, PROBLEMS, control_text

name = sys.argv[1]
floor = float(os.environ.get('FLOOR', '1e-20'))
txt = control_text(*PROBLEMS[name])
op = O.Problem.from_control(txt, DATA)
ref = op.run(nzone=1, hist_cap=6000, fast=False)
alt = op.run(nzone=1, hist_cap=6000, fast=True)
ctl = X.read_controls(txt)
net = X.Network(os.path.join(DATA, ctl['data_dir']), ctl['iweak0'])
fx = [X.load_fixture(net, os.path.join(DATA, ctl['inab_file'][0]), os.path.join(DATA, ctl['thermo_file'][0]))]
names = net.names()


def dev(a, b, n):
    ya, yb = a['hist_y'][0, :n], b['hist_y'][0, :n]
    m = yb > floor
    return np.where(m, np.abs(ya - yb) / np.where(m, yb, 1.0), 0.0)


for v in sys.argv[2:]:
    if v == 'dense':
        ctx = X.Context(net, ctl, nzevolve=1, arith=0, solver=0)
    else:
        os.environ['XB_SLU_TAU'] = v
        ctx = X.Context(net, ctl, nzevolve=1, arith=0, solver=1)
    r = ctx.full_net(X.build_zone_set(fx, nzone=1), hist_cap=6000)
    ph = ctx.phase_cycles() if hasattr(ctx, 'phase_cycles') else None
    ctx.close()
    n = min(int(ref['hist_n'][0]), int(alt['hist_n'][0]), int(r['hist_n'][0]))
    dg, da = dev(r, ref, n), dev(alt, ref, n)
    env = np.maximum(np.maximum.accumulate(da.max(axis=1)), 1e-11)
    ratio = dg.max(axis=1) / env
    k = int(np.argmax(ratio)); sp = int(np.argmax(dg[k]))
    first = int(np.argmax(ratio > 100)) if (ratio > 100).any() else -1
    print('%s tau %s: records %d/%d  max dev %.3g  worst ratio %.3g at record %d species %s Y=%.3g dev=%.3g env=%.3g  first>100 at %d  ktot %s' % (
        name, v, int(r['hist_n'][0]), int(ref['hist_n'][0]), dg.max(), ratio[k], k, names[sp], ref['hist_y'][0, k, sp], dg[k, sp], env[k], first, r['ktot'][0].tolist()), flush=True)
    if first >= 0:
        for kk in range(first, min(first + 3, n)):
            top = np.argsort(-dg[kk])[:4]
            print('   record %d t9 %.4f: ' % (kk, ref['hist_t9'][0, kk]) + ', '.join('%s Y=%.2e dev=%.2e' % (names[s], ref['hist_y'][0, kk, s], dg[kk, s]) for s in top))
    if ph is not None: print('   phases', ph)
