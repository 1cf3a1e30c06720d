#!/usr/bin/env python
"""How often would a structured LU (light particles last, no interchanges) have to fall back to pivoting?
Along complete oracle trajectories (every `stride`-th accepted step) the Jacobian is permuted, eliminated without
interchanges, and every column is tested with the fast mode's acceptance rule tau*max|a(i,k)| <= |a(k,k)|, tau = 0.25.
Reports the fraction of factorisations with no violated column, the violated columns per factorisation and where they are.
usage: python tools/structured_lu_accept.py [stride [tau]] > profiles/rNN_structured_lu_accept.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import DATA, PROBLEMS, control_text   # noqa: E402
from oracle import xnet_oracle as O                 # noqa: E402
import xnet_b200 as X                                # noqa: E402  (host-side loader only: the permutation)


def main():
    stride = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    tau = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
    for name, net in (('torch47', 'torch47'), ('sn160', 'SN160'), ('sn160_scrn', 'SN160'), ('xrb304', 'XRB304')):
        op = O.Problem.from_control(control_text(*PROBLEMS[name]), DATA)
        ny = op.nuc['ny']
        perm = X.Network(os.path.join(DATA, 'Data_' + net), 1).itable('slu_perm') - 1
        names = [op.nuc['names'][i].strip() for i in perm]
        h = op.run(nzone=1, hist_cap=4000)
        n = int(h['hist_n'][0])
        nfac = 0; clean = 0; viol = 0; where = {}; worst = 0.0
        for k in range(1, n - 1, stride):
            cs, yd, jac, ye = op.eval_point(h['hist_t9'][0, k], h['hist_rho'][0, k], h['hist_y'][0, k], 1.0 / h['hist_tdel'][0, k + 1], -1.0)
            a = jac[np.ix_(perm, perm)]
            bad = 0
            for c in range(ny - 1):
                col = np.abs(a[c:, c])
                if tau * col.max() > col[0]:
                    bad += 1; where[names[c]] = where.get(names[c], 0) + 1
                    worst = max(worst, col.max() / max(col[0], 1e-300))
                    p = c + int(np.argmax(col)); a[[c, p]] = a[[p, c]]          # pivot and carry on
                a[c + 1:, c] /= a[c, c]
                a[c + 1:, c + 1:] -= np.outer(a[c + 1:, c], a[c, c + 1:])
            nfac += 1; clean += bad == 0; viol += bad
        top = sorted(where.items(), key=lambda kv: -kv[1])[:8]
        print('%-11s N %3d  factorisations %4d  without any violated column %5.1f %%  violated columns per factorisation %.3f  '
              'worst |a_ik|/|a_kk| %.1f  most frequent: %s' % (name, ny, nfac, 100.0 * clean / nfac, viol / nfac, worst, top))


if __name__ == '__main__':
    main()
