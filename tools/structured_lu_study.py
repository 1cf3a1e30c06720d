#!/usr/bin/env python
"""Feasibility numbers for a structured (bordered-band) LU of XNet's Jacobians.
CPU only, uses the oracle to produce Jacobians along a trajectory.  For each network it reports, for the natural
species order and for the order with the light particles (n, p, d, t, he3, he4) moved LAST:
  nnz of the assembled matrix, nnz of L+U after partial-pivoting LU (the fill), the flops actually needed when zeros
  are skipped, against the dense 2/3 N^3, and the profile (per-row first/last non-zero column) storage in doubles.
usage: python tools/structured_lu_study.py > profiles/rNN_structured_lu_study.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import DATA, PROBLEMS, control_text   # noqa: E402
from oracle import xnet_oracle as O                 # noqa: E402


def lu_stats(a, tau=1.0):
    """Right-looking LU with (threshold) partial pivoting that skips structural zeros; returns (nnz(L+U), flops)."""
    a = a.copy(); n = a.shape[0]; flops = 0
    rowp = np.arange(n)
    for k in range(n - 1):
        col = np.abs(a[k:, k])
        p = k + int(np.argmax(col))
        if tau * col[p - k] <= col[0] and col[0] > 0:
            p = k
        if p != k:
            a[[k, p]] = a[[p, k]]; rowp[[k, p]] = rowp[[p, k]]
        if a[k, k] == 0.0:
            continue
        rows = np.nonzero(a[k + 1:, k])[0] + k + 1
        cols = np.nonzero(a[k, k + 1:])[0] + k + 1
        a[rows, k] /= a[k, k]
        flops += len(rows)
        if len(rows) and len(cols):
            a[np.ix_(rows, cols)] -= np.outer(a[rows, k], a[k, cols])
            flops += 2 * len(rows) * len(cols)
    return int(np.count_nonzero(a)), flops, a, rowp


def profile_storage(lu):
    n = lu.shape[0]; tot = 0
    for i in range(n):
        nzr = np.nonzero(lu[i])[0]; nzc = np.nonzero(lu[:, i])[0]
        lo = min(nzr.min() if len(nzr) else i, i); up = min(nzc.min() if len(nzc) else i, i)
        tot += (i - lo) + (i - up) + 1                  # skyline: row part of L, column part of U, diagonal
    return tot


def main():
    light = {'n', 'p', 'd', 't', 'he3', 'he4'}
    for name in ('torch47', 'sn160', 'xrb304'):
        op = O.Problem.from_control(control_text(*PROBLEMS[name]), DATA)
        ny = op.nuc['ny']; names = [s.strip().lower() for s in op.nuc['names']]
        h = op.run(nzone=1, hist_cap=4000)
        n = int(h['hist_n'][0])
        perm = [i for i in range(ny) if names[i] not in light] + [i for i in range(ny) if names[i] in light]
        print('== %s  N = %d  dense LU = %.3g flops, %d doubles' % (name, ny, 2.0 / 3.0 * ny ** 3, ny * ny))
        for k in np.linspace(n // 4, n - 2, 3).astype(int):
            y = h['hist_y'][0, k]; t9 = h['hist_t9'][0, k]; rho = h['hist_rho'][0, k]; rdt = 1.0 / h['hist_tdel'][0, k]
            cs, yd, jac, ye = op.eval_point(t9, rho, y, rdt, -1.0)          # jac[row, col]
            x_ref = np.linalg.solve(jac, yd)
            for label, pm in (('natural order', list(range(ny))), ('light particles last', perm)):
                a = jac[np.ix_(pm, pm)]
                for tau in (1.0, 0.25):
                    nnz, fl, lu, rowp = lu_stats(a, tau)
                    growth = np.abs(np.triu(lu)).max() / np.abs(a).max()
                    from scipy.linalg import solve_triangular
                    b = yd[pm][rowp]
                    xs = solve_triangular(np.triu(lu), solve_triangular(np.tril(lu, -1) + np.eye(ny), b, lower=True))
                    x = np.empty(ny); x[pm] = xs
                    big = np.abs(x_ref) > 1e-10 * np.abs(x_ref).max()
                    dsol = np.abs(x[big] / x_ref[big] - 1).max()
                    print('  step %4d T9 %.2f  %-22s tau %.2f: nnz(A) %5d (%.2f)  nnz(L+U) %6d (%.2f)  flops %.3g (%.3f of dense)  skyline %6d doubles (%.2f)  max|U|/max|A| %.2f  max rel. diff. of the solution vs dense DGESV %.1e'
                          % (k, t9, label, tau, np.count_nonzero(a), np.count_nonzero(a) / ny ** 2, nnz, nnz / ny ** 2, fl,
                             fl / (2.0 / 3.0 * ny ** 3), profile_storage(lu), profile_storage(lu) / ny ** 2, growth, dsol))
            del x_ref


if __name__ == '__main__':
    main()
