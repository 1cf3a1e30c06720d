"""Structured LU, host side: the index tables of xb_loader.cpp drive a numpy restatement of the device
algorithm (xnet_b200/csrc/xb_slu.cuh: skyline store, elimination without interchanges over the heavy part with the
threshold test, partial pivoting inside the light-particle border, unscaled multipliers, forward/backward solves) on
Jacobians the oracle assembles along real trajectories.  Checks: every structural non-zero lies inside the envelope,
the elimination never leaves it, and the solution agrees with LAPACK's dense solve to the conditioning of the system.
CPU only: the CUDA kernels are checked against the same dense solves in tests/test_gpu_slu.py."""
import os

import numpy as np
import pytest

from conftest import DATA, PROBLEMS, control_text

TAU = 1.0e-3     # kSluTauDefault (xb_slu.cuh)


class Slu:
    """numpy mirror of slu_factor / slu_solve: same storage, same elimination order, same pivot rule in the border"""
    def __init__(self, net, heat):
        self.heat = heat
        t = {k: net.slu_table(k, heat) for k in ('perm', 'pos', 'rf', 'cf', 'cb', 'cbot', 'nrb', 'ncb', 'sizes')}
        self.n, self.nstore, self.nb = int(t['sizes'][0]), int(t['sizes'][1]), int(t['sizes'][2])
        self.bo, self.bs = int(t['sizes'][4]), int(t['sizes'][5])
        self.perm, self.pos, self.rf, self.cf, self.cb, self.cbot, self.nrb, self.ncb = (
            t[k].astype(int) for k in ('perm', 'pos', 'rf', 'cf', 'cb', 'cbot', 'nrb', 'ncb'))
        self.nh = self.n - self.nb

    def addr(self, i, j):
        return self.bo + j * self.bs + (i - self.nh) if i >= self.nh else self.cb[j] + i

    def inside(self, i, j):
        return i >= self.nh or self.cf[j] <= i <= self.cbot[j]

    def structural(self, i, j):
        """the row/column skyline of structured_lu_plan (a subset of the stored envelope)"""
        return self.cf[j] <= i if i <= j else self.rf[i] <= j

    def rows(self, k):
        """rows that take part in pivot k: the band span, then the border"""
        return list(range(k + 1, k + 1 + self.nrb[k])) + list(range(self.nh, self.n))

    def cols(self, k):
        c = list(range(k + 1, k + 1 + self.ncb[k])) + list(range(self.nh, self.n))
        return [j for j in c if self.cf[j] <= k]

    def scatter(self, a):
        """a[row, col] in equation order -> store; asserts nothing non-zero falls outside the structural pattern, that
        the pattern lies inside the stored envelope and that no two elements share an address"""
        S = np.zeros(self.nstore)
        used = np.zeros(self.nstore, bool)
        for r in range(self.n):
            for c in range(self.n):
                pi, pj = self.pos[r], self.pos[c]
                if self.inside(pi, pj):
                    ad = self.addr(pi, pj)
                    assert 0 <= ad < self.nstore and not used[ad]
                    used[ad] = True; S[ad] = a[r, c]
                if not self.structural(pi, pj):
                    assert a[r, c] == 0.0, 'non-zero outside the structural pattern at (%d,%d)' % (r, c)
                else:
                    assert self.inside(pi, pj), 'structural element (%d,%d) outside the stored envelope' % (pi, pj)
        return S

    def factor(self, S, b=None):
        n, nh, nb = self.n, self.nh, self.nb
        rinv = np.zeros(n); fail = False
        b = None if b is None else b.copy()
        ntest = n - self.heat
        for k in range(nh):
            piv = S[self.cb[k] + k]
            rinv[k] = 1.0 / piv
            rows = self.rows(k)
            l = np.array([S[self.addr(i, k)] for i in rows])
            if any(TAU * abs(v) > abs(piv) for v, i in zip(l, rows) if i < ntest):   # the T9 row is exempt (xb_slu.cuh)
                fail = True
            for j in self.cols(k):
                u = S[self.cb[j] + k] * rinv[k]
                for v, i in zip(l, rows):
                    S[self.addr(i, j)] -= v * u
            if b is not None:
                u = b[k] * rinv[k]
                for v, i in zip(l, rows):
                    b[i] -= v * u
        # border block: the diagonal is kept while it passes the threshold test against the tested rows below it
        v = np.array([[S[self.addr(nh + r, nh + c)] for c in range(nb)] for r in range(nb)])
        vb = None if b is None else b[nh:].copy()
        ipb = list(range(nb))
        nbt = ntest - nh
        for j in range(nb):
            col = np.abs(v[j:, j])
            p = j + int(np.argmax(col))
            bestt = col[:max(nbt - j, 0)].max() if nbt > j else -1.0
            if col[0] > 0.0 and not TAU * bestt > col[0]:
                p = j
            ipb[j] = p
            if p != j:
                v[[j, p]] = v[[p, j]]
                if vb is not None:
                    vb[[j, p]] = vb[[p, j]]
            rinv[nh + j] = 1.0 / v[j, j]
            for r in range(j + 1, nb):
                lr = v[r, j] * rinv[nh + j]
                v[r, j + 1:] -= lr * v[j, j + 1:]
                if vb is not None:
                    vb[r] -= lr * vb[j]
        for r in range(nb):
            for c in range(nb):
                S[self.addr(nh + r, nh + c)] = v[r, c]
        if b is not None:
            b[nh:] = vb
        for j in range(nb):
            if ipb[j] != j:
                ia = [self.addr(nh + j, c) for c in range(nh)]; ib = [self.addr(nh + ipb[j], c) for c in range(nh)]
                tmp = S[ia].copy(); S[ia] = S[ib]; S[ib] = tmp
        return rinv, ipb, fail, b

    def solve(self, S, rinv, ipb, b, fwd_done):
        n, nh, nb = self.n, self.nh, self.nb
        b = b.copy()
        if not fwd_done:
            for j in range(nb):
                p = ipb[j]
                if p != j:
                    b[[nh + j, nh + p]] = b[[nh + p, nh + j]]
            for k in range(n):
                f = b[k] * rinv[k]
                rows = self.rows(k) if k < nh else range(k + 1, n)
                for i in rows:
                    b[i] -= S[self.addr(i, k)] * f
        for k in range(n - 1, -1, -1):
            x = b[k] * rinv[k]
            b[k] = x
            for i in range(self.cf[k], k):
                b[i] -= S[self.addr(i, k)] * x
        return b


_HIST = {}


def _systems(op, name, heat, nsamp=5):
    """(a[row, col], rhs) in equation order at nsamp accepted steps of the oracle's trajectory"""
    from oracle import xnet_oracle as O
    if name not in _HIST:                              # the oracle's trajectory (XRB304: 35 s) is shared by the tests of this file
        _HIST[name] = op.run(nzone=1, hist_cap=6000)
    h = _HIST[name]
    n = int(h['hist_n'][0])
    out = []
    for k in np.linspace(2, n - 2, nsamp).astype(int):
        rdt = 1.0 / h['hist_tdel'][0, k + 1]
        if heat:
            r = O.eval_point_heat(op, h['hist_t9'][0, k], h['hist_rho'][0, k], h['hist_y'][0, k], rdt, -1.0)
            a = r['jac'].T.copy()                   # [row, col]
            rhs = np.concatenate([r['ydot'], [r['t9dot']]])
        else:
            cs, yd, jac, ye = op.eval_point(h['hist_t9'][0, k], h['hist_rho'][0, k], h['hist_y'][0, k], rdt, -1.0)
            a, rhs = jac, yd
        out.append((a, rhs))
    return out


@pytest.mark.parametrize('name', ['alpha', 'torch47', 'sn160_scrn', 'torch47_heat', 'sn160_heat', 'xrb304'])
def test_structured_lu_restatement(name, oracle_problem, gpu_problem):
    op = oracle_problem(name)
    ctl, net, fx = gpu_problem(name)
    heat = 1 if ctl['iheat'] > 0 else 0
    slu = Slu(net, heat)
    n = slu.n
    assert n == net.ny + heat and slu.nb >= 1 and slu.nstore < 65536
    assert sorted(slu.perm.tolist()) == list(range(n)) and (slu.pos[slu.perm] == np.arange(n)).all()
    if heat:
        assert slu.perm[n - 1] == net.ny           # T9 stays the last equation
    nfail = 0
    for a, rhs in _systems(op, name, heat, 3 if n > 200 else 5):
        S = slu.scatter(a)
        x_ref = np.linalg.solve(a, rhs)
        rinv, ipb, fail, bf = slu.factor(S.copy(), rhs[slu.perm])
        nfail += fail
        Sf = S.copy()
        rinv2, ipb2, fail2, _ = slu.factor(Sf)
        x1 = slu.solve(Sf, rinv2, ipb2, bf, True)[slu.pos]            # carried right-hand side
        x2 = slu.solve(Sf, rinv2, ipb2, rhs[slu.perm], False)[slu.pos]   # separate forward solve
        scale = np.abs(x_ref).max()
        cond = np.linalg.cond(a)
        tol = max(1e-13 * cond, 1e-12) * scale
        assert np.abs(x1 - x_ref).max() <= tol, (name, np.abs(x1 - x_ref).max() / scale, cond)
        assert np.abs(x2 - x1).max() <= 1e-12 * scale + 1e-3 * tol
        # residual of the structured solve is as small as LAPACK's
        r1 = np.abs(a @ x1 - rhs).max(); r0 = np.abs(a @ x_ref - rhs).max()
        assert r1 <= 50 * r0 + 1e-14 * np.abs(rhs).max()
    assert nfail == 0, 'threshold test (tau = %.2f) failed on %d sampled systems' % (TAU, nfail)


def _replay_schedule(slu, net, heat, S0, b, nwl):
    """numpy replay of slu_factor's pivot loop as the kernel runs it (xb_slu.cuh): round by round, warp by warp, lane by
    lane from the pivot schedule (structured_lu_schedule, xb_loader.cpp); the right-hand side is the pseudo-column at
    rhs_off.  Returns the store, the carried right-hand side and the first failing pivot (or None)."""
    nround, rhs_off = (int(v) for v in net.slu_table('sched_sizes/%d' % nwl, heat))
    warp = net.slu_table('sched_warp/%d' % nwl, heat).view(np.uint32).reshape(-1, nwl, 2)
    lane = net.slu_table('sched_lane/%d' % nwl, heat).view(np.uint32).reshape(-1, nwl, 32)
    assert nround % 3 == 0 and warp.shape[0] == nround + 3 and lane.shape[0] == nround + 3      # three idle rounds for the prefetch
    assert (warp[nround:, :, 1] & 0xff == 0).all()
    n, nh, nb = slu.n, slu.nh, slu.nb
    S = np.zeros(rhs_off + n + 16)
    S[:slu.nstore] = S0
    S[rhs_off:rhs_off + n] = b
    S[slu.bo + n * slu.bs: slu.bo + n * slu.bs + nb] = b[nh:]        # border entries ride in column n of the border block
    kfail = None
    touched = {}
    for r in range(nround):
        snap = S.copy()                                # every warp of a round reads the state the barrier left
        for w in range(nwl):
            lp, pv = int(warp[r, w, 0] & 0xffff), int(warp[r, w, 0] >> 16)
            nrow, ntest, k = int(warp[r, w, 1] & 15), int((warp[r, w, 1] >> 4) & 15), int(warp[r, w, 1] >> 8)
            assert pv == slu.cb[k] + k
            piv = snap[pv]
            rk = 1.0 / piv
            for t in range(ntest):
                if TAU * abs(snap[lp + t]) > abs(piv):
                    kfail = k if kfail is None else min(kfail, k)
            for l in range(32):
                ap, up = int(lane[r, w, l] & 0xffff), int(lane[r, w, l] >> 16)
                if ap == lp:
                    continue
                un = snap[up] * rk
                for t in range(nrow):
                    key = (k, ap + t)
                    assert key not in touched, 'element updated twice by one pivot'
                    touched[key] = 1
                    assert snap[ap + t] == S[ap + t], 'two warps of a round touch the same element'
                    S[ap + t] = snap[ap + t] - snap[lp + t] * un
    return S, kfail


@pytest.mark.parametrize('name,nwl', [('alpha', 4), ('torch47', 4), ('sn160_scrn', 4), ('sn160_heat', 4), ('sn160_scrn', 2), ('xrb304', 8)])
def test_pivot_schedule_replay(name, nwl, oracle_problem, gpu_problem):
    """The pivot schedule does exactly what the index arithmetic of the restatement does: same store and carried
    right-hand side, bit for bit (every element receives the same single update per pivot), same verdict of the column test."""
    op = oracle_problem(name)
    ctl, net, fx = gpu_problem(name)
    heat = 1 if ctl['iheat'] > 0 else 0
    slu = Slu(net, heat)
    nh = slu.nh
    for a, rhs in _systems(op, name, heat, 2):
        S0 = slu.scatter(a)
        Sref = S0.copy()
        # the restatement's heavy part only: stop before the border block by replaying its loop here
        b = rhs[slu.perm].copy()
        bref = b.copy()
        fail_ref = None
        ntest = slu.n - heat
        for k in range(nh):
            piv = Sref[slu.cb[k] + k]
            rk = 1.0 / piv
            rows = slu.rows(k)
            l = np.array([Sref[slu.addr(i, k)] for i in rows])
            if any(TAU * abs(v) > abs(piv) for v, i in zip(l, rows) if i < ntest) and fail_ref is None:
                fail_ref = k
            for j in slu.cols(k):
                u = Sref[slu.cb[j] + k] * rk
                for v, i in zip(l, rows):
                    Sref[slu.addr(i, j)] -= v * u
            u = bref[k] * rk
            for v, i in zip(l, rows):
                bref[i] -= v * u
        S, kfail = _replay_schedule(slu, net, heat, S0, b, nwl)
        rhs_off = int(net.slu_table('sched_sizes/%d' % nwl, heat)[1])
        assert np.array_equal(S[:slu.bo + slu.n * slu.bs], Sref[:slu.bo + slu.n * slu.bs])
        assert np.array_equal(S[rhs_off:rhs_off + nh], bref[:nh])
        assert np.array_equal(S[slu.bo + slu.n * slu.bs: slu.bo + slu.n * slu.bs + slu.nb], bref[nh:])
        assert kfail == fail_ref


def _replay_solve_table(slu, net, heat, S, b):
    """numpy replay of slu_solve's phase C (xb_slu.cuh) as the warp runs it: 32 window rows in 'registers' (lane = row mod
    32), one table entry per lane and step, the owner's product broadcast, the row leaving the window stored and the row
    entering it fetched; rows above the window updated in memory at the listed steps.  b: right-hand side after phases A
    and B (border unknowns solved, their columns subtracted from the heavy rows); returns b with the heavy unknowns."""
    words = net.slu_table('sol_lane', heat).view(np.uint32).reshape(-1, 32, 2)
    outl = net.slu_table('sol_out', heat).tolist()
    ent = np.empty((words.shape[0], 32, 4), np.int64)
    ent[:, :, 0] = words[:, :, 0] & 0xffff; ent[:, :, 1] = words[:, :, 0] >> 16
    ent[:, :, 2] = words[:, :, 1] & 0xffff; ent[:, :, 3] = words[:, :, 1] >> 16
    nh = slu.nh
    ngroup = words.shape[0] - 2
    assert ngroup % 2 == 0 and 4 * ngroup >= nh and (ent[ngroup:] == 0xffff).all() and outl[-1] < -nh
    b = b.copy()
    k = nh - 1
    bv = np.array([b[k - ((k - l) & 31)] if k - ((k - l) & 31) >= 0 else 0.0 for l in range(32)])
    be = np.zeros(32)
    enter = np.zeros(32, bool)
    oi = 0
    for g in range(ngroup):
        for s4 in range(4):
            v = np.array([S[a] if a != 0xffff else 0.0 for a in ent[g, :, s4]])
            own = np.arange(32) == (k & 31)
            x = (bv * v)[k & 31]
            bv = np.where(enter, be, bv)
            if k == outl[oi]:
                for i in range(slu.cf[k], k - 31):
                    b[i] -= S[slu.cb[k] + i] * x
                oi += 1
            bv = bv - v * x
            if k >= 0:
                b[k] = x
                be[k & 31] = b[k - 32] if k >= 32 else b[0]
            enter = own
            k -= 1
    assert outl[oi] < 0
    return b


@pytest.mark.parametrize('name', ['alpha', 'torch47', 'sn160_scrn', 'sn160_heat', 'xrb304'])
def test_solve_table_replay(name, oracle_problem, gpu_problem):
    """The table-driven band back-substitution equals the restatement's column sweep (same products, same order per row:
    bit for bit), including the columns that reach above the 32-row window."""
    op = oracle_problem(name)
    ctl, net, fx = gpu_problem(name)
    heat = 1 if ctl['iheat'] > 0 else 0
    slu = Slu(net, heat)
    n, nh = slu.n, slu.nh
    nout = 0
    for a, rhs in _systems(op, name, heat, 2):
        S = slu.scatter(a)
        rinv, ipb, fail, bf = slu.factor(S, rhs[slu.perm])
        x_ref = slu.solve(S, rinv, ipb, bf, True)
        # phases A and B in the restatement's arithmetic: border unknowns, then their columns off the heavy rows
        b = bf.copy()
        for k in range(n - 1, nh - 1, -1):
            x = b[k] * rinv[k]
            b[k] = x
            for i in range(slu.cf[k], k):
                b[i] -= S[slu.addr(i, k)] * x
        Sr = S.copy()
        for k in range(nh):
            Sr[slu.cb[k] + k] = rinv[k]              # the kernel keeps the reciprocals over the pivots
        got = _replay_solve_table(slu, net, heat, Sr, b)
        assert np.array_equal(got[:nh], x_ref[:nh])
        nout += sum(1 for k in range(nh) if slu.cf[k] <= k - 32)
    if name in ('sn160_scrn', 'xrb304'):
        assert nout > 0                             # the out-of-window path is exercised


def _replay_schedule2(slu, net, heat, S0, b, nwl):
    """numpy replay of the rank-2 pivot loop (slu_factor with the pair schedule, xb_slu.cuh): one round eliminates the
    pivots (k, k+1); every lane forms U(k+1, j) for its own column, the flagged item stores it one barrier later."""
    nround, rhs_off, pair = (int(v) for v in net.slu_table('sched2_sizes/%d' % nwl, heat))
    assert pair == 1
    warp = net.slu_table('sched2_warp/%d' % nwl, heat).view(np.uint32).reshape(-1, nwl, 4)
    lane = net.slu_table('sched2_lane/%d' % nwl, heat).view(np.uint32).reshape(-1, nwl, 32, 2)
    assert nround % 3 == 0 and warp.shape[0] == nround + 3 and lane.shape[0] == nround + 3
    n, nh, nb = slu.n, slu.nh, slu.nb
    S = np.zeros(rhs_off + n + 16)
    S[:slu.nstore] = S0
    S[rhs_off:rhs_off + n] = b
    S[slu.bo + n * slu.bs: slu.bo + n * slu.bs + nb] = b[nh:]
    kfail = None
    pend = []                                         # deferred stores of U(k+1, .): after the next barrier
    recip = None                                      # (pv1, r1, pv2 or None, r2): stored once the schedule has left the pair
    for r in range(nround):
        for ad, v in pend:
            S[ad] = v
        pend = []
        pv1_r = int(warp[r, 0, 2] & 0xffff)
        if recip is not None and recip[0] != pv1_r:
            S[recip[0]] = recip[1]
            if recip[2] is not None:
                S[recip[2]] = recip[3]
            recip = None
        snap = S.copy()
        for w in range(nwl):
            lp1, lp2 = int(warp[r, w, 0] & 0xffff), int(warp[r, w, 0] >> 16)
            f = int(warp[r, w, 1])
            nrow, nt1, nt2, commit, single = f & 15, (f >> 4) & 15, (f >> 8) & 15, (f >> 12) & 1, (f >> 13) & 1
            pv1, u12 = int(warp[r, w, 2] & 0xffff), int(warp[r, w, 2] >> 16)
            k = int(warp[r, w, 3])
            assert pv1 == slu.cb[k] + k and pv1 == pv1_r
            piv1, l21, u12v, piv2 = snap[pv1], snap[pv1 + 1], snap[u12], snap[u12 + 1]
            rk1 = 1.0 / piv1
            un1k1 = -(u12v * rk1)
            p2 = piv2 + l21 * un1k1
            rk2 = 0.0 if single else 1.0 / p2
            if not single:
                assert u12 == slu.cb[k + 1] + k
                if TAU * abs(l21) > abs(piv1):
                    kfail = k if kfail is None else min(kfail, k)
            for t in range(nt1):
                if TAU * abs(snap[lp1 + t]) > abs(piv1):
                    kfail = k if kfail is None else min(kfail, k)
            for t in range(nt2):
                l2t = snap[lp2 + t] + snap[lp1 + t] * un1k1
                if TAU * abs(l2t) > abs(p2):
                    kfail = k + 1 if kfail is None else min(kfail, k + 1)
            recip = (pv1, rk1, None if single else u12 + 1, rk2)
            for l in range(32):
                ap, up1 = int(lane[r, w, l, 0] & 0xffff), int(lane[r, w, l, 0] >> 16)
                up2, col2 = int(lane[r, w, l, 1] & 0xffff), int(lane[r, w, l, 1] >> 16) & 1
                if ap == lp1:
                    continue
                un1 = -(snap[up1] * rk1)
                upr = snap[up2] + l21 * un1
                un2 = 0.0 if (col2 or single) else -(upr * rk2)
                for t in range(nrow):
                    l1 = snap[lp1 + t]
                    l2 = snap[lp2 + t] + l1 * un1k1
                    assert snap[ap + t] == S[ap + t], 'two warps of a round touch the same element'
                    S[ap + t] = (snap[ap + t] + l1 * un1) + l2 * un2
                if commit and not col2:
                    pend.append((up2, upr))
    for ad, v in pend:
        S[ad] = v
    return S, kfail, recip


@pytest.mark.parametrize('name,nwl', [('alpha', 4), ('torch47', 4), ('sn160_scrn', 4), ('sn160_heat', 4), ('sn160_scrn', 2), ('xrb304', 8)])
def test_pair_schedule_replay(name, nwl, oracle_problem, gpu_problem, monkeypatch):
    """The rank-2 schedule (two pivots per round; XB_SLU_PAIR=1, parked: measured slower on B200) leaves the same store,
    carried right-hand side and column-test verdict as the pivot-by-pivot elimination: each element receives the same
    two updates in the same order, so the match is exact."""
    monkeypatch.setenv('XB_SLU_PAIR', '1')
    op = oracle_problem(name)
    ctl, net, fx = gpu_problem(name)
    heat = 1 if ctl['iheat'] > 0 else 0
    slu = Slu(net, heat)
    nh = slu.nh
    for a, rhs in _systems(op, name, heat, 2):
        S0 = slu.scatter(a)
        Sref = S0.copy()
        b = rhs[slu.perm].copy()
        bref = b.copy()
        fail_ref = None
        ntest = slu.n - heat
        for k in range(nh):
            piv = Sref[slu.cb[k] + k]
            rk = 1.0 / piv
            rows = slu.rows(k)
            l = np.array([Sref[slu.addr(i, k)] for i in rows])
            if any(TAU * abs(v) > abs(piv) for v, i in zip(l, rows) if i < ntest) and fail_ref is None:
                fail_ref = k
            for j in slu.cols(k):
                u = -(Sref[slu.cb[j] + k] * rk)
                for v, i in zip(l, rows):
                    Sref[slu.addr(i, j)] += v * u
            u = -(bref[k] * rk)
            for v, i in zip(l, rows):
                bref[i] += v * u
        S, kfail, recip = _replay_schedule2(slu, net, heat, S0, b, nwl)
        rhs_off = int(net.slu_table('sched2_sizes/%d' % nwl, heat)[1])
        end = slu.bo + slu.n * slu.bs
        # pivots: the replay has stored the reciprocals of all pairs but the last
        for k in range(nh):
            ad = slu.cb[k] + k
            if recip is not None and ad in (recip[0], recip[2]):
                assert S[ad] == Sref[ad]
            else:
                assert S[ad] == 1.0 / Sref[ad]
                S[ad] = Sref[ad]
        assert np.array_equal(S[:end], Sref[:end])
        assert np.array_equal(S[rhs_off:rhs_off + nh], bref[:nh])
        assert np.array_equal(S[end:end + slu.nb], bref[nh:])
        assert kfail == fail_ref
