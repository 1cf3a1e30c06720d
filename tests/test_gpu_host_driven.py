"""The drop-in scenario of SURVEY §8b end to end: the HOST keeps XNet's control flow -- full_net's step loop
(xnet_evolve.F90:136-222), timestep (xnet_integrate.F90:13-268) and solve_be's retry ladder
(xnet_integrate_be.F90:32-231), restated here in Python as the stand-in for the Fortran driver -- and calls the library
only where the reference calls its plug-ins: cross_sect/yderiv for the first time step and step_be for every Newton
loop (xb_step_be: one launch per attempt).  In strict arithmetic the resulting accepted-step sequence and abundances
must equal the oracle's complete-run history bit for bit, and therefore also the fused xb_evolve path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DBL_MIN = np.finfo(np.float64).tiny


def _t9rhofind(fx, tf):
    """xnet_conditions.F90:72-99 (linear scan from the start, linear interpolation)."""
    th, t9h, rhoh = fx['th'], fx['t9h'], fx['rhoh']
    ns = len(th)
    if ns == 1:
        return 1, t9h[0], rhoh[0]
    n = int(np.searchsorted(th, tf, side='left')) + 1        # first n (1-based) with tf <= th(n)
    if 1 < n <= ns:
        rdt = 1.0 / (th[n - 1] - th[n - 2]); dt = tf - th[n - 2]
        return n, dt * rdt * (t9h[n - 1] - t9h[n - 2]) + t9h[n - 2], dt * rdt * (rhoh[n - 1] - rhoh[n - 2]) + rhoh[n - 2]
    if n == 1:
        return n, t9h[0], rhoh[0]
    return n, t9h[ns - 1], rhoh[ns - 1]


class HostDriver:
    """One zone, iheat = 0, trajectory-driven: the state full_net keeps on the host."""
    def __init__(self, ctx, ctl, fx):
        self.ctx, self.ctl, self.fx = ctx, ctl, fx
        self.y = fx['ystart'].copy(); self.yo = self.y.copy(); self.yt = self.y.copy()
        self.t = fx['tstart']; self.tstop = fx['tstop']; self.tdelstart = fx['tdelstart']
        self.tdel = 0.0; self.tdel_next = 0.0; self.tdel_old = 0.0
        self.nt = fx['nstart']; self.ntt = self.nt
        self.t9 = fx['t9start']; self.rho = fx['rhostart']; self.t9t = self.t9; self.rhot = self.rho; self.tt = self.t
        self.nh = np.array([fx['nh']], np.int32)
        self.ktot = np.zeros(5, int)
        self.ydot = np.zeros_like(self.y)

    def timestep(self):
        c = self.ctl
        self.tdel_old = self.tdel
        mask_init = abs(self.tdel_old) < DBL_MIN
        if mask_init:
            self.ctx.cross_sect(self.yt[None], [self.t9t], [self.rhot]); self.ktot[4] += 1
            self.ydot = self.ctx.yderiv(self.yt[None])[0]; self.ktot[3] += 1
        tdel_stop = self.tstop - self.t
        if mask_init:
            changest = 0.5 if self.tdelstart < 0.0 else 0.01
            changey = changest * min(0.1, c['changemx'])
            tdel_init = min(changest * abs(self.tdelstart), tdel_stop)
            self.tdel_next = 1.0e-4 * tdel_stop if self.fx['nh'] > 2 else tdel_stop
        else:
            changey = c['changemx']; tdel_init = 0.0
        if self.tdel_old >= 0.0:
            y = self.y
            yfloor = np.maximum(y, c['yacc'])
            with np.errstate(all='ignore'):
                r = np.abs(self.ydot / yfloor) if mask_init else np.abs((y - self.yo) / yfloor) / self.tdel_old
            rtau = np.where(y > c['ymin'], r, 0.0)
            imax = int(np.argmax(rtau))
            tdel_dy = changey / rtau[imax] if abs(rtau[imax]) > 0.0 else self.tdel_next
            self.tdel = max(tdel_init, min(tdel_stop, self.tdel_next, tdel_dy, self.tdel_next))   # tdel_dt9 = tdel_next (iheat = 0)
        else:
            self.tdel = -self.tdel_old
        self.tt = min(self.t + self.tdel, self.tstop)
        if self.fx['nh'] > 1:
            self.ntt, self.t9t, self.rhot = _t9rhofind(self.fx, self.tt)
            if self.ntt - 1 > self.nt:
                for j in range(self.nt, self.ntt):
                    t9j = self.fx['t9h'][j - 1]; rhoj = self.fx['rhoh'][j - 1]
                    if (t9j > self.t9 and t9j > self.t9t) or (t9j < self.t9 and t9j < self.t9t) or \
                            (rhoj > self.rho and rhoj > self.rhot) or (rhoj < self.rho and rhoj < self.rhot):
                        self.tdel = self.fx['th'][j - 1] - self.t
                        break
                self.tt = min(self.t + self.tdel, self.tstop)
            dtherm = 0.0
            for _ in range(10):
                self.ntt, self.t9t, self.rhot = _t9rhofind(self.fx, self.tt)
                if self.t9 > 0.0:
                    dtherm = 10.0 * abs(self.t9t - self.t9) / self.t9 + abs(self.rhot - self.rho) / self.rho
                if dtherm < 0.1:
                    break
                self.tdel = 0.5 * self.tdel
                self.tt = min(self.t + self.tdel, self.tstop)

    def solve_be(self):
        """Returns its (0 = step accepted, 1 = failed)."""
        c = self.ctl; ktsmx = 10; inr = 0; mykts = 0
        for kts in range(1, ktsmx + 1):
            r = self.ctx.step_be(self.y[None], self.yt[None], [self.tdel], [self.t9], [self.t9t], [self.rhot], nh=self.nh)
            inr = int(r['inr'][0]); self.yt = r['yt'][0]; self.ydot = r['ydot'][0]
            self.ktot[2:5] += r['kcount'][0][2:5]
            if inr == 0:
                self.tdel = self.tdel / c['tdel_maxmult']; self.tt = self.t + self.tdel
                mykts = kts + 1
                self.yt = self.y.copy()
                self.ktot[1] += c['kitmx'] + 1
                self.ntt, self.t9t, self.rhot = _t9rhofind(self.fx, self.tt)
                if kts == ktsmx - 1:
                    self.tdel = 0.0; self.tdelstart = 0.0
                    self.timestep()
            else:
                mykts = kts; self.ktot[1] += inr
                break
        self.ktot[0] += mykts
        if inr > 0:
            self.tdel_next = self.tdel * c['tdel_maxmult']
            self.nt = self.ntt; self.t = self.tt; self.t9 = self.t9t; self.rho = self.rhot
            self.yo = self.y; self.y = self.yt.copy()
            return 0
        return 1


@pytest.mark.parametrize('name,nsteps', [('alpha', None), ('torch47', 400)])
def test_host_driven_evolve_matches_oracle_bit_for_bit(gpu_problem, oracle_problem, name, nsteps):
    import xnet_b200 as X
    ctl, net, fx = gpu_problem(name)
    ctx = X.Context(net, ctl, nzevolve=1, arith=1)
    ref = oracle_problem(name).run(nzone=1, hist_cap=4000, fast='pm')
    n = int(ref['hist_n'][0])
    d = HostDriver(ctx, ctl, fx[0])
    launches0 = X.lib().xb_launch_count(0)
    rec = [(0, d.t, d.tdel, d.y.copy())]
    kmax = ctl['kstmx'] if nsteps is None else nsteps
    for kstep in range(1, kmax + 1):
        d.timestep()
        if d.solve_be() != 0:
            break
        rec.append((kstep, d.t, d.tdel, d.y.copy()))
        if d.t >= d.tstop:
            break
    m = len(rec)
    assert m == (n if nsteps is None else nsteps + 1) and m <= n
    assert np.array_equal([r[0] for r in rec], ref['hist_kstep'][0, :m])
    assert np.array_equal([r[1] for r in rec], ref['hist_t'][0, :m])
    assert np.array_equal([r[2] for r in rec], ref['hist_tdel'][0, :m])
    assert np.array_equal(np.stack([r[3] for r in rec]), ref['hist_y'][0, :m])
    if nsteps is None:      # complete run: the counters too (TS, NR, Jacobian builds, yderiv and cross_sect calls)
        assert np.array_equal(d.ktot, ref['ktot'][0])
    # one library launch per Newton loop (plus the two of the very first time step), not one per Newton iteration
    assert X.lib().xb_launch_count(0) - launches0 == d.ktot[0] + 2
    ctx.close()
