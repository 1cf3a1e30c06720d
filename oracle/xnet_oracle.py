"""ctypes binding of the C oracle (TEST INFRASTRUCTURE ONLY; product code must not import this).

Usage:
    prob = Problem.from_control(control_text, root)       # loads network + fixtures via xnet_loader
    res  = prob.run(nzone=..., hist_cap=..., nthreads=...) # ox_full_net
"""
import ctypes as C
import os
import subprocess
import numpy as np
from . import xnet_loader as L

_HERE = os.path.dirname(os.path.abspath(__file__))
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class OxNet(C.Structure):
    _fields_ = [('ny', C.c_int), ('nffn', C.c_int), ('nnnu', C.c_int),
                ('nr', C.c_int * 4), ('nan', C.c_int * 4),
                ('rc', c_dp * 4), ('ni', c_ip * 4), ('iwk', c_ip * 4), ('irev', c_ip * 4),
                ('iffn', c_ip), ('la', c_ip), ('le', c_ip),
                ('mu', c_ip * 4), ('a', c_dp * 4),
                ('aa', c_dp), ('zz', c_dp), ('mex', c_dp), ('be', c_dp),
                ('mex_n', C.c_double), ('mex_p', C.c_double),
                ('g', c_dp), ('t9i', c_dp),
                ('ffnsum', c_dp), ('ffnenu', c_dp), ('ffn_ec', c_dp), ('ffn_beta', c_dp),
                ('ffn_qval', c_dp), ('has_logft', c_ip), ('helm', c_dp),
                ('mflx', C.c_int), ('ifl', c_ip * 4), ('dcflx', c_dp)]


class OxCtl(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('iweak0', 'iscrn', 'iheat', 'isolv', 'kstmx', 'kitmx', 'ijac', 'iconvc')] + \
               [(n, C.c_double) for n in ('changemx', 'changemxt', 'yacc', 'tolm', 'tolc', 'tolt9', 'ymin', 'tdel_maxmult')]


class OxZones(C.Structure):
    _fields_ = [('nfix', C.c_int), ('nh', c_ip), ('off', c_ip), ('th', c_dp), ('t9h', c_dp), ('rhoh', c_dp),
                ('tstart', c_dp), ('tstop', c_dp), ('tdelstart', c_dp), ('nstart', c_ip),
                ('t9start', c_dp), ('rhostart', c_dp), ('yestart', c_dp), ('ystart', c_dp),
                ('xext', c_dp), ('aext', c_dp), ('zext', c_dp),
                ('nzone', C.c_int), ('zfix', c_ip), ('rho_scale', c_dp),
                ('rs_t', c_dp), ('rs_tdel', c_dp), ('rs_t9', c_dp), ('rs_rho', c_dp), ('rs_t9o', c_dp),
                ('rs_y', c_dp), ('rs_yo', c_dp), ('rs_kstep0', c_ip)]


class OxOut(C.Structure):
    _fields_ = [('y', c_dp), ('t', c_dp), ('t9', c_dp), ('rho', c_dp), ('tdel', c_dp),
                ('kstep', c_ip), ('ktot', c_ip), ('status', c_ip),
                ('hist_cap', C.c_int), ('hist_n', c_ip), ('hist_kstep', c_ip), ('hist_kmon', c_ip),
                ('hist_t', c_dp), ('hist_t9', c_dp), ('hist_rho', c_dp), ('hist_tdel', c_dp),
                ('hist_edot', c_dp), ('hist_sqnu', c_dp), ('hist_y', c_dp), ('hist_flx', c_dp)]


def build(force=False):
    so = os.path.join(_HERE, 'liboracle.so')
    if force or not all(os.path.exists(os.path.join(_HERE, n)) for n in ('liboracle.so', 'liboracle_fast.so', 'liboracle_pm.so')):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return so


_libs = {}


def lib(fast=False):
    key = fast if isinstance(fast, str) else ('fast' if fast else 'strict')
    if key not in _libs:
        build()
        l = C.CDLL(os.path.join(_HERE, {'fast': 'liboracle_fast.so', 'strict': 'liboracle.so', 'pm': 'liboracle_pm.so'}[key]))
        l.ox_full_net.restype = C.c_int
        l.ox_full_net.argtypes = [C.POINTER(OxNet), C.POINTER(OxCtl), C.POINTER(OxZones), C.POINTER(OxOut), C.c_int]
        l.ox_dgetrf.argtypes = [C.c_int, C.c_int, c_dp, C.c_int, c_ip, c_ip]
        l.ox_dgetrs.argtypes = [C.c_int, C.c_int, c_dp, C.c_int, c_ip, c_dp, C.c_int]
        l.ox_eval_point.restype = C.c_int
        l.ox_eval_point.argtypes = [C.POINTER(OxNet), C.POINTER(OxCtl)] + [C.c_double] * 2 + [c_dp] + \
            [C.c_double] * 5 + [c_dp, c_dp, c_dp, c_dp]
        l.ox_eval_point_heat.restype = C.c_int
        l.ox_eval_point_heat.argtypes = [C.POINTER(OxNet), C.POINTER(OxCtl)] + [C.c_double] * 2 + [c_dp] + \
            [C.c_double] * 5 + [c_dp] * 7
        l.ox_flux_point.restype = C.c_int
        l.ox_flux_point.argtypes = [C.POINTER(OxNet), C.POINTER(OxCtl)] + [C.c_double] * 2 + [c_dp] + [C.c_double] * 3 + [c_dp]
        l.ox_step_be_point.restype = C.c_int
        l.ox_step_be_point.argtypes = [C.POINTER(OxNet), C.POINTER(OxCtl), c_dp, c_dp, C.c_double, C.c_double, c_dp,
                                       C.c_double, C.c_int] + [C.c_double] * 3 + [c_ip, c_dp, c_dp, c_ip]
        _libs[key] = l
    return _libs[key]


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _pd(a):
    return a.ctypes.data_as(c_dp)


def _pi(a):
    return a.ctypes.data_as(c_ip)


def load_helm(root):
    """13 planes [13][201][541] of the Helmholtz table (data/helm_table.bin, see tools/make_helm_table.py)"""
    a = np.fromfile(os.path.join(root, 'helm_table.bin'), dtype='<f8')
    assert a.size == 13 * 201 * 541, a.size
    return a


class Problem:
    def __init__(self, ctl, nuc, net, ffn, fixtures, helm=None):
        self.ctl, self.nuc, self.net, self.ffn, self.fixtures = ctl, nuc, net, ffn, fixtures
        self.helm = helm
        self._keep = []
        self._build_net()
        self._build_ctl()
        self._build_fix()

    @classmethod
    def from_control(cls, control_text, root, helm=None):
        ctl, nuc, net, ffn, zones = L.load_problem(control_text, root)
        if helm is None and (ctl['iscrn'] > 0 or ctl['iheat'] > 0):
            helm = load_helm(root)
        return cls(ctl, nuc, net, ffn, zones, helm)

    def _k(self, a):
        self._keep.append(a)
        return a

    def _build_net(self):
        nuc, net, ffn = self.nuc, self.net, self.ffn
        n = OxNet()
        n.ny = nuc['ny']; n.nffn = net['nffn']; n.nnnu = net['nnnu']
        ny = nuc['ny']
        for c in range(4):
            n.nr[c] = net['nreac'][c]; n.nan[c] = net['nan'][c]
            n.rc[c] = _pd(self._k(_d(net['rc%d' % (c + 1)])))
            n.ni[c] = _pi(self._k(_i(net['n%d' % (c + 1)])))
            n.iwk[c] = _pi(self._k(_i(net['iwk%d' % (c + 1)])))
            n.irev[c] = _pi(self._k(_i(net['irev%d' % (c + 1)])))
            n.mu[c] = _pi(self._k(_i(net['mu%d' % (c + 1)] if net['nan'][c] else np.zeros(1))))
            n.a[c] = _pd(self._k(_d(net['a%d' % (c + 1)] if net['nan'][c] else np.zeros(1))))
        n.iffn = _pi(self._k(_i(net['iffn'])))
        n.la = _pi(self._k(_i(net['la']))); n.le = _pi(self._k(_i(net['le'])))
        n.aa = _pd(self._k(_d(nuc['aa']))); n.zz = _pd(self._k(_d(nuc['zz'])))
        n.mex = _pd(self._k(_d(nuc['mex']))); n.be = _pd(self._k(_d(nuc['be'])))
        n.mex_n = float(nuc['mex_n']); n.mex_p = float(nuc['mex_p'])
        n.g = _pd(self._k(_d(nuc['g']))); n.t9i = _pd(self._k(_d(nuc['t9i'])))
        for nm in ('ffnsum', 'ffnenu', 'ffn_ec', 'ffn_beta', 'ffn_qval'):
            setattr(n, nm, _pd(self._k(_d(ffn[nm]))))
        n.has_logft = _pi(self._k(_i(ffn['has_logft'])))
        if self.helm is not None:
            n.helm = _pd(self._k(_d(self.helm)))
        n.mflx = int(net.get('mflx', 0))
        if n.mflx > 0:
            for c in range(4):
                n.ifl[c] = _pi(self._k(_i(net['ifl%d' % (c + 1)] if net['nreac'][c] else np.zeros(1))))
            n.dcflx = _pd(self._k(_d(net['dcflx'])))
        self.cnet = n

    def _build_ctl(self):
        c = OxCtl()
        for nm, _ in OxCtl._fields_:
            setattr(c, nm, self.ctl[nm])
        self.cctl = c

    def _build_fix(self):
        fx = self.fixtures
        self.fix_arrays = dict(
            nh=_i([f['nh'] for f in fx]),
            off=_i(np.concatenate([[0], np.cumsum([f['nh'] for f in fx])[:-1]])),
            th=_d(np.concatenate([f['th'] for f in fx])), t9h=_d(np.concatenate([f['t9h'] for f in fx])),
            rhoh=_d(np.concatenate([f['rhoh'] for f in fx])),
            tstart=_d([f['tstart'] for f in fx]), tstop=_d([f['tstop'] for f in fx]),
            tdelstart=_d([f['tdelstart'] for f in fx]), nstart=_i([f['nstart'] for f in fx]),
            t9start=_d([f['t9start'] for f in fx]), rhostart=_d([f['rhostart'] for f in fx]),
            yestart=_d([f['yestart'] for f in fx]), ystart=_d(np.stack([f['ystart'] for f in fx])),
            xext=_d([f['xext'] for f in fx]), aext=_d([f['aext'] for f in fx]), zext=_d([f['zext'] for f in fx]))

    def flux_point(self, t9, rho, y, which='strict', xext=0.0, aext=1.0, zext=0.0):
        """flux (xnet_flux.F90:67-150) of one state after a rate evaluation at (t9, rho, y): flx[mflx]."""
        flx = np.zeros(max(int(self.net['mflx']), 1)); yy = _d(y)
        rc = lib(which).ox_flux_point(C.byref(self.cnet), C.byref(self.cctl), t9, rho, _pd(yy), xext, aext, zext, _pd(flx))
        if rc:
            raise RuntimeError('ox_flux_point rc=%d' % rc)
        return flx[:int(self.net['mflx'])]

    def run(self, nzone=None, hist_cap=0, nthreads=0, rho_scale=None, fast=False, ctl_override=None,
            restart=None, zfix=None, flux=False):
        fa = self.fix_arrays
        K = len(self.fixtures)
        nzone = nzone or K
        ny = self.nuc['ny']
        z = OxZones()
        z.nfix = K
        for nm in ('nh', 'off', 'nstart'):
            setattr(z, nm, _pi(fa[nm]))
        for nm in ('th', 't9h', 'rhoh', 'tstart', 'tstop', 'tdelstart', 't9start', 'rhostart', 'yestart',
                   'ystart', 'xext', 'aext', 'zext'):
            setattr(z, nm, _pd(fa[nm]))
        z.nzone = nzone
        zfix = _i(np.arange(nzone) % K) if zfix is None else _i(zfix)
        z.zfix = _pi(zfix)
        if rho_scale is not None:
            rs = _d(rho_scale); z.rho_scale = _pd(rs)
        if restart is not None:
            rs = {k: (_i(v) if k == 'kstep0' else _d(v)) for k, v in restart.items()}
            for nm in ('t', 'tdel', 't9', 'rho', 't9o', 'y', 'yo'):
                setattr(z, 'rs_' + nm, _pd(rs[nm]))
            z.rs_kstep0 = _pi(rs['kstep0'])
        cctl = self.cctl
        if ctl_override:
            cctl = OxCtl()
            for nm, _ in OxCtl._fields_:
                setattr(cctl, nm, ctl_override.get(nm, self.ctl[nm]))
        o = OxOut()
        res = dict(y=np.zeros((nzone, ny)), t=np.zeros(nzone), t9=np.zeros(nzone), rho=np.zeros(nzone),
                   tdel=np.zeros(nzone), kstep=np.zeros(nzone, np.int32), ktot=np.zeros((nzone, 5), np.int32),
                   status=np.zeros(nzone, np.int32))
        o.y = _pd(res['y']); o.t = _pd(res['t']); o.t9 = _pd(res['t9']); o.rho = _pd(res['rho']); o.tdel = _pd(res['tdel'])
        o.kstep = _pi(res['kstep']); o.ktot = _pi(res['ktot']); o.status = _pi(res['status'])
        o.hist_cap = hist_cap
        if hist_cap > 0:
            res.update(hist_n=np.zeros(nzone, np.int32), hist_kstep=np.zeros((nzone, hist_cap), np.int32),
                       hist_kmon=np.zeros((nzone, hist_cap, 2), np.int32),
                       hist_t=np.zeros((nzone, hist_cap)), hist_t9=np.zeros((nzone, hist_cap)),
                       hist_rho=np.zeros((nzone, hist_cap)), hist_tdel=np.zeros((nzone, hist_cap)),
                       hist_edot=np.zeros((nzone, hist_cap)), hist_sqnu=np.zeros((nzone, hist_cap)),
                       hist_y=np.zeros((nzone, hist_cap, ny)))
            o.hist_n = _pi(res['hist_n']); o.hist_kstep = _pi(res['hist_kstep']); o.hist_kmon = _pi(res['hist_kmon'])
            for nm in ('hist_t', 'hist_t9', 'hist_rho', 'hist_tdel', 'hist_edot', 'hist_sqnu', 'hist_y'):
                setattr(o, nm, _pd(res[nm]))
            if flux and int(self.net.get('mflx', 0)) > 0:
                res['hist_flx'] = np.zeros((nzone, hist_cap, int(self.net['mflx'])))
                o.hist_flx = _pd(res['hist_flx'])
        rc = lib(fast).ox_full_net(C.byref(self.cnet), C.byref(cctl), C.byref(z), C.byref(o), nthreads)
        res['rc'] = rc
        return res

    def eval_point(self, t9, rho, y, diag, mult, xext=0.0, aext=1.0, zext=0.0):
        ny = self.nuc['ny']
        nrt = sum(self.net['nreac'])
        cs = np.zeros(nrt); yd = np.zeros(ny); jac = np.zeros(ny * ny); ye = np.zeros(1)
        yy = _d(y)
        rc = lib().ox_eval_point(C.byref(self.cnet), C.byref(self.cctl), t9, rho, _pd(yy), xext, aext, zext,
                                 diag, mult, _pd(cs), _pd(yd), _pd(jac), _pd(ye))
        if rc:
            raise RuntimeError('ox_eval_point rc=%d' % rc)
        return cs, yd, jac.reshape(ny, ny).T.copy(), ye[0]   # jac returned as [row, col]


def eval_point_heat(prob, t9, rho, y, diag, mult, which='strict', xext=0.0, aext=1.0, zext=0.0):
    """cross_sect -> yderiv -> jacobian_build on one state with the self-heating outputs.
    Returns dict(csect, dcsect, ydot, jac [col,row] (msize x msize), ye, eos (cv, etae, detaedt9), t9dot)."""
    ny = prob.nuc['ny']; nrt = sum(prob.net['nreac']); m = ny + (1 if prob.ctl['iheat'] > 0 else 0)
    cs = np.zeros(nrt); dcs = np.zeros(nrt); yd = np.zeros(ny); jac = np.zeros(m * m); ye = np.zeros(1)
    eos = np.zeros(3); t9d = np.zeros(1)
    yy = _d(y)
    rc = lib(which).ox_eval_point_heat(C.byref(prob.cnet), C.byref(prob.cctl), t9, rho, _pd(yy), xext, aext, zext, diag, mult,
                                       _pd(cs), _pd(dcs), _pd(yd), _pd(jac), _pd(ye), _pd(eos), _pd(t9d))
    if rc:
        raise RuntimeError('ox_eval_point_heat rc=%d' % rc)
    return dict(csect=cs, dcsect=dcs, ydot=yd, jac=jac.reshape(m, m), ye=ye[0], eos=eos, t9dot=t9d[0])


def step_be_point(prob, y, yt, tdel, t9, t9t, rhot, nh=2, which='strict', xext=0.0, aext=1.0, zext=0.0):
    """One step_be call on an explicit state. Returns dict(inr, yt, t9t, ydot, t9dot, kcount)."""
    ny = prob.nuc['ny']
    yy = _d(y); ytt = _d(yt).copy(); t9tt = np.array([t9t], dtype=np.float64)
    inr = np.zeros(1, dtype=np.int32); yd = np.zeros(ny); t9d = np.zeros(1); kc = np.zeros(5, dtype=np.int32)
    rc = lib(which).ox_step_be_point(C.byref(prob.cnet), C.byref(prob.cctl), _pd(yy), _pd(ytt), tdel, t9, _pd(t9tt), rhot, nh,
                                     xext, aext, zext, _pi(inr), _pd(yd), _pd(t9d), _pi(kc))
    if rc:
        raise RuntimeError('ox_step_be_point rc=%d' % rc)
    return dict(inr=int(inr[0]), yt=ytt, t9t=t9tt[0], ydot=yd, t9dot=t9d[0], kcount=kc)


def dgetrf(a):
    """a: (n,n) array (row,col). Returns (lu, ipiv 1-based, info) via the LAPACK restatement."""
    n = a.shape[0]
    af = np.asfortranarray(a, dtype=np.float64).copy(order='F')
    ipiv = np.zeros(n, np.int32); info = np.zeros(1, np.int32)
    lib().ox_dgetrf(n, n, af.ctypes.data_as(c_dp), n, _pi(ipiv), _pi(info))
    return af, ipiv, int(info[0])


def dgetrs(lu, ipiv, b):
    n = lu.shape[0]
    luf = np.asfortranarray(lu, dtype=np.float64)
    bb = _d(b).copy()
    lib().ox_dgetrs(n, 1, luf.ctypes.data_as(c_dp), n, _pi(_i(ipiv)), _pd(bb), n)
    return bb
