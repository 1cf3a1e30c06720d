"""ctypes binding of libxnet_b200.so + the host-side mirror of XNet's driver inputs.

Names follow the reference: `read_controls` (xnet_controls.F90:164-396) parses a `control` file,
`Network` is the nuclear_data/reaction_data module state, `Context` the device-resident state that
`read_jacobian_data`/`be_init` allocate, and its methods are the module procedures
(`cross_sect`, `yderiv`, `jacobian_build`, `jacobian_decomp`, `jacobian_bksub`, `full_net`).
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class XnetB200Error(RuntimeError):
    pass


def lib_path():
    # XNET_B200_LIB: development aid for A/B-testing kernel variants built next to the library
    return os.environ.get('XNET_B200_LIB') or os.path.join(_HERE, 'libxnet_b200.so')


class XbCtl(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('iweak0', 'iscrn', 'iheat', 'isolv', 'kstmx', 'kitmx', 'ijac', 'iconvc')] + \
               [(n, C.c_double) for n in ('changemx', 'changemxt', 'yacc', 'tolm', 'tolc', 'tolt9', 'ymin', 'tdel_maxmult')] + \
               [('arith', C.c_int32), ('solver', C.c_int32)]


class XbZones(C.Structure):
    _fields_ = [('nfix', C.c_int32), ('nh', c_ip), ('off', c_ip), ('th', c_dp), ('t9h', c_dp), ('rhoh', c_dp),
                ('tstart', c_dp), ('tstop', c_dp), ('tdelstart', c_dp), ('nstart', c_ip),
                ('t9start', c_dp), ('rhostart', c_dp), ('yestart', c_dp), ('ystart', c_dp),
                ('xext', c_dp), ('aext', c_dp), ('zext', c_dp),
                ('nzone', C.c_int32), ('zfix', c_ip), ('rho_scale', c_dp),
                ('rs_t', c_dp), ('rs_tdel', c_dp), ('rs_t9', c_dp), ('rs_rho', c_dp), ('rs_t9o', c_dp),
                ('rs_y', c_dp), ('rs_yo', c_dp), ('rs_kstep0', c_ip)]


class XbOut(C.Structure):
    _fields_ = [('y', c_dp), ('t', c_dp), ('t9', c_dp), ('rho', c_dp), ('tdel', c_dp),
                ('kstep', c_ip), ('ktot', c_ip), ('status', c_ip),
                ('hist_cap', C.c_int32), ('hist_n', c_ip), ('hist_kstep', c_ip), ('hist_kmon', c_ip),
                ('hist_t', c_dp), ('hist_t9', c_dp), ('hist_rho', c_dp), ('hist_tdel', c_dp),
                ('hist_edot', c_dp), ('hist_sqnu', c_dp), ('hist_y', c_dp), ('hist_flx', c_dp)]


class XbNetArrays(C.Structure):
    """struct xb_net_arrays: the network tables as XNet's Fortran modules hold them (what the Fortran shim passes)."""
    _fields_ = [('ny', C.c_int32), ('nffn', C.c_int32), ('nnnu', C.c_int32), ('nreac', C.c_int32 * 4), ('nan', C.c_int32 * 4),
                ('rc', c_dp * 4), ('ni', c_ip * 4), ('ni_ld', C.c_int32 * 4), ('iwk', c_ip * 4), ('irev', c_ip * 4),
                ('la', c_ip), ('le', c_ip), ('mu', c_ip * 4), ('a', c_dp * 4),
                ('aa', c_dp), ('zz', c_dp), ('mex', c_dp), ('be', c_dp), ('mex_n', C.c_double), ('mex_p', C.c_double),
                ('g', c_dp), ('t9i', c_dp), ('ffnsum', c_dp), ('ffnenu', c_dp),
                ('ffn_ec', c_dp), ('ffn_beta', c_dp), ('ffn_qval', c_dp), ('has_logft', c_ip),
                ('mflx', C.c_int32), ('ifl', c_ip * 4), ('nflx', c_ip), ('iwflx', c_ip), ('qflx', c_dp), ('dcflx', c_dp)]


class XbTsHeader(C.Structure):
    _fields_ = [('descript', C.c_char_p * 3), ('data_desc', C.c_char_p), ('ctl', C.POINTER(XbCtl)),
                ('inab_file', C.c_char_p), ('abund_desc', C.c_char_p), ('thermo_file', C.c_char_p), ('thermo_desc', C.c_char_p),
                ('ny', C.c_int32), ('zz', c_dp), ('aa', c_dp), ('mflx', C.c_int32), ('ifl_orig', c_ip), ('ifl_term', c_ip)]


_lib = None

# every symbol include/xnet_b200.h declares (tests check the library exports all of them)
EXPORTS = ['xb_last_error', 'xb_version', 'xb_launch_count', 'xb_net_load', 'xb_net_load_helm', 'xb_net_from_arrays', 'xb_net_free',
           'xb_net_int', 'xb_net_itable', 'xb_net_dtable', 'xb_read_thermo', 'xb_read_inab', 'xb_zone_start',
           'xb_ctx_create', 'xb_ctx_destroy', 'xb_cross_sect', 'xb_get_csect', 'xb_yderiv', 'xb_jacobian_build', 'xb_jacobian_scale',
           'xb_get_jac', 'xb_set_jac', 'xb_jacobian_decomp', 'xb_get_indx', 'xb_jacobian_bksub',
           'xb_jacobian_solve', 'xb_get_dcsect', 'xb_get_eos', 'xb_flux', 'xb_set_flux_output', 'xb_step_be', 'xb_ts_write', 'xb_ev_write', 'xb_evolve', 'xb_zones_upload', 'xb_evolve_resident', 'xb_results_download',
           'xb_phase_cycles', 'xb_last_copy_bytes', 'xb_launch_config', 'xb_measure_fp64_peak', 'xb_last_kernel_ms', 'xb_solver_info', 'xb_slu_table', 'xb_slu_solve']


def lib():
    """Loads libxnet_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise XnetB200Error('libxnet_b200.so is not built (run __graft_entry__.build()); no CPU fallback exists')
        l = C.CDLL(p)
        l.xb_last_error.restype = C.c_char_p
        l.xb_launch_count.restype = C.c_longlong
        l.xb_launch_count.argtypes = [C.c_int]
        l.xb_net_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        l.xb_net_free.argtypes = [C.c_void_p]
        l.xb_net_from_arrays.argtypes = [C.POINTER(XbNetArrays), C.POINTER(C.c_void_p)]
        l.xb_net_load_helm.argtypes = [C.c_void_p, C.c_char_p]
        l.xb_net_int.argtypes = [C.c_void_p, C.c_char_p, c_ip]
        l.xb_net_itable.argtypes = [C.c_void_p, C.c_char_p, c_ip, C.c_int64, C.POINTER(C.c_int64)]
        l.xb_net_dtable.argtypes = [C.c_void_p, C.c_char_p, c_dp, C.c_int64, C.POINTER(C.c_int64)]
        l.xb_read_thermo.argtypes = [C.c_char_p, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp]
        l.xb_read_inab.argtypes = [C.c_void_p, C.c_char_p, c_dp, c_dp]
        l.xb_zone_start.argtypes = [C.c_void_p, C.c_int32, c_dp, c_dp, c_dp, c_dp, C.c_double, c_dp, c_dp, c_ip, c_dp]
        l.xb_ctx_create.argtypes = [C.c_void_p, C.POINTER(XbCtl), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        l.xb_ctx_destroy.argtypes = [C.c_void_p]
        l.xb_cross_sect.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
        l.xb_get_csect.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_dp]
        l.xb_yderiv.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_ip, c_dp, c_dp]
        l.xb_jacobian_build.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_ip, c_dp, c_dp]
        l.xb_jacobian_scale.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_ip, c_dp, c_dp]
        l.xb_get_jac.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_dp]
        l.xb_set_jac.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_dp]
        l.xb_jacobian_decomp.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_ip, c_ip]
        l.xb_get_indx.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_ip]
        l.xb_jacobian_bksub.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp]
        l.xb_jacobian_solve.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, c_ip]
        l.xb_get_dcsect.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_dp]
        l.xb_get_eos.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_dp, c_dp, c_dp, c_dp]
        l.xb_step_be.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_ip,
                                 c_dp, c_dp, c_ip]
        l.xb_flux.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_ip, c_dp, c_dp]
        l.xb_set_flux_output.argtypes = [C.c_void_p, C.c_int32]
        l.xb_ts_write.argtypes = [C.c_char_p, C.POINTER(XbTsHeader), C.c_int32, c_ip] + [c_dp] * 8
        l.xb_ev_write.argtypes = [C.c_char_p, C.c_int32, c_ip] + [c_dp] * 6 + [C.c_int32, C.c_char_p, c_dp, c_ip]
        l.xb_evolve.argtypes = [C.c_void_p, C.POINTER(XbZones), C.POINTER(XbOut)]
        l.xb_zones_upload.argtypes = [C.c_void_p, C.POINTER(XbZones), C.c_int32]
        l.xb_evolve_resident.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        l.xb_results_download.argtypes = [C.c_void_p, C.POINTER(XbOut)]
        l.xb_phase_cycles.argtypes = [C.c_void_p, c_dp]
        l.xb_last_copy_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        l.xb_launch_config.argtypes = [C.c_void_p, c_ip]
        l.xb_measure_fp64_peak.argtypes = [C.c_int32, c_dp]
        l.xb_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        l.xb_solver_info.argtypes = [C.c_void_p, c_dp]
        l.xb_slu_table.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, c_ip, C.c_int64, C.POINTER(C.c_int64)]
        l.xb_slu_solve.argtypes = [C.c_void_p, C.c_int32, c_dp, c_dp, c_dp, c_ip, C.c_int32]
        _lib = l
    return _lib


def _chk(rc):
    if rc != 0:
        raise XnetB200Error('xnet_b200 error %d: %s' % (rc, lib().xb_last_error().decode()))


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _pd(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _pi(a):
    return None if a is None else a.ctypes.data_as(c_ip)


# --------------------------------------------------------------------------------------------------
class Controls(dict):
    """Parsed `control` file (keys named as the xnet_controls module variables)."""
    def to_struct(self, arith=0, solver=0, **override):
        c = XbCtl()
        for nm, _ in XbCtl._fields_:
            if nm in ('arith', 'solver'):
                continue
            setattr(c, nm, override.get(nm, self[nm]))
        c.arith = arith
        c.solver = solver
        return c


def _block(lines, key):
    for i, ln in enumerate(lines[:1000]):
        if ln.startswith('##') and key in ln[2:82]:
            return i + 1
    return None


def _tok(line):
    return line.replace(',', ' ').split()[0]


def _f(s):
    return float(s.replace('D', 'E').replace('d', 'e'))


def read_controls(text):
    """Host mirror of read_controls (xnet_controls.F90:164-396): `## <Block>` markers located by
    substring search, values list-directed one per line in fixed order, optional blocks defaulted."""
    ln = text.split('\n')
    c = Controls()
    p = _block(ln, 'Job Controls')
    for k, nm in enumerate(('szone', 'nzone', 'iweak0', 'iscrn', 'iprocess')):
        c[nm] = int(_tok(ln[p + k]))
    p = _block(ln, 'Zone Batching Controls')
    c['nzbatchmx'] = int(_tok(ln[p])) if p is not None else 1
    p = _block(ln, 'Integration Controls')
    for k, nm in enumerate(('isolv', 'kstmx', 'kitmx', 'ijac', 'iconvc')):
        c[nm] = int(_tok(ln[p + k]))
    for k, nm in enumerate(('changemx', 'yacc', 'tolm', 'tolc', 'ymin', 'tdel_maxmult')):
        c[nm] = _f(_tok(ln[p + 5 + k]))
    p = _block(ln, 'Self-heating Controls')
    if p is not None:
        c['iheat'] = int(_tok(ln[p])); c['changemxt'] = _f(_tok(ln[p + 1])); c['tolt9'] = _f(_tok(ln[p + 2]))
    else:
        c['iheat'] = 0; c['changemxt'] = 1.0e-2; c['tolt9'] = 1.0e-4
    if c['isolv'] == 3:
        c['changemx'] = 1.0e+10; c['changemxt'] = 1.0e+10
    p = _block(ln, 'NSE Initial Conditions')
    c['t9nse'] = _f(_tok(ln[p])) if p is not None else 8.0
    p = _block(ln, 'Neutrinos')
    c['ineutrino'] = int(_tok(ln[p])) if p is not None else 0
    p = _block(ln, 'Output Controls')
    c['idiag'] = int(_tok(ln[p])); c['itsout'] = int(_tok(ln[p + 1]))
    c['ev_file_base'] = ln[p + 3][:80].strip(); c['bin_file_base'] = ln[p + 5][:80].strip()
    p = _block(ln, 'Input Controls')
    c['data_dir'] = ln[p + 1][:80].strip()
    files = [x[:80].strip() for x in ln[p + 3:p + 3 + 2 * c['nzone']] if x.strip()]
    inab, th = files[0::2], files[1::2]
    n = min(len(inab), len(th))
    inab, th = inab[:n], th[:n]
    if n == 1 and c['nzone'] > 1:  # base name + zero padded zone number (xnet_util.F90:341-370)
        w = len(str(c['nzone']))
        inab = [inab[0] + str(k + 1).zfill(w) for k in range(c['nzone'])]
        th = [th[0] + str(k + 1).zfill(w) for k in range(c['nzone'])]
    c['inab_file'] = inab; c['thermo_file'] = th
    return c


class Network:
    """Zone-invariant tables (nuclear_data + reaction_data), loaded and pre-processed by the C++
    host side of the library (xb_net_load)."""
    def __init__(self, data_dir, iweak0=1):
        self.h = C.c_void_p()
        _chk(lib().xb_net_load(data_dir.encode(), iweak0, C.byref(self.h)))
        self.data_dir = data_dir
        self.ny = self.int('ny')
        self.nreac = [self.int('nreac%d' % k) for k in (1, 2, 3, 4)]
        self.nan = [self.int('nan%d' % k) for k in (1, 2, 3, 4)]
        self.nffn = self.int('nffn'); self.nnnu = self.int('nnnu')
        self.mflx = self.int('mflx')     # matched reaction pairs (match_react, xnet_preprocess.F90:108-289)
        self.helm_loaded = False

    @classmethod
    def from_fortran_arrays(cls, t, iweak0=1):
        """xb_net_from_arrays: builds a Network from tables laid out as the Fortran modules hold them (the path the
        Fortran shim takes, fortran/xnet_jacobian_b200.F90 read_jacobian_data).  `t` maps the field names of
        xb_net_arrays to arrays: rc1..4 (7,nr) column-major i.e. shape [nr,7]; n1i..n4i shape [nr,ld] with ld = 5,6,6,6;
        la/le shape [ny,4]; ffn tables shape [143,nffn]; optional ifl1..4, nflx [mflx,8], iwflx, qflx, dcflx [mflx,2]."""
        a = XbNetArrays(); keep = []

        def pd(x):
            v = _d(x); keep.append(v); return _pd(v)

        def pi(x):
            v = _i(x); keep.append(v); return _pi(v)
        a.ny = int(t['ny']); a.nffn = int(t['nffn']); a.nnnu = int(t.get('nnnu', 0))
        for c in range(4):
            a.nreac[c] = int(t['nreac'][c]); a.nan[c] = int(t['nan'][c])
            a.rc[c] = pd(t['rc%d' % (c + 1)]); ni = _i(t['n%di' % (c + 1)]); keep.append(ni)
            a.ni[c] = _pi(ni); a.ni_ld[c] = ni.shape[1] if ni.ndim == 2 and ni.size else (5 if c == 0 else 6)
            a.iwk[c] = pi(t['iwk%d' % (c + 1)]); a.irev[c] = pi(t['irev%d' % (c + 1)])
            a.mu[c] = pi(t['mu%d' % (c + 1)]); a.a[c] = pd(t['a%d' % (c + 1)])
        a.la = pi(t['la']); a.le = pi(t['le'])
        for nm in ('aa', 'zz', 'mex', 'be', 'g', 't9i', 'ffnsum', 'ffnenu'):
            setattr(a, nm, pd(t[nm]))
        a.mex_n = float(t['mex_n']); a.mex_p = float(t['mex_p'])
        for nm in ('ffn_ec', 'ffn_beta', 'ffn_qval'):
            if t.get(nm) is not None:
                setattr(a, nm, pd(t[nm]))
        if t.get('has_logft') is not None:
            a.has_logft = pi(t['has_logft'])
        if t.get('mflx'):
            a.mflx = int(t['mflx'])
            for c in range(4):
                a.ifl[c] = pi(t['ifl%d' % (c + 1)])
            a.nflx = pi(t['nflx'])
            if t.get('iwflx') is not None:
                a.iwflx = pi(t['iwflx'])
            if t.get('qflx') is not None:
                a.qflx = pd(t['qflx'])
            if t.get('dcflx') is not None:
                a.dcflx = pd(t['dcflx'])
        self = cls.__new__(cls)
        self.h = C.c_void_p()
        _chk(lib().xb_net_from_arrays(C.byref(a), C.byref(self.h)))
        self.data_dir = t.get('data_dir', '')
        self.ny = self.int('ny')
        self.nreac = [self.int('nreac%d' % k) for k in (1, 2, 3, 4)]
        self.nan = [self.int('nan%d' % k) for k in (1, 2, 3, 4)]
        self.nffn = self.int('nffn'); self.nnnu = self.int('nnnu'); self.mflx = self.int('mflx')
        self.helm_loaded = False
        return self

    def load_helm(self, path=None):
        """Helmholtz table for screening / self-heating (XNet reads helm_table.dat from the run directory);
        default: helm_table.bin or helm_table.dat next to the network directory."""
        if self.helm_loaded:
            return
        if path is None:
            root = os.path.dirname(os.path.abspath(self.data_dir.rstrip('/')))
            cands = [os.path.join(root, n) for n in ('helm_table.bin', 'helm_table.dat')]
            path = next((p for p in cands if os.path.exists(p)), cands[0])
        _chk(lib().xb_net_load_helm(self.h, path.encode()))
        self.helm_loaded = True

    def names(self):
        """Species names in network order (the sunet file of the data directory, xnet_data.F90:285-324)."""
        with open(os.path.join(self.data_dir, 'sunet')) as f:
            nm = [ln[:5] for ln in f.read().split('\n') if ln.strip()]
        if len(nm) != self.ny:
            raise XnetB200Error('sunet lists %d species, the network has %d' % (len(nm), self.ny))
        return nm

    def int(self, what):
        v = C.c_int32()
        _chk(lib().xb_net_int(self.h, what.encode(), C.byref(v)))
        return v.value

    def slu_table(self, what, heat=0):
        """structured-LU index tables (xb_slu_table)"""
        n = C.c_int64()
        _chk(lib().xb_slu_table(self.h, heat, what.encode(), None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), np.int32)
        _chk(lib().xb_slu_table(self.h, heat, what.encode(), _pi(out), out.size, C.byref(n)))
        return out[:n.value]

    def itable(self, what):
        n = C.c_int64()
        _chk(lib().xb_net_itable(self.h, what.encode(), None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), np.int32)
        _chk(lib().xb_net_itable(self.h, what.encode(), _pi(out), out.size, C.byref(n)))
        return out[:n.value]

    def dtable(self, what):
        n = C.c_int64()
        _chk(lib().xb_net_dtable(self.h, what.encode(), None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1))
        _chk(lib().xb_net_dtable(self.h, what.encode(), _pd(out), out.size, C.byref(n)))
        return out[:n.value]

    def __del__(self):
        try:
            if self.h:
                lib().xb_net_free(self.h); self.h = None
        except Exception:
            pass


class Fixture(dict):
    """One (trajectory, initial abundance) input pair with its start-of-run conditions."""


def load_fixture(net, inab_path, th_path):
    """read_thermo_file + load_initial_abundances + t9rhofind(tstart) through the C ABI."""
    l = lib()
    nh = C.c_int32(); hdr = np.zeros(3)
    _chk(l.xb_read_thermo(th_path.encode(), 0, C.byref(nh), None, None, None, None, _pd(hdr)))
    th = np.zeros(nh.value); t9h = np.zeros(nh.value); rhoh = np.zeros(nh.value); yeh = np.zeros(nh.value)
    _chk(l.xb_read_thermo(th_path.encode(), nh.value, C.byref(nh), _pd(th), _pd(t9h), _pd(rhoh), _pd(yeh), _pd(hdr)))
    y = np.zeros(net.ny); scal = np.zeros(4)
    _chk(l.xb_read_inab(net.h, inab_path.encode(), _pd(y), _pd(scal)))
    nstart = C.c_int32(); out4 = np.zeros(4)
    _chk(l.xb_zone_start(net.h, nh.value, _pd(th), _pd(t9h), _pd(rhoh), _pd(yeh), hdr[0], _pd(y), _pd(scal),
                         C.byref(nstart), _pd(out4)))
    return Fixture(nh=nh.value, th=th, t9h=t9h, rhoh=rhoh, yeh=yeh, tstart=hdr[0], tstop=hdr[1], tdelstart=hdr[2],
                   nstart=nstart.value, t9start=out4[0], rhostart=out4[1], yestart=out4[2], ystart=y,
                   xext=scal[1], aext=scal[2], zext=scal[3])


def _pin(a):
    """copy of a numpy array in page-locked host memory (torch is the allocator; the product path itself does not use it)"""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).clone().pin_memory()
    return t.numpy()


def build_zone_set(fixtures, nzone=None, zfix=None, rho_scale=None, restart=None, pinned=False):
    """Packs fixtures into the xb_zones layout; zone z uses fixture z mod K unless zfix is given.  pinned: the arrays
    are placed in page-locked memory (the host buffers of an end-to-end xb_evolve call)."""
    K = len(fixtures)
    nzone = nzone or K
    a = dict(
        nh=_i([f['nh'] for f in fixtures]),
        off=_i(np.concatenate([[0], np.cumsum([f['nh'] for f in fixtures])[:-1]])),
        th=_d(np.concatenate([f['th'] for f in fixtures])), t9h=_d(np.concatenate([f['t9h'] for f in fixtures])),
        rhoh=_d(np.concatenate([f['rhoh'] for f in fixtures])),
        tstart=_d([f['tstart'] for f in fixtures]), tstop=_d([f['tstop'] for f in fixtures]),
        tdelstart=_d([f['tdelstart'] for f in fixtures]), nstart=_i([f['nstart'] for f in fixtures]),
        t9start=_d([f['t9start'] for f in fixtures]), rhostart=_d([f['rhostart'] for f in fixtures]),
        yestart=_d([f['yestart'] for f in fixtures]), ystart=_d(np.stack([f['ystart'] for f in fixtures])),
        xext=_d([f['xext'] for f in fixtures]), aext=_d([f['aext'] for f in fixtures]),
        zext=_d([f['zext'] for f in fixtures]),
        zfix=_i(np.arange(nzone) % K) if zfix is None else _i(zfix))
    if rho_scale is not None:
        a['rho_scale'] = _d(rho_scale)
    if restart is not None:
        for nm in ('t', 'tdel', 't9', 'rho', 't9o', 'y', 'yo'):
            a['rs_' + nm] = _d(restart[nm])
        a['rs_kstep0'] = _i(restart['kstep0'])
    if pinned:
        a = {k: _pin(v) for k, v in a.items()}
    z = XbZones()
    z.nfix = K; z.nzone = nzone
    for nm, tp in XbZones._fields_:
        if nm in a:
            setattr(z, nm, _pi(a[nm]) if tp is c_ip else _pd(a[nm]))
    z._keep = a
    return z


def _alloc_out(nzone, ny, hist_cap, pinned=False, mflx=0):
    def arr(shape, dt):
        if pinned:
            import torch
            t = torch.zeros(shape, dtype=torch.float64 if dt == np.float64 else torch.int32).pin_memory()
            return t.numpy()
        return np.zeros(shape, dt)
    r = dict(y=arr((nzone, ny), np.float64), t=arr(nzone, np.float64), t9=arr(nzone, np.float64),
             rho=arr(nzone, np.float64), tdel=arr(nzone, np.float64), kstep=arr(nzone, np.int32),
             ktot=arr((nzone, 5), np.int32), status=arr(nzone, np.int32))
    if hist_cap > 0:
        r.update(hist_n=np.zeros(nzone, np.int32), hist_kstep=np.zeros((nzone, hist_cap), np.int32),
                 hist_kmon=np.zeros((nzone, hist_cap, 2), np.int32))
        for nm in ('hist_t', 'hist_t9', 'hist_rho', 'hist_tdel', 'hist_edot', 'hist_sqnu'):
            r[nm] = np.zeros((nzone, hist_cap))
        r['hist_y'] = np.zeros((nzone, hist_cap, ny))
        if mflx > 0:
            r['hist_flx'] = np.zeros((nzone, hist_cap, mflx))
    o = XbOut()
    o.hist_cap = hist_cap
    for nm, tp in XbOut._fields_:
        if nm in r:
            setattr(o, nm, _pi(r[nm]) if tp is c_ip else _pd(r[nm]))
    return r, o


class Context:
    """Device-resident state for `nzevolve` zones (what read_jacobian_data + be_init allocate)."""
    def __init__(self, net, controls, nzevolve=1, device=0, arith=0, solver=0, **override):
        """arith: 0 = fast, 1 = strict (bit-reproducible).  solver (fused evolve, fast arithmetic only): 0 = dense
        partial-pivoting LU (DGETRF order), 1 = structured LU (light particles last, skyline envelope, threshold-checked
        elimination without interchanges, partial pivoting inside the light-particle border, dense fallback)."""
        self.net = net
        self.ctl = controls.to_struct(arith=arith, solver=solver, **override)
        if self.ctl.iscrn > 0 or self.ctl.iheat > 0:
            net.load_helm()
        self.h = C.c_void_p()
        self.nzevolve = nzevolve
        _chk(lib().xb_ctx_create(net.h, C.byref(self.ctl), nzevolve, device, C.byref(self.h)))
        cfg = np.zeros(6, np.int32)
        _chk(lib().xb_launch_config(self.h, _pi(cfg)))
        self.launch_config = dict(grid=int(cfg[0]), threads=int(cfg[1]), smem_bytes=int(cfg[2]),
                                  jac_in_smem=int(cfg[3]), nsm=int(cfg[4]), msize=int(cfg[5]))
        self.msize = int(cfg[5])

    def close(self):
        if self.h:
            lib().xb_ctx_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- module procedures (host arrays in, host arrays out) ----
    def cross_sect(self, yt, t9t, rhot, mask=None, xext=None, aext=None, zext=None):
        yt = _d(yt); nz = yt.shape[0]
        t9t = _d(t9t); rhot = _d(rhot); yet = np.zeros(nz)
        m = None if mask is None else _i(mask)
        ex = [None if v is None else _d(v) for v in (xext, aext, zext)]
        _chk(lib().xb_cross_sect(self.h, 1, nz, _pi(m), _pd(yt), _pd(t9t), _pd(rhot), _pd(ex[0]), _pd(ex[1]), _pd(ex[2]), _pd(yet)))
        return yet

    def csect(self, nz):
        out = np.zeros((nz, sum(self.net.nreac)))
        _chk(lib().xb_get_csect(self.h, 1, nz, _pd(out)))
        return out

    def yderiv(self, yt, mask=None):
        yt = _d(yt); nz = yt.shape[0]; ydot = np.zeros_like(yt)
        m = None if mask is None else _i(mask)
        _chk(lib().xb_yderiv(self.h, 1, nz, _pi(m), _pd(yt), _pd(ydot)))
        return ydot

    def jacobian_build(self, nz, diag=None, mult=None, mask=None):
        d = None if diag is None else _d(diag); mu = None if mult is None else _d(mult)
        m = None if mask is None else _i(mask)
        _chk(lib().xb_jacobian_build(self.h, 1, nz, _pi(m), _pd(d), _pd(mu)))

    def jacobian_scale(self, nz, diag=None, mult=None, mask=None):
        """jacobian_scale: rescales the dydotdy of the last jacobian_build (no re-assembly)"""
        d = None if diag is None else _d(diag); mu = None if mult is None else _d(mult)
        m = None if mask is None else _i(mask)
        _chk(lib().xb_jacobian_scale(self.h, 1, nz, _pi(m), _pd(d), _pd(mu)))

    def jac(self, nz):
        """jac[z, col, row] as it lies in memory (column-major per zone)."""
        out = np.zeros((nz, self.msize, self.msize))
        _chk(lib().xb_get_jac(self.h, 1, nz, _pd(out)))
        return out

    def set_jac(self, jac_colmajor):
        j = _d(jac_colmajor)
        _chk(lib().xb_set_jac(self.h, 1, j.shape[0], _pd(j)))

    def jacobian_decomp(self, nz, mask=None):
        info = np.zeros(nz, np.int32)
        m = None if mask is None else _i(mask)
        _chk(lib().xb_jacobian_decomp(self.h, 1, nz, _pi(m), _pi(info)))
        return info

    def indx(self, nz):
        out = np.zeros((nz, self.msize), np.int32)
        _chk(lib().xb_get_indx(self.h, 1, nz, _pi(out)))
        return out

    def jacobian_bksub(self, yrhs, mask=None, t9rhs=None):
        """jacobian_bksub(kstep,yrhs,dy,t9rhs,dt9,mask_in) (xnet_jacobian_dense.F90:618).  Returns dy, or (dy, dt9)
        when t9rhs is given (self-heating, msize = ny+1)."""
        yrhs = _d(yrhs); nz = yrhs.shape[0]; dy = np.zeros_like(yrhs)
        m = None if mask is None else _i(mask)
        if t9rhs is None:
            _chk(lib().xb_jacobian_bksub(self.h, 1, nz, _pi(m), _pd(yrhs), None, _pd(dy), None))
            return dy
        t9rhs = _d(t9rhs); dt9 = np.zeros(nz)
        _chk(lib().xb_jacobian_bksub(self.h, 1, nz, _pi(m), _pd(yrhs), _pd(t9rhs), _pd(dy), _pd(dt9)))
        return dy, dt9

    def dcsect(self, nz):
        out = np.zeros((nz, sum(self.net.nreac)))
        _chk(lib().xb_get_dcsect(self.h, 1, nz, _pd(out)))
        return out

    def eos(self, nz):
        """(cv, etae, detaedt9) of the last cross_sect and t9dot of the last yderiv, per zone."""
        cv = np.zeros(nz); etae = np.zeros(nz); deta = np.zeros(nz); t9dot = np.zeros(nz)
        _chk(lib().xb_get_eos(self.h, 1, nz, _pd(cv), _pd(etae), _pd(deta), _pd(t9dot)))
        return cv, etae, deta, t9dot

    def step_be(self, y, yt, tdel, t9, t9t, rhot, inr=None, nh=None, kstep=1):
        """step_be(kstep,inr) (xnet_integrate_be.F90:233): one launch for the whole Newton loop.
        Returns dict(inr, yt, t9t, ydot, t9dot, kcount[nz,5]); inputs are not modified."""
        y = _d(y); nz = y.shape[0]
        yt = _d(yt).copy(); t9t = _d(t9t).copy(); tdel = _d(tdel); t9 = _d(t9); rhot = _d(rhot)
        inr = np.zeros(nz, np.int32) if inr is None else _i(inr).copy()
        nhh = None if nh is None else _i(nh)
        ydot = np.zeros_like(y); t9dot = np.zeros(nz); kc = np.zeros((nz, 5), np.int32)
        _chk(lib().xb_step_be(self.h, 1, nz, kstep, _pi(inr), _pd(y), _pd(yt), _pd(tdel), _pd(t9), _pd(t9t), _pd(rhot),
                              _pi(nhh), _pd(ydot), _pd(t9dot), _pi(kc)))
        return dict(inr=inr, yt=yt, t9t=t9t, ydot=ydot, t9dot=t9dot, kcount=kc)

    # ---- fused evolve ----
    def full_net(self, zones, hist_cap=0, pinned=False, flux=False, check=True):
        """xb_evolve: host buffers in, host buffers out (H2D + kernel + D2H).  flux=True adds hist_flx, the
        matched-pair reaction fluxes of every ts record (xnet_flux.F90:67-150)."""
        res, o = _alloc_out(zones.nzone, self.net.ny, hist_cap, pinned, self.net.mflx if flux else 0)
        _chk(lib().xb_set_flux_output(self.h, 1 if flux else 0))
        _chk(lib().xb_evolve(self.h, C.byref(zones), C.byref(o)))
        if check:
            self.check_status(res)
        return res

    @staticmethod
    def check_status(res):
        """The reference stops the run when a zone cannot continue ('Evolution failed to converge',
        xnet_evolve.F90:226-237: its = 1 after ten time-step retries).  status: -1 = reached tstop, 0 = still evolving
        when kstmx steps were taken (callers that cap kstmx on purpose), 2 = failed to converge."""
        bad = np.nonzero(res['status'] > 0)[0]
        if bad.size:
            raise XnetB200Error('evolution failed to converge in %d zone(s), first: %d' % (bad.size, int(bad[0])))

    def flux(self, yt=None, nz=None, mask=None):
        """flux(mask_in) (xnet_flux.F90:67): fluxes [nz, mflx] from the device-resident rates and yt."""
        if yt is not None:
            yt = _d(yt); nz = yt.shape[0]
        flx = np.zeros((nz, max(self.net.mflx, 1)))
        m = None if mask is None else _i(mask)
        _chk(lib().xb_flux(self.h, 1, nz, _pi(m), _pd(yt), _pd(flx)))
        return flx[:, :self.net.mflx]

    def upload(self, zones, hist_cap=0, flux=False):
        self._nzone = zones.nzone; self._hist_cap = hist_cap; self._flux = flux
        _chk(lib().xb_set_flux_output(self.h, 1 if flux else 0))
        _chk(lib().xb_zones_upload(self.h, C.byref(zones), hist_cap))

    def run_resident(self):
        ms = C.c_float()
        _chk(lib().xb_evolve_resident(self.h, C.byref(ms)))
        return ms.value

    def download(self, pinned=False):
        res, o = _alloc_out(self._nzone, self.net.ny, self._hist_cap, pinned, self.net.mflx if getattr(self, '_flux', False) else 0)
        _chk(lib().xb_results_download(self.h, C.byref(o)))
        return res

    def phase_cycles(self):
        c = np.zeros(24)
        _chk(lib().xb_phase_cycles(self.h, _pd(c)))
        names = ('rates', 'yderiv', 'jac', 'lu', 'bksub', 'other', 'total', 'zones',
                 'lu_panel', 'lu_swap_trsm', 'lu_gemm', 'spare') + tuple('dbg%d' % k for k in range(12))
        return dict(zip(names, c))

    def solver_info(self):
        """None with the dense solver, else the structured solver's sizes (xb_solver_info)."""
        v = np.zeros(8)
        _chk(lib().xb_solver_info(self.h, _pd(v)))
        if v[0] == 0:
            return None
        return dict(msize=int(v[1]), skyline_doubles=int(v[2]), flops=float(v[3]), zones_per_sm=int(v[4]), border=int(v[5]),
                    lu_warps=int(v[6]), rates_cached=int(v[7]))

    def slu_solve(self, a, b, mode=0):
        """xb_slu_solve: a [nz, n, n] as [zone, col, row] (column-major per zone, equation order), b [nz, n] -> (x, flags)"""
        a = _d(a); b = _d(b); nz = a.shape[0]
        x = np.zeros_like(b); fl = np.zeros(nz, np.int32)
        _chk(lib().xb_slu_solve(self.h, nz, _pd(a), _pd(b), _pd(x), _pi(fl), mode))
        return x, fl

    @property
    def kernel_name(self):
        return 'evolve_slu_kernel' if self.solver_info() else ('evolve_kernel<strict>' if self.ctl.arith else 'evolve_kernel<fast>')

    def last_kernel_ms(self):
        """CUDA-event duration of the last kernel launched by this context (module-level drop-in or evolve)."""
        ms = C.c_float()
        _chk(lib().xb_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def copy_bytes(self):
        a = C.c_int64(); b = C.c_int64()
        _chk(lib().xb_last_copy_bytes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value


def measure_fp64_peak(device=0):
    v = C.c_double()
    _chk(lib().xb_measure_fp64_peak(device, C.byref(v)))
    return v.value
