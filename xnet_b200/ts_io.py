"""XNet binary time-series files (`ts_<zone>`): writer binding, reader and comparator.

Format = the reference's Write statements (net.F90:288-301 header records, xnet_output.F90:78-79 one record per
output step: kstep, t, t9, rho, sqnu, tdel, edot, y(ny), flx(mflx)), Fortran sequential-unformatted with 4-byte
record markers.  The writer is the C++ side of the library (xb_ts_write); reading and comparing is host tooling
in Python like the reference's own (tools/python/read_ts_file.py, find_difference.py, test/compare_ts.py).

Note on the reference reader: read_ts_file.py reads five float64 after kstep (t, t9, rho, tdel, edot) although the
writer emits six (sqnu sits between rho and tdel), so with current ts files its `timestep` is sqnu, its `edot` is
tdel and its abundances are shifted by one slot.  `read_ts` below follows the writer; `read_ts(..., legacy=True)`
reproduces the reference reader's slicing for anyone who needs to compare against numbers produced with it.
"""
import ctypes as C
import os

import numpy as np

from . import api


def name_ordered(base, n, nmax):
    """base + n zero-padded to the number of digits of nmax (xnet_util.F90:341-370)."""
    return base + str(n).zfill(len(str(int(nmax))))


def write_ts(path, net, controls, fixture, hist, zone=0, descript=('', '', ''), data_desc='', inab_file='', abund_desc='',
             thermo_file='', thermo_desc=''):
    """Writes the records of zone `zone` of a history dict (Context.full_net(..., hist_cap>0[, flux=True])) to `path`.
    Without hist_flx the flux block is written as zeros (the reference's initial record is zero too)."""
    n = int(hist['hist_n'][zone])
    cap = hist['hist_kstep'].shape[1]
    if n > cap:
        raise ValueError('history overflow: %d records, capacity %d' % (n, cap))
    ctl = controls.to_struct() if hasattr(controls, 'to_struct') else controls
    h = api.XbTsHeader()
    enc = [str(s).encode() for s in descript]
    for i in range(3):
        h.descript[i] = enc[i]
    h.data_desc = str(data_desc).encode(); h.ctl = C.pointer(ctl)
    h.inab_file = str(inab_file).encode(); h.abund_desc = str(abund_desc).encode()
    h.thermo_file = str(thermo_file).encode(); h.thermo_desc = str(thermo_desc).encode()
    zz = api._d(net.dtable('zz')); aa = api._d(net.dtable('aa'))
    h.ny = net.ny; h.zz = api._pd(zz); h.aa = api._pd(aa)
    mflx = net.mflx
    orig = api._i(net.itable('ifl_orig')) if mflx else None; term = api._i(net.itable('ifl_term')) if mflx else None
    h.mflx = mflx; h.ifl_orig = api._pi(orig); h.ifl_term = api._pi(term)
    col = {k: np.ascontiguousarray(hist['hist_' + k][zone, :n]) for k in ('kstep', 't', 't9', 'rho', 'sqnu', 'tdel', 'edot', 'y')}
    flx = np.ascontiguousarray(hist['hist_flx'][zone, :n]) if 'hist_flx' in hist and mflx else None
    api._chk(api.lib().xb_ts_write(os.fsencode(path), C.byref(h), n, api._pi(api._i(col['kstep'])), api._pd(col['t']),
                                   api._pd(col['t9']), api._pd(col['rho']), api._pd(col['sqnu']), api._pd(col['tdel']),
                                   api._pd(col['edot']), api._pd(col['y']), api._pd(flx)))
    return n


def write_ev(path, net, hist, zone=0, names=()):
    """ASCII evolution file ev_<zone> (itsout >= 2; net.F90:303-314, xnet_output.F90:72,83-85): one line per record with
    the mass fractions of the species in `names` (the control file's output nuclei; 5-character names as in sunet,
    matched after stripping blanks)."""
    n = int(hist['hist_n'][zone])
    cap = hist['hist_kstep'].shape[1]
    if n > cap:
        raise ValueError('history overflow: %d records, capacity %d' % (n, cap))
    aa = net.dtable('aa')
    known = [s.strip().lower() for s in net.names()] if hasattr(net, 'names') else None
    idx = []
    for nm in names:
        if known is None:
            raise ValueError('species names are not available for this network')
        idx.append(known.index(nm.strip().lower()))
    y = hist['hist_y'][zone, :n]
    xout = np.ascontiguousarray(y[:, idx] * aa[idx]) if idx else None
    names5 = ''.join(nm.strip().rjust(5)[:5] for nm in names).encode()
    col = {k: np.ascontiguousarray(hist['hist_' + k][zone, :n]) for k in ('kstep', 't', 't9', 'rho', 'edot', 'sqnu', 'tdel')}
    kmon = np.ascontiguousarray(hist['hist_kmon'][zone, :n])
    api._chk(api.lib().xb_ev_write(os.fsencode(path), n, api._pi(api._i(col['kstep'])), api._pd(col['t']), api._pd(col['t9']),
                                   api._pd(col['rho']), api._pd(col['edot']), api._pd(col['sqnu']), api._pd(col['tdel']),
                                   len(idx), names5, api._pd(xout), api._pi(api._i(kmon))))
    return n


def _records(buf):
    """Splits a Fortran sequential-unformatted byte string into payloads, checking both record markers."""
    out = []; p = 0
    while p < len(buf):
        if p + 4 > len(buf):
            raise ValueError('truncated record marker at byte %d' % p)
        n = int(np.frombuffer(buf, '<i4', 1, p)[0])
        if n < 0 or p + 8 + n > len(buf):
            raise ValueError('truncated record at byte %d' % p)
        if int(np.frombuffer(buf, '<i4', 1, p + 4 + n)[0]) != n:
            raise ValueError('record markers disagree at byte %d' % p)
        out.append(buf[p + 4:p + 4 + n]); p += 8 + n
    return out


def read_ts(path, legacy=False):
    """Reads a ts file.  Returns a dict: descript[3], data_desc, controls (dict), inab_file, abund_desc, thermo_file,
    thermo_desc, ny, zz, aa, mflx, ifl_orig, ifl_term and per record kstep, t, t9, rho, sqnu, tdel, edot, y [n,ny]
    (abundances), xmf = y*aa (mass fractions), flx [n,mflx]."""
    with open(path, 'rb') as f:
        rec = _records(f.read())
    if len(rec) < 6:
        raise ValueError('not a ts file: %d records' % len(rec))
    txt = lambda b: b.decode('latin-1').rstrip()
    r = dict(descript=[txt(rec[0][80 * i:80 * i + 80]) for i in range(3)], data_desc=txt(rec[0][240:320]))
    i5 = np.frombuffer(rec[1], '<i4', 5, 0); d6 = np.frombuffer(rec[1], '<f8', 6, 20); i2 = np.frombuffer(rec[1], '<i4', 2, 68)
    r['controls'] = dict(zip(('kstmx', 'kitmx', 'iweak0', 'iscrn', 'iconvc'), (int(v) for v in i5)))
    r['controls'].update(zip(('changemx', 'tolm', 'tolc', 'yacc', 'ymin', 'tdel_maxmult'), (float(v) for v in d6)))
    r['controls'].update(iheat=int(i2[0]), isolv=int(i2[1]))
    r['inab_file'] = txt(rec[2][:80]); r['abund_desc'] = txt(rec[2][80:160])
    r['thermo_file'] = txt(rec[3][:80]); r['thermo_desc'] = txt(rec[3][80:160])
    ny = int(np.frombuffer(rec[4], '<i4', 1)[0])
    r['ny'] = ny; r['zz'] = np.frombuffer(rec[4], '<f8', ny, 4).copy(); r['aa'] = np.frombuffer(rec[4], '<f8', ny, 4 + 8 * ny).copy()
    mflx = int(np.frombuffer(rec[5], '<i4', 1)[0])
    r['mflx'] = mflx
    r['ifl_orig'] = np.frombuffer(rec[5], '<i4', mflx, 4).copy(); r['ifl_term'] = np.frombuffer(rec[5], '<i4', mflx, 4 + 4 * mflx).copy()
    steps = rec[6:]
    n = len(steps); want = 4 + 8 * (6 + ny + mflx)
    for k, s in enumerate(steps):
        if len(s) != want:
            raise ValueError('step record %d has %d bytes, expected %d' % (k, len(s), want))
    body = np.frombuffer(b''.join(s[4:] for s in steps), '<f8').reshape(n, 6 + ny + mflx) if n else np.zeros((0, 6 + ny + mflx))
    r['kstep'] = np.array([np.frombuffer(s, '<i4', 1)[0] for s in steps], np.int32)
    if legacy:   # slicing of tools/python/read_ts_file.py (five scalars): see the module docstring
        names = ('t', 't9', 'rho', 'tdel', 'edot'); off = 5
        r['sqnu'] = np.zeros(n)
    else:
        names = ('t', 't9', 'rho', 'sqnu', 'tdel', 'edot'); off = 6
    for j, nm in enumerate(names):
        r[nm] = body[:, j].copy()
    r['y'] = body[:, off:off + ny].copy()
    r['xmf'] = r['y'] * r['aa']
    r['flx'] = body[:, 6 + ny:6 + ny + mflx].copy()
    return r


def compare_ts(file1, file2, rtol=1e-10, floor=1e-20, steps=True):
    """The north-star parity check between two ts files: identical accepted-step sequence (kstep, and t and tdel to
    `rtol`) and abundances Y > floor within `rtol` relative at every step; plus the numbers test/compare_ts.py prints
    for the final record (find_difference.compare_final): |dX|, |dX|/|X1| per species and their 2-norms, T9 and
    density differences.  Returns a dict with `ok` and the measured maxima; raises nothing on mismatch."""
    a = file1 if isinstance(file1, dict) else read_ts(file1)
    b = file2 if isinstance(file2, dict) else read_ts(file2)
    res = dict(nstep1=len(a['kstep']), nstep2=len(b['kstep']))
    same_shape = a['ny'] == b['ny'] and res['nstep1'] == res['nstep2'] and res['nstep1'] > 0
    res['same_species'] = a['ny'] == b['ny'] and np.array_equal(a['zz'], b['zz']) and np.array_equal(a['aa'], b['aa'])
    ok = res['same_species'] and res['nstep1'] > 0 and res['nstep2'] > 0
    if steps:
        res['same_step_count'] = res['nstep1'] == res['nstep2']
        ok = ok and same_shape
        if same_shape:
            rel = lambda x, y: float(np.max(np.abs(x - y) / np.maximum(np.abs(x), np.finfo(float).tiny)))
            res['kstep_equal'] = bool(np.array_equal(a['kstep'], b['kstep']))
            res['t_rel'] = rel(a['t'], b['t']); res['tdel_rel'] = rel(a['tdel'][1:], b['tdel'][1:]) if res['nstep1'] > 1 else 0.0
            m = a['y'] > floor
            res['y_rel_max'] = float(np.max(np.where(m, np.abs(b['y'] - a['y']) / np.where(m, a['y'], 1.0), 0.0)))
            ok = ok and res['kstep_equal'] and res['t_rel'] <= rtol and res['tdel_rel'] <= rtol and res['y_rel_max'] <= rtol
    if res['same_species'] and res['nstep1'] > 0 and res['nstep2'] > 0:
        x1 = a['xmf'][-1]; x2 = b['xmf'][-1]
        aerr = np.abs(x2 - x1)
        with np.errstate(divide='ignore', invalid='ignore'):
            res['final_rerr'] = aerr / np.abs(x1)
        res['final_aerr'] = aerr
        res['final_aerr_norm'] = float(np.linalg.norm(aerr)); res['final_rerr_norm'] = float(np.linalg.norm(aerr) / np.linalg.norm(x1))
        res['final_t9_rerr'] = float(abs(b['t9'][-1] - a['t9'][-1]) / abs(a['t9'][-1]))
        res['final_rho_rerr'] = float(abs(b['rho'][-1] - a['rho'][-1]) / abs(a['rho'][-1]))
        if not steps:
            m = x1 > floor
            ok = ok and float(np.max(np.where(m, aerr / np.where(m, x1, 1.0), 0.0))) <= rtol
    res['ok'] = bool(ok)
    return res
