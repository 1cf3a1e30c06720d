"""Oracle (TEST INFRASTRUCTURE ONLY) -- Python restatement of XNet's input readers and network
pre-processing.  Product code must never import this module.

Follows, in the reference tree (/root/reference/source):
  read_sunet / read_netwinv / read_nuclear_data   xnet_data.F90:285-513
  net_preprocess (netsu -> reaction tables + species->reaction maps)  xnet_preprocess.F90:505-1011
  read_reaction_data (map expansion n_n1..n_n4, iffn)                 xnet_data.F90:654-956
  read_ffn_data_table / read_ffn_data_logft                           xnet_ffn.F90:73-137
  read_controls                                                       xnet_controls.F90:164-396
  read_thermo_file / load_initial_abundances / read_inab_file         model_input_ascii.F90:5-350
  y_moment                                                            xnet_abundances.F90:93-185

Parity pins: the shipped `net_desc` files (species count, la/le per REACLIB chapter, nreac, nan,
nffn/nnu) are reproduced exactly (tests/test_oracle_loader.py).
All indices returned here are 1-based as in the Fortran (0 = blank slot).
"""
import os
import re
import numpy as np

NCHAP = 11
NNUCR = [2, 3, 4, 3, 4, 5, 6, 4, 5, 6, 5]        # xnet_preprocess.F90:21
NREACTANT = [1, 1, 1, 2, 2, 2, 2, 3, 3, 4, 1]    # xnet_preprocess.F90:22
NG = 24

# xnet_preprocess.F90:534-556 (171 labels, order matters: first match wins)
DESC_KNOWN = [
    '  ec', 'bet+', 'betp', 'bet-', 'betm', 'btyk', ' bec', 'bkmo',
    'mo92', ' ffn', ' lmp', 'fkth', 'cf88', 'wies', 'laur', 'bb92',
    'baka', 'rolf', 'wawo', 'mafo', 'ms01', ' wag', 'bb93', 'ca85',
    'ornl', ' wfh', 'rath', 'rpsm', 'hg94', 'ro95', 'wien', 'iucf',
    'smwa', 'smmo', 'wi2p', 'aexp', 'pexp', 'bex+', 'bex-', 'bqa+',
    'bhi+', 'bec ', 'bklq', 'nuba', 'bede', 'bqa+', 'gamm', 'ja01',
    'nosm', 'nacr', 'tang', 'ta04', 'ch04', 'ku02', 'il04', 'go00',
    'most', 'il03', 'de04', 'da03', 'vanc', 'il01', 'ha00', 'ku92',
    'ro04', 'sc83', 'he98', 'po00', 'il00', 'he95', 'sc98', 'vo00',
    'il99', 'beau', 'da04', 'ca00', 'da77', 'bl03', 'ch01', 'fynb',
    'De99', 'bu96', 'ba00', 'Ha96', 'sh03', 'tr04', 'sc05', 'nfis',
    'GW95', 'ch05', 're98', 'nb03', 'mo03', 'wc07', 'an06', 'ka02',
    'gl07', 'GG90', 'lg06', 'im05', 'GG92', 'ct07', 'ga00', 'ww02',
    'wh87', 'dh03', 'ha01', 'cb01', 'HF99', 'hg95', 'pg06', 'jm06',
    'sb05', 'SM93', 'SM86', 'thra', 'fy05', 'ha04', 'pt05', 'bk91',
    'bk92', 'kg91', 'SM91', 'iwam', ' nen', 'nebp', 'lznu', 'lzan',
    'ths8', 'wc12', 'nk06', 'mb11', 'jz10', 'cd08', 'nw00', 'il10',
    'li10', 'chw0', 'co10', 'kd02', 'st08', 'pkrF', 'br10', 'hi09',
    'ld11', 'hd08', 'ua08', 'ol11', 'ls09', 'ia08', 'dc11', 'mb07',
    'wc17', 'mo97', 'ks03', 'nac2', 'mp17', 'cb09', 'li12', 'ma10',
    'mm11', 'si13', 'sa12', 'ks12', 'hg12', 'rk12', 'ac12', 'gl12',
    'mv09', 'lhnn', 'lhna']
assert len(DESC_KNOWN) == 171

# xnet_constants.F90:44-56 (CODATA 2014)
AVN = 6.022140857e+23
EPMEV = 1.6021766208e-06
M_E = 0.5109989461
M_N = 939.5654133
M_P = 938.2720813
CLT = 2.99792458e+10
AMU = 1.0 / (AVN * EPMEV)
M_U = AMU * CLT ** 2
THBIM1 = 1.58
BIP1 = 1.86
FIVE3RD = 5.0 / 3.0


def _ffloat(s):
    """Fortran formatted real field -> float (blank = 0, D exponent allowed)."""
    s = s.strip()
    if not s:
        return 0.0
    return float(s.replace('D', 'E').replace('d', 'e'))


def _ld_tokens(text):
    """Tokens of a Fortran list-directed record stream (comma/blank separated)."""
    return text.replace(',', ' ').split()


def read_sunet(data_dir):
    # xnet_data.F90:285-324 : one (a5) name per line, line count = ny
    with open(os.path.join(data_dir, 'sunet')) as f:
        lines = f.read().split('\n')
    while lines and lines[-1] == '':
        lines.pop()
    return [(ln + '     ')[:5] for ln in lines]


def read_netwinv(data_dir):
    # xnet_data.F90:326-389
    with open(os.path.join(data_dir, 'netwinv')) as f:
        lines = f.read().split('\n')
    ny = int(lines[0][:5])
    grid = (lines[1] + ' ' * 72)[:72]
    it9i = np.array([int(grid[3 * j:3 * j + 3]) for j in range(NG)], dtype=np.int32)
    t9i = it9i.astype(np.float64)
    t9i[:NG - 1] = 0.01 * t9i[:NG - 1]
    t9i[NG - 1] = 0.1 * t9i[NG - 1]
    names = [(lines[2 + i] + '     ')[:5] for i in range(ny)]
    toks = _ld_tokens('\n'.join(lines[2 + ny:]))
    aa = np.zeros(ny); iz = np.zeros(ny, np.int32); inn = np.zeros(ny, np.int32)
    spin = np.zeros(ny); mex = np.zeros(ny); g = np.zeros((ny, NG))
    p = 0
    for i in range(ny):
        nam = toks[p]; aa[i] = float(toks[p + 1]); iz[i] = int(toks[p + 2]); inn[i] = int(toks[p + 3])
        spin[i] = float(toks[p + 4]); mex[i] = float(toks[p + 5]); p += 6
        g[i, :] = [float(x) for x in toks[p:p + NG]]; p += NG
        if nam.strip() != names[i].strip():
            raise ValueError('netwinv data entry name does not match header for inuc=%d' % (i + 1))
    return dict(ny=ny, it9i=it9i, t9i=t9i, names=names, aa=aa, iz=iz, inn=inn, spin=spin, mex=mex, g=g)


def read_nuclear_data(data_dir):
    """xnet_data.F90:391-513 (+ format_nuclear_data consistency checks :540-585)."""
    sun = read_sunet(data_dir)
    w = read_netwinv(data_dir)
    ny = w['ny']
    if len(sun) != ny:
        raise ValueError('netwinv /= sunet')
    for a, b in zip(sun, w['names']):
        if a.strip() != b.strip():
            raise ValueError('netwinv /= sunet')
    zz = w['iz'].astype(np.float64)
    nn = w['inn'].astype(np.float64)
    ineut = iprot = 0
    for i in range(ny):
        if w['iz'][i] == 0 and w['inn'][i] == 1:
            ineut = i + 1
        if w['iz'][i] == 1 and w['inn'][i] == 0:
            iprot = i + 1
    mex = w['mex']
    mex_n = mex[ineut - 1] if ineut > 0 else M_N - M_U
    mex_p = mex[iprot - 1] if iprot > 0 else M_P + M_E - M_U
    be = mex_n * nn + mex_p * zz - mex
    desc = ''
    p = os.path.join(data_dir, 'net_desc')
    if os.path.exists(p):
        with open(p) as f:
            desc = f.readline().rstrip('\n')
    d = dict(w)
    d.update(zz=zz, nn=nn, be=be, mex_n=mex_n, mex_p=mex_p, ineut=ineut, iprot=iprot,
             izmax=int(w['iz'].max()), izmin=int(w['iz'].min()), data_desc=(desc + ' ' * 80)[:80],
             angm=2.0 * w['spin'] + 1.0)
    return d


def index_from_name(name, names_stripped, ny):
    # xnet_data.F90:54-82 : 0 blank, ny+1 not found
    s = name.strip()
    if not s:
        return 0
    try:
        return names_stripped.index(s) + 1
    except ValueError:
        return ny + 1


def net_preprocess(data_dir, nuc):
    """xnet_preprocess.F90:573-951 : parse netsu, classify, double-count factors, species->reaction
    registration.  Returns dict with per-class (c=0..3 for 1..4 reactants) arrays."""
    ny = nuc['ny']
    names_stripped = [n.strip() for n in nuc['names']]
    iz = nuc['iz']; be = nuc['be']; nn = nuc['nn']; zz = nuc['zz']
    mex_n = nuc['mex_n']; mex_p = nuc['mex_p']
    with open(os.path.join(data_dir, 'netsu')) as f:
        lines = f.read().split('\n')
    nffn = int(lines[0][0:5]); nnnu = int(lines[0][5:11])
    la = [1] * (NCHAP + 1); le = [0] * (NCHAP + 1)   # 1-based chapter index
    # per class storage: dict index n (1-based) -> record, as the Fortran overwrites slot n
    cls = [dict() for _ in range(4)]
    nk = [1, 1, 1, 1]
    n = 1
    krt = 1
    body = lines[1:]
    while body and body[-1].strip() == '':
        body.pop()
    pos = 0
    while pos + 3 <= len(body):
        hdr = (body[pos] + ' ' * 80)[:80]
        # (i2,3x,6a5,8x,a4,a1,a1,3x,1pe12.5)
        ktxt = hdr[0:2].strip()
        k = int(ktxt) if ktxt else 0
        nucname = [hdr[5 + 5 * j:10 + 5 * j] for j in range(6)]
        desc = hdr[43:47]; nr = hdr[47:48]; vw = hdr[48:49]
        q = _ffloat(hdr[52:64])
        c1 = (body[pos + 1] + ' ' * 52)[:52]
        c2 = (body[pos + 2] + ' ' * 52)[:52]
        p = [_ffloat(c1[13 * j:13 * j + 13]) for j in range(4)] + \
            [_ffloat(c2[13 * j:13 * j + 13]) for j in range(3)]
        pos += 3
        if k == 0:
            ni = [index_from_name(s, names_stripped, ny) for s in nucname]
            if any(x > ny for x in ni):
                continue
            try:
                idesc = DESC_KNOWN.index(desc) + 1
            except ValueError:
                idesc = 172
            if idesc == 1:
                iec = 1
            elif idesc in (10, 11):
                if iz[ni[0] - 1] > iz[ni[1] - 1] and iz[ni[0] - 1] != 1:
                    iec = 3
                else:
                    iec = 2
            elif (2 <= idesc <= 9) or (38 <= idesc <= 45) or (92 <= idesc <= 94) or idesc == 130 \
                    or idesc in (153, 154) or idesc == 157:
                iec = 4
            elif idesc in (125, 127):
                iec = 7
            elif idesc in (126, 128):
                iec = 8
            elif idesc == 170:
                iec = 9
            elif idesc == 171:
                iec = 10
            else:
                iec = 0
            if nr == ' ':
                ier = 0
            elif nr == 'w':
                ier = 0
            elif nr == 'n':
                ier = 1
            else:
                ier = 2
            iev = 1 if vw == 'v' else 0
            nrt = NREACTANT[krt - 1]
            ki = [0] * 6
            for i in range(6):
                if i + 1 > nrt:
                    ki[i] = 0 if nucname[i] == '     ' else 1
                else:
                    ki[i] = -1
            qtemp = 0.0
            for i in range(6):
                if i + 1 > nrt:
                    if nucname[i] != '     ':
                        j = ni[i] - 1
                        qtemp = qtemp + be[j] - mex_n * nn[j] - mex_p * zz[j]
                else:
                    j = ni[i] - 1
                    qtemp = qtemp - be[j] + mex_n * nn[j] + mex_p * zz[j]
            q = qtemp
            # identical reactants (xnet_preprocess.F90:729-738), 1-based i in the Fortran
            j = 1
            for i in range(2, nrt + 1):
                if nucname[i - 2] != nucname[i - 1]:
                    j = i
                else:
                    ki[j - 1] += ki[i - 1]
                    ki[i - 1] = 0
            # identical products (:741-750)
            kl = nrt + 2
            j = kl - 1
            for i in range(kl, 7):
                if nucname[i - 2] != nucname[i - 1]:
                    j = i
                else:
                    ki[j - 1] += ki[i - 1]
                    ki[i - 1] = 0
            if nrt == 1:
                r = 1.0
                nm = 5
            else:
                lo = 0
                for i in range(nrt):
                    if ki[i] < lo:
                        lo = ki[i]
                fact = 1
                for i in range(1, -lo + 1):
                    fact *= i
                r = 1.0 / float(fact)
                nm = 6
            rec = dict(n=[ni[i] for i in range(nm)] + [0] * (6 - nm),
                       an=[float(ki[i]) * r for i in range(nm)] + [0.0] * (6 - nm),
                       q=q, iwk=iec, ires=ier, irev=iev, desc=desc, rc=p)
            cls[nrt - 1][n] = rec
            nk[nrt - 1] += 1
            n += 1
        elif k == 1:
            krt = k; la[k] = 1; n = 1
        elif k in (4, 8, 10):
            krt = k; la[k] = 1; le[k - 1] = n - 1; n = 1
        elif k == 11:
            krt = k; la[k] = le[3] + 1; le[k - 1] = n - 1; n = le[3] + 1
        else:
            krt = k; la[k] = n; le[k - 1] = n - 1
    le[krt] = n - 1
    nreac = [nk[c] - 1 for c in range(4)]

    out = dict(nffn=nffn, nnnu=nnnu, nreac=nreac, chap_la=la[1:], chap_le=le[1:])
    for c in range(4):
        nrc = nreac[c]
        N = np.zeros((nrc, 6), np.int32); AN = np.zeros((nrc, 6)); RC = np.zeros((nrc, 7))
        Q = np.zeros(nrc); IWK = np.zeros(nrc, np.int32); IRES = np.zeros(nrc, np.int32)
        IREV = np.zeros(nrc, np.int32); DESC = [''] * nrc
        for kk in range(1, nrc + 1):
            rec = cls[c][kk]
            N[kk - 1] = rec['n']; AN[kk - 1] = rec['an']; RC[kk - 1] = rec['rc']; Q[kk - 1] = rec['q']
            IWK[kk - 1] = rec['iwk']; IRES[kk - 1] = rec['ires']; IREV[kk - 1] = rec['irev']
            DESC[kk - 1] = rec['desc']
        out['n%d' % (c + 1)] = N; out['an%d' % (c + 1)] = AN; out['rc%d' % (c + 1)] = RC
        out['q%d' % (c + 1)] = Q; out['iwk%d' % (c + 1)] = IWK; out['ires%d' % (c + 1)] = IRES
        out['irev%d' % (c + 1)] = IREV; out['desc%d' % (c + 1)] = DESC

    # species -> reaction registration (xnet_preprocess.F90:863-951)
    tiny = np.finfo(np.float64).tiny
    lsa = np.zeros((4, ny), np.int32); lse = np.zeros((4, ny), np.int32)
    nan = [0] * 4
    for c in range(4):
        N = out['n%d' % (c + 1)]; AN = out['an%d' % (c + 1)]
        a_list = []; mu_list = []
        # slot-valid mask per reaction: chapter j covers reactions la(j)..le(j), slots 1..nnucr(j)
        # a reaction index may (pathologically) be covered by several chapters; emulate loop order.
        chap_ranges = [(j, la[j], le[j]) for j in range(1, NCHAP + 1) if NREACTANT[j - 1] == c + 1]
        cnt = 0
        for i in range(1, ny + 1):
            lsa[c, i - 1] = cnt + 1
            for (j, a0, e0) in chap_ranges:
                if e0 < a0:
                    continue
                sub = N[a0 - 1:e0, :NNUCR[j - 1]]
                suba = AN[a0 - 1:e0, :NNUCR[j - 1]]
                hit = np.argwhere((sub == i) & (np.abs(suba) > tiny))
                for (kk, l) in hit:
                    a_list.append(suba[kk, l]); mu_list.append(a0 + kk)
                    cnt += 1
            lse[c, i - 1] = cnt
        nan[c] = cnt
        out['a%d' % (c + 1)] = np.array(a_list, np.float64)
        out['mu%d' % (c + 1)] = np.array(mu_list, np.int32)
    out['la'] = lsa; out['le'] = lse; out['nan'] = nan

    # xnet_data.F90:855-866 : FFN / NNU table pointers smuggled in rc1(1,:)
    iwk1 = out['iwk1']; rc1 = out['rc1']
    iffn = np.zeros(nreac[0], np.int32); innu = np.zeros(nreac[0], np.int32)
    for kk in range(nreac[0]):
        if iwk1[kk] in (2, 3):
            iffn[kk] = int(np.rint(rc1[kk, 0]))
        elif iwk1[kk] in (7, 8):
            innu[kk] = int(np.rint(rc1[kk, 0]))
    out['iffn'] = iffn; out['innu'] = innu
    # xnet_data.F90:901-924 : expanded reactant indices
    for c in range(4):
        mu = out['mu%d' % (c + 1)]; N = out['n%d' % (c + 1)]
        for r in range(c + 1):
            out['n%d%d' % (c + 1, r + 1)] = (N[mu - 1, r] if len(mu) else np.zeros(0, np.int32)).astype(np.int32)
    return out


def net_desc_text(nuc, net):
    """Text of net_desc as written at xnet_preprocess.F90:971-980 (without sparsity lines)."""
    L = []
    L.append('Number of Nuclear Species=%5d' % nuc['ny'])
    L.append('Reaction Count for the 11 different types')
    for j in range(NCHAP):
        L.append('%10d%10d%10d' % (j + 1, net['chap_la'][j], net['chap_le'][j]))
    L.append('Necessary dimensions')
    L.append('%17s' % 'nreac(1,2,3,4)=' + ''.join('%8d' % x for x in net['nreac']))
    L.append('%17s' % 'nan(1,2,3,4)=' + ''.join('%8d' % x for x in net['nan']))
    L.append('Reaction Count for non-REACLIB rates')
    L.append('%7s%8d%7s%8d' % ('nffn=', net['nffn'], ' nnu=', net['nnnu']))
    return L


def read_ffn(data_dir, nffn, iweak0):
    """xnet_ffn.F90:37-137. Returns tables shaped (nffn,143), grid index j = 13*(ie-1)+it."""
    ngrid = 143
    d = dict(ffnsum=np.zeros((max(nffn, 1), ngrid)), ffnenu=np.zeros((max(nffn, 1), ngrid)),
             ffn_ec=np.zeros((max(nffn, 1), ngrid)), ffn_beta=np.zeros((max(nffn, 1), ngrid)),
             ffn_qval=np.zeros(max(nffn, 1)), has_logft=np.zeros(max(nffn, 1), np.int32))
    if nffn <= 0:
        return d
    with open(os.path.join(data_dir, 'netweak')) as f:
        lines = f.read().split('\n')
    pos = 0
    if abs(iweak0) == 2:
        for i in range(nffn):
            h = _ld_tokens(lines[pos]); pos += 1
            desc = h[2][:3]
            d['ffn_qval'][i] = float(h[3])
            d['has_logft'][i] = 2 if desc == 'ft+' else (1 if desc == 'ft-' else 0)
            vals = []
            while len(vals) < 4 * ngrid:
                vals += [float(x) for x in _ld_tokens(lines[pos])]; pos += 1
            v = np.array(vals[:4 * ngrid]).reshape(ngrid, 4)
            d['ffn_beta'][i] = v[:, 0]; d['ffn_ec'][i] = v[:, 1]; d['ffnsum'][i] = v[:, 2]; d['ffnenu'][i] = v[:, 3]
        ec = d['ffn_ec']
        ec[ec < -30.0] = 99.0
    else:
        for i in range(nffn):
            pos += 1
            vals = []
            while len(vals) < 2 * ngrid:
                vals += [float(x) for x in _ld_tokens(lines[pos])]; pos += 1
            v = np.array(vals[:2 * ngrid]).reshape(ngrid, 2)
            d['ffnsum'][i] = v[:, 0]; d['ffnenu'][i] = v[:, 1]
    return d


# --------------------------------------------------------------------------------------------------
# control file
# --------------------------------------------------------------------------------------------------
def _find_block(lines, key):
    for i, ln in enumerate(lines[:1000]):
        if ln[:2] == '##' and key in ln[2:82]:
            return i + 1
    return None


def _ld_first(line):
    return _ld_tokens(line)[0]


def read_controls(text):
    """xnet_controls.F90:164-396.  `text` = contents of the `control` file."""
    lines = text.split('\n')
    c = {}
    p = _find_block(lines, 'Problem Description')
    c['descript'] = [(lines[p + i] + ' ' * 80)[:80] for i in range(3)]
    p = _find_block(lines, 'Job Controls')
    c['szone'] = int(_ld_first(lines[p])); c['nzone'] = int(_ld_first(lines[p + 1]))
    c['iweak0'] = int(_ld_first(lines[p + 2])); c['iscrn'] = int(_ld_first(lines[p + 3]))
    c['iprocess'] = int(_ld_first(lines[p + 4]))
    p = _find_block(lines, 'Zone Batching Controls')
    c['nzbatchmx'] = int(_ld_first(lines[p])) if p is not None else 1
    p = _find_block(lines, 'Integration Controls')
    c['isolv'] = int(_ld_first(lines[p])); c['kstmx'] = int(_ld_first(lines[p + 1]))
    c['kitmx'] = int(_ld_first(lines[p + 2])); c['ijac'] = int(_ld_first(lines[p + 3]))
    c['iconvc'] = int(_ld_first(lines[p + 4])); c['changemx'] = _ffloat(_ld_first(lines[p + 5]))
    c['yacc'] = _ffloat(_ld_first(lines[p + 6])); c['tolm'] = _ffloat(_ld_first(lines[p + 7]))
    c['tolc'] = _ffloat(_ld_first(lines[p + 8])); c['ymin'] = _ffloat(_ld_first(lines[p + 9]))
    c['tdel_maxmult'] = _ffloat(_ld_first(lines[p + 10]))
    p = _find_block(lines, 'Self-heating Controls')
    if p is not None:
        c['iheat'] = int(_ld_first(lines[p])); c['changemxt'] = _ffloat(_ld_first(lines[p + 1]))
        c['tolt9'] = _ffloat(_ld_first(lines[p + 2]))
    else:
        c['iheat'] = 0; c['changemxt'] = 1.0e-2; c['tolt9'] = 1.0e-4
    if c['isolv'] == 3:
        c['changemx'] = 1.0e+10; c['changemxt'] = 1.0e+10
    p = _find_block(lines, 'NSE Initial Conditions')
    c['t9nse'] = _ffloat(_ld_first(lines[p])) if p is not None else 8.0
    p = _find_block(lines, 'Neutrinos')
    c['ineutrino'] = int(_ld_first(lines[p])) if p is not None else 0
    p = _find_block(lines, 'Output Controls')
    c['idiag'] = int(_ld_first(lines[p])); c['itsout'] = int(_ld_first(lines[p + 1]))
    c['ev_file_base'] = lines[p + 3][:80].strip(); c['bin_file_base'] = lines[p + 5][:80].strip()
    nn = (lines[p + 6] + ' ' * 54)[50:54].strip()
    c['nnucout'] = int(nn) if nn else 0
    names = (lines[p + 7] + ' ' * 70)[:70]
    c['output_nuc'] = [names[5 * i:5 * i + 5] for i in range(min(c['nnucout'], 14))]
    p = _find_block(lines, 'Input Controls')
    c['data_dir'] = lines[p + 1][:80].strip()
    q = p + 3
    inab = []; th = []
    for izone in range(c['nzone']):
        if q >= len(lines) or (q == len(lines) - 1 and lines[q] == ''):
            break
        a = lines[q][:80].strip()
        if q + 1 >= len(lines) or (q + 1 == len(lines) - 1 and lines[q + 1] == ''):
            break
        t = lines[q + 1][:80].strip()
        inab.append(a); th.append(t); q += 2
    nread = len(inab)
    if nread == 1 and nread < c['nzone']:
        base_a, base_t = inab[0], th[0]
        inab = [name_ordered(base_a, i + 1, c['nzone']) for i in range(c['nzone'])]
        th = [name_ordered(base_t, i + 1, c['nzone']) for i in range(c['nzone'])]
    c['inab_file'] = inab; c['thermo_file'] = th
    return c


def name_ordered(base, n, nmax):
    """xnet_util.F90:341-370 : append n zero-padded to the digit count of nmax."""
    nd = len(str(nmax))
    return base + str(n).zfill(nd)


# --------------------------------------------------------------------------------------------------
# per-zone inputs
# --------------------------------------------------------------------------------------------------
def read_thermo_file(path, nhmx=50000):
    """model_input_ascii.F90:5-110."""
    with open(path) as f:
        lines = f.read().split('\n')
    desc = (lines[0] + ' ' * 80)[:80]
    tstart = _ffloat(_ld_first(lines[1])); tstop = _ffloat(_ld_first(lines[2]))
    tdelstart = _ffloat(_ld_first(lines[3]))
    th = []; t9h = []; rhoh = []; yeh = []
    body = lines[4:]
    if body and body[-1] == '':
        body = body[:-1]
    for ln in body[:nhmx]:
        tok = ln.replace('\t', ' ').split()
        if len(tok) < 3:
            raise ValueError('Not enough columns in thermo file: ' + path)
        th.append(_ffloat(tok[0])); t9h.append(_ffloat(tok[1])); rhoh.append(_ffloat(tok[2]))
        yeh.append(_ffloat(tok[3]) if len(tok) > 3 else 0.0)
    tdelstart = min(0.0, tdelstart)
    return dict(desc=desc, tstart=tstart, tstop=tstop, tdelstart=tdelstart, nh=len(th),
                th=np.array(th), t9h=np.array(t9h), rhoh=np.array(rhoh), yeh=np.array(yeh))


def y_moment(y, nuc, xext, aext, zext):
    """xnet_abundances.F90:93-185 (scalar form); sequential sums."""
    aa = nuc['aa']; zz = nuc['zz']
    zz2 = zz * zz
    zzi = zz ** THBIM1
    yext = xext / aext
    ntot = atot = ztot = z2tot = zitot = 0.0
    for k in range(len(y)):
        ntot += y[k]; atot += y[k] * aa[k]; ztot += y[k] * zz[k]
        z2tot += y[k] * zz2[k]; zitot += y[k] * zzi[k]
    ntot += yext; atot += xext; ztot += yext * zext
    z2tot += yext * zext * zext; zitot += yext * zext ** THBIM1
    return dict(ye=ztot / atot, ytot=ntot, abar=atot / ntot, zbar=ztot / ntot,
                z2bar=z2tot / ntot, zibar=zitot / ntot)


def read_inab_file(path, nuc):
    """model_input_ascii.F90:223-350."""
    ny = nuc['ny']
    names_stripped = [n.strip() for n in nuc['names']]
    aa = nuc['aa']; zz = nuc['zz']
    with open(path) as f:
        lines = f.read().split('\n')
    desc = (lines[0] + ' ' * 80)[:80]
    toks = _ld_tokens('\n'.join(lines[1:]))
    yein = 0.0; yin = np.zeros(ny); yext = 0.0
    for i in range(0, len(toks) - 1, 2):
        namein = toks[i].strip("'\"")[:5].strip()
        val = _ffloat(toks[i + 1])
        if namein and val > 0.0:
            namein = namein.lower()
            if namein == 'ye':
                yein = val
            else:
                inuc = index_from_name(namein, names_stripped, ny)
                if inuc < 1 or inuc > ny:
                    yext += val
                else:
                    yin[inuc - 1] = val
    xnet = 0.0; znet = 0.0
    for i in range(ny):          # sum(yin*aa), sum(yin*zz) evaluated left to right
        xnet += yin[i] * aa[i]; znet += yin[i] * zz[i]
    if yext > 0.0 and xnet < 1.0 and znet < yein:
        xext = 1.0 - xnet
        aext = xext / yext
        if yein <= 0.0 or yein >= 1.0:
            zext = 0.0
        else:
            zext = (yein - znet) * aext / xext
    else:
        xext = 0.0; aext = 1.0; zext = 0.0
        yin = yin / xnet
    return dict(desc=desc, yein=yein, yin=yin, xext=xext, aext=aext, zext=zext, xnet=xnet)


def t9rhofind(tf, th, t9h, rhoh):
    """xnet_conditions.F90:50-102 ; returns (nf, t9f, rhof), nf 1-based."""
    ns = len(th)
    if ns == 1:
        return 1, t9h[0], rhoh[0]
    n = ns + 1
    for i in range(1, ns + 1):
        if tf <= th[i - 1]:
            n = i
            break
    if 1 < n <= ns:
        rdt = 1.0 / (th[n - 1] - th[n - 2])
        dt = tf - th[n - 2]
        dt9 = t9h[n - 1] - t9h[n - 2]
        drho = rhoh[n - 1] - rhoh[n - 2]
        return n, dt * rdt * dt9 + t9h[n - 2], dt * rdt * drho + rhoh[n - 2]
    if n == 1:
        return n, t9h[0], rhoh[0]
    return n, t9h[ns - 1], rhoh[ns - 1]


def load_zone(inab_path, th_path, nuc, t9nse=8.0):
    """net.F90:246-272 + load_initial_abundances (model_input_ascii.F90:112-221), non-NSE branch."""
    th = read_thermo_file(th_path)
    nstart, t9start, rhostart = t9rhofind(th['tstart'], th['th'], th['t9h'], th['rhoh'])
    nh = th['nh']
    if 1 < nstart <= nh:
        rdt = 1.0 / (th['th'][nstart - 1] - th['th'][nstart - 2])
        dt = th['tstart'] - th['th'][nstart - 2]
        dye = th['yeh'][nstart - 1] - th['yeh'][nstart - 2]
        yestart = dt * rdt * dye + th['yeh'][nstart - 2]
    elif nstart == 1:
        yestart = th['yeh'][0]
    else:
        yestart = th['yeh'][nh - 1]
    if t9start > t9nse:
        raise NotImplementedError('NSE initial abundances are out of scope (SURVEY 2a)')
    ab = read_inab_file(inab_path, nuc)
    xext, aext, zext = ab['xext'], ab['aext'], ab['zext']
    yein = ab['yein']
    if yein <= 0.0 or yein >= 1.0:
        yein = y_moment(ab['yin'], nuc, xext, aext, zext)['ye']
    yestart = yein
    z = dict(th)
    z.update(nstart=nstart, t9start=t9start, rhostart=rhostart, yestart=yestart, ystart=ab['yin'],
             xext=xext, aext=aext, zext=zext, abund_desc=ab['desc'], thermo_desc=th['desc'])
    return z


def match_react(net):
    """Reaction-pair matching: match_react + flux_search (xnet_preprocess.F90:55-289) and flux_init
    (xnet_flux.F90:20-65).  Adds mflx, ifl1..ifl4 (signed, 1-based), nflx [mflx,8], iwflx, qflx, dcflx [mflx,2],
    ifl_orig, ifl_term to `net`.  Independent of the product's C++ version: sets are looked up in a dict keyed by
    the 8-tuple (first occurrence wins, which is what the reference's linear search returns)."""
    sets = []; index = {}; iw_of = []; q_of = []
    ifl = [np.zeros(net['nreac'][c], np.int32) for c in range(4)]
    for j in range(1, NCHAP + 1):
        nreact = NREACTANT[j - 1]; nnuc = NNUCR[j - 1]; c = nreact - 1
        N = net['n%d' % nreact]; Q = net['q%d' % nreact]; IWK = net['iwk%d' % nreact]
        for k in range(net['chap_la'][j - 1], net['chap_le'][j - 1] + 1):
            if k < 1 or k > net['nreac'][c]:
                continue
            reac = [int(v) for v in N[k - 1, :nreact]]; prod = [int(v) for v in N[k - 1, nreact:nnuc]]
            fwd = Q[k - 1] > 0.0
            lhs, rhs = (reac, prod) if fwd else (prod, reac)
            key = tuple(lhs + [0] * (4 - len(lhs)) + rhs + [0] * (4 - len(rhs)))
            iw = int(IWK[k - 1])
            m = index.get(key) if iw == 0 else None
            if m is None:
                sets.append(key); iw_of.append(iw); q_of.append(abs(float(Q[k - 1])))
                m = len(sets)
                # only non-weak lookups may hit it later, but a weak set can be found by a later non-weak reaction:
                # the reference searches ALL existing sets, whatever their flag
                index.setdefault(key, m)
            ifl[c][k - 1] = m if fwd else -m
    mflx = len(sets)
    nflx = np.array(sets, np.int32).reshape(mflx, 8)
    fact = {1: 1.0, 2: 0.5, 3: 1.0 / 6.0, 4: 1.0 / 24.0, 5: 0.25}
    dc = np.ones((mflx, 2)); orig = np.zeros(mflx, np.int32); term = np.zeros(mflx, np.int32)
    for m in range(mflx):
        for side in (0, 1):
            f = nflx[m, 4 * side:4 * side + 4]
            cnt = 1 + sum(1 for s in (1, 2, 3) if f[s] != 0 and f[s] == f[s - 1])
            if cnt == 3 and f[1] != f[2]:
                cnt = 5
            dc[m, side] = fact[cnt]
            nz = int(np.count_nonzero(f))
            (orig if side == 0 else term)[m] = f[nz - 1] if nz else 0
    net.update(mflx=mflx, nflx=nflx, iwflx=np.array(iw_of, np.int32), qflx=np.array(q_of), dcflx=dc,
               ifl_orig=orig, ifl_term=term)
    for c in range(4):
        net['ifl%d' % (c + 1)] = ifl[c]
    return net


def load_network(data_dir, iweak0=1):
    nuc = read_nuclear_data(data_dir)
    net = net_preprocess(data_dir, nuc)
    match_react(net)
    ffn = read_ffn(data_dir, net['nffn'], iweak0)
    return nuc, net, ffn


def load_problem(control_text, root):
    """Convenience: parse control, load network and all zones (paths relative to `root`)."""
    ctl = read_controls(control_text)
    ddir = os.path.join(root, ctl['data_dir'])
    nuc, net, ffn = load_network(ddir, ctl['iweak0'])
    zones = [load_zone(os.path.join(root, a), os.path.join(root, t), nuc, ctl['t9nse'])
             for a, t in zip(ctl['inab_file'], ctl['thermo_file'])]
    return ctl, nuc, net, ffn, zones
