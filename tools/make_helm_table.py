#!/usr/bin/env python
"""Packs the 13 Helmholtz-table arrays XNet's pruned EOS reads (f, fd, ft, fdd, ftt, fdt, fddt, fdtt, fddtt,
ef, efd, eft, efdt) out of the ASCII helm_table.dat (tools/starkiller-helmholtz/helm_table.dat in the XNet
tree, 60 MB) into data/helm_table.bin: little-endian float64 [13][201][541] (density index fastest),
11.3 MB.  The ASCII file is an input-data fixture, not source; the product loader (xb_loader.cpp) reads
either form.  usage: tools/make_helm_table.py [helm_table.dat] [out.bin]"""
import os
import sys
import numpy as np

IMAX, JMAX = 541, 201
src = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/tools/starkiller-helmholtz/helm_table.dat'
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'data', 'helm_table.bin')
n = IMAX * JMAX
with open(src) as fh:
    lines = fh.readlines()
assert len(lines) == 4 * n, len(lines)
free = np.array([ln.split() for ln in lines[:n]], dtype=np.float64)            # [j*IMAX+i][9]  (actual_eos.F90:1620-1625)
ef = np.array([ln.split() for ln in lines[2 * n:3 * n]], dtype=np.float64)     # [..][4]  ef, efd, eft, efdt (:1635-1639)
assert free.shape == (n, 9) and ef.shape == (n, 4)
planes = np.concatenate([free.T, ef.T], axis=0).reshape(13, JMAX, IMAX)
planes.astype('<f8').tofile(dst)
print('wrote', dst, planes.nbytes, 'bytes')
