"""Full-length strict goldens (run from the repo root):  python tests/golden/make_golden_full.py
Complete trajectories of the portable-math oracle, stored as every STRIDE-th accepted-step record plus SHA-256 digests of
the complete per-step arrays (strict arithmetic is bit-exact, so a digest pins every step without a multi-megabyte
fixture).  ts_<problem>_full_pm.npz: kstep, t, tdel, t9, y of the stored records; n (records), idx (their indices);
sha_tdel, sha_t9, sha_y, sha_kmon (hex digests of the little-endian bytes of the complete arrays); ktot; final y."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import xnet_oracle as O  # noqa: E402

DATA = os.path.join(ROOT, 'data')
OUT = os.path.dirname(os.path.abspath(__file__))
STRIDE = 50
CASES = {'xrb304': ('test_settings', 'setup_xrb_fisker', 4000)}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).astype(a.dtype.newbyteorder('<')).tobytes()).hexdigest()


if __name__ == '__main__':
    for name, (st, su, cap) in CASES.items():
        txt = open(os.path.join(DATA, st)).read() + open(os.path.join(DATA, 'Test_Problems', su)).read()
        p = O.Problem.from_control(txt, DATA)
        r = p.run(nzone=1, hist_cap=cap, fast='pm')
        n = int(r['hist_n'][0])
        assert n <= cap and r['status'][0] == -1, (n, r['status'])
        idx = np.unique(np.concatenate([np.arange(0, n, STRIDE), [n - 1]]))
        np.savez_compressed(os.path.join(OUT, 'ts_%s_full_pm.npz' % name), n=np.int32(n), idx=idx.astype(np.int32),
                            kstep=r['hist_kstep'][0, idx], t=r['hist_t'][0, idx], tdel=r['hist_tdel'][0, idx],
                            t9=r['hist_t9'][0, idx], y=r['hist_y'][0, idx], ktot=r['ktot'][0], y_final=r['y'][0],
                            sha_tdel=sha(r['hist_tdel'][0, :n]), sha_t9=sha(r['hist_t9'][0, :n]),
                            sha_y=sha(r['hist_y'][0, :n]), sha_kmon=sha(r['hist_kmon'][0, :n]))
        print(name, n, 'records,', len(idx), 'stored')
