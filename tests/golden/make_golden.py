"""Generates the committed golden vectors from the CPU oracle (run from the repo root):
    python tests/golden/make_golden.py
* net_desc_<net>.txt : copies of the reference's shipped net_desc files (the reference's own
  known-answer files for the reaction-map / indexing work).
* eos_points.npz : Helmholtz/Fermi-Dirac/screening scalars of the oracle at a grid of (T9, rho) points.
* ts_<problem>_<math>.npz : accepted-step records (kstep, t, tdel, t9, rho, y, kmon) of the oracle
  for `libm` (reference-faithful elementary functions) and `pm` (portable-math) builds.
"""
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import xnet_oracle as O  # noqa: E402

DATA = os.path.join(ROOT, 'data')
OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {'alpha': ('test_settings_small', 'setup_tnsn_alpha', 3000, None),
         'torch47': ('test_settings', 'setup_tnsn_torch47', 400, 399),
         'sn160': ('test_settings', 'setup_tnsn_sn160', 120, 119),
         'xrb304': ('test_settings', 'setup_xrb_fisker', 40, 39),
         # screening switched on (Helmholtz eta, Fermi-Dirac integrals, Graboske/DeWitt blend)
         'torch47_scrn': ('test_settings', 'setup_tnsn_torch47', 400, 399),
         'sn160_scrn': ('test_settings', 'setup_tnsn_sn160', 120, 119),
         # self-heating + screening (test_settings_heat)
         'torch47_heat': ('test_settings_heat', 'setup_heat_torch47', 300, 299),
         'sn160_heat': ('test_settings_heat', 'setup_heat_sn160', 120, 119),
         # logft weak rates + self-heating + screening, SN231 (N = 232: global-memory Jacobian)
         'sn231_logft': ('test_settings_logft', 'setup_logft_sn231', 60, 59)}

for net in ('alpha', 'torch47', 'SN160', 'SN231', 'XRB304'):
    shutil.copy(os.path.join(DATA, 'Data_' + net, 'net_desc'), os.path.join(OUT, 'net_desc_%s.txt' % net))

for name, (st, su, cap, kstmx) in CASES.items():
    txt = open(os.path.join(DATA, st)).read() + open(os.path.join(DATA, 'Test_Problems', su)).read()
    if name.endswith('_scrn'):
        txt = txt.replace('0         Include Screening (yes=1)', '1         Include Screening (yes=1)')
    p = O.Problem.from_control(txt, DATA)
    for math, fast in (('libm', False), ('pm', 'pm')):
        r = p.run(nzone=1, hist_cap=cap, fast=fast, ctl_override=dict(kstmx=kstmx) if kstmx else None)
        n = min(int(r['hist_n'][0]), cap)
        np.savez_compressed(os.path.join(OUT, 'ts_%s_%s.npz' % (name, math)),
                            kstep=r['hist_kstep'][0, :n], t=r['hist_t'][0, :n], tdel=r['hist_tdel'][0, :n],
                            t9=r['hist_t9'][0, :n], rho=r['hist_rho'][0, :n], y=r['hist_y'][0, :n],
                            kmon=r['hist_kmon'][0, :n], edot=r['hist_edot'][0, :n], ktot=r['ktot'][0],
                            kstmx=np.int32(kstmx or p.ctl['kstmx']))
        print(name, math, n, 'records')
