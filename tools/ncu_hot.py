#!/usr/bin/env python3
"""Per-function and per-instruction warp-stall samples from `ncu --page source --csv` output.
usage: tools/ncu_hot.py src.csv [topN]   (function boundaries = CALL.REL targets inside the kernel)"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][ix['Address']], 16)
reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
ins = []
for r in data:
    if len(r) < len(hdr): continue
    a = int(r[ix['Address']], 16) - base
    st = {h: int(r[ix[h]] or 0) for h in reasons}
    ins.append((a, r[ix['Source']].strip(), int(r[ix['# Samples']] or 0), st, int(r[ix['Instructions Executed']] or 0)))
targets = sorted({int(m.group(1), 16) - (base if int(m.group(1), 16) >= base else 0) for a, s, n, st, ex in ins for m in [re.search(r'CALL\.REL\.NOINC (0x[0-9a-f]+)', s)] if m})
bounds = [0] + targets + [ins[-1][0] + 16]
tot = sum(n for a, s, n, st, ex in ins)
print('total samples', tot)
print('%-10s %7s %8s %6s %10s  %s' % ('func@', 'instrs', 'samples', 'pct', 'executed', 'top stall reasons (non-barrier)'))
for k in range(len(bounds) - 1):
    lo, hi = bounds[k], bounds[k + 1]
    sel = [x for x in ins if lo <= x[0] < hi]
    n = sum(x[2] for x in sel)
    if n == 0: continue
    agg = collections.Counter()
    for x in sel:
        for h, v in x[3].items(): agg[h] += v
    bar = agg.pop('stall_barrier', 0)
    ex = sum(x[4] for x in sel)
    print('%-10s %7d %8d %5.1f%% %10d  barrier:%d %s' % (hex(lo), len(sel), n, 100.0 * n / tot, ex, bar, ' '.join('%s:%d' % (h[6:], v) for h, v in agg.most_common(5))))
print('\ntop instructions by non-barrier samples')
key = lambda x: x[2] - x[3].get('stall_barrier', 0)
for x in sorted(ins, key=key, reverse=True)[:topn]:
    st = {h[6:]: v for h, v in x[3].items() if v and h != 'stall_barrier'}
    print('%8s %6d  %-70s %s' % (hex(x[0]), key(x), x[1][:70], ' '.join('%s:%d' % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])))
