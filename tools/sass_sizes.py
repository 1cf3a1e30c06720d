#!/usr/bin/env python3
"""Sizes (instructions) of the noinline device functions inside a kernel + opcode histogram of the largest.
usage: tools/sass_sizes.py <kernel-name-substring>"""
import re, sys, subprocess, collections
out = subprocess.run(['cuobjdump', '-sass', 'xnet_b200/libxnet_b200.so'], capture_output=True, text=True).stdout
key = sys.argv[1] if len(sys.argv) > 1 else 'evolve_kernelILb0'
on = False; ins = []
for l in out.split('\n'):
    if 'Function :' in l:
        on = key in l
        continue
    if on:
        m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
        if m: ins.append((int(m.group(1), 16), m.group(2)))
calls = collections.Counter()
for a, i in ins:
    m = re.search(r'CALL\.REL\.NOINC (0x[0-9a-f]+)', i)
    if m: calls[int(m.group(1), 16)] += 1
t = sorted(calls); end = ins[-1][0] + 16
print('main body', t[0] // 16, 'total', end // 16)
for k, x in enumerate(t):
    nxt = t[k + 1] if k + 1 < len(t) else end
    h = collections.Counter()
    for a, i in ins:
        if x <= a < nxt:
            w = i.split(); op = w[1] if w[0].startswith('@') else w[0]
            h[op.split('.')[0]] += 1
    print(hex(x), (nxt - x) // 16, 'calls', calls[x], ' '.join('%s:%d' % kv for kv in h.most_common(8)))
