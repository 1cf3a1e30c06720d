#!/bin/bash
# A/B of launch variants of the structured kernel through its environment knobs (no rebuild), short bench each.
# usage: bash tools/gpu_variants.sh "VAR=val ..." "VAR=val ..." ...     ("" = defaults)
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-extras ${BENCH_ARGS}"
for v in "$@"; do
  echo "== variant [$v]"
  env $v $B > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/v.json')); r=d['roofline']
    print('value %.4g cycles/iter %.0f launch %s' % (d['value'], r['cycles_per_newton_iter'], d['config']['launch']))
    cyc = r['cycles_per_newton_iter']
    print(' cycles', {k: int(v * cyc) for k, v in r['phase_share'].items() if v}, 'fb', r.get('dense_fallbacks_per_lu'), r.get('structured'))
except Exception as e: print('no result', e)
PY
done
