#!/bin/bash
# parity tests + short benches of several workloads.  usage: gpu_iter2.sh "<workload:zones> ..." [bench args]
wl=$1; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for w in $wl; do
  name=${w%%:*}; nz=${w##*:}
  timeout 900 python bench.py --workload $name --zones $nz --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "== $name x $nz exit $?"; python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$name.json'))
    r = d['roofline']
    print('value %.4g e2e %.4g ms/step %.1f cycles/iter %.0f  frac %.4f finished %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['cycles_per_newton_iter'], r['frac'], d['config']['all_zones_finished']))
    print(' shares', r['phase_share']); print(' launch', d['config']['launch'], 'pivots/lu %.2f' % r['pivot_steps_per_lu'])
except Exception as e:
    print('no result', e)
PY
  tail -2 gpurun_out/bench_$name.err
done
