#!/bin/bash
# short benches of several workloads, optionally with a library variant. usage: gpu_wl.sh <lib|default> "<workload:zones> ..." [bench args]
lib=$1; wl=$2; shift; shift
mkdir -p gpurun_out
if [ "$lib" != default ]; then export XNET_B200_LIB=$PWD/xnet_b200/$lib; fi
for w in $wl; do
  name=${w%%:*}; nz=${w##*:}
  timeout 900 python bench.py --workload $name --zones $nz --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "== $name x $nz ($lib) exit $?"; python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$name.json'))
    r = d['roofline']
    print('value %.4g e2e %.4g ms/step %.1f cycles/iter %.0f  frac %.4f finished %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['cycles_per_newton_iter'], r['frac'], d['config']['all_zones_finished']))
    print(' shares', r['phase_share']); print(' pivots/lu %.2f partial %.2f dbg %s' % (r['pivot_steps_per_lu'], r['partially_accepted_panels_per_lu'], r['dbg_cycles_per_newton_iter']))
except Exception as e:
    print('no result', e)
PY
  tail -2 gpurun_out/bench_$name.err
done
