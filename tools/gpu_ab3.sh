#!/bin/bash
# A/B of library variants on short benches of two workloads. usage: gpu_ab3.sh "<lib> ..." "<workload:zones> ..."
libs=$1; wls=$2
mkdir -p gpurun_out
for l in $libs; do
  if [ "$l" = default ]; then unset XNET_B200_LIB; else export XNET_B200_LIB=$PWD/xnet_b200/$l; fi
  for w in $wls; do
    name=${w%%:*}; nz=${w##*:}
    timeout 600 python bench.py --workload $name --zones $nz --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ab_${l}_$name.json 2> gpurun_out/ab_${l}_$name.err
    echo "== $l $name exit $?"; python - <<PY
import json
try:
    d = json.load(open('gpurun_out/ab_${l}_$name.json'))
    r = d['roofline']
    print('value %.4g  cycles/iter %.0f  frac %.4f finished %s' % (d['value'], r['cycles_per_newton_iter'], r['frac'], d['config']['all_zones_finished']))
    print(' shares', r['phase_share'])
    print(' pivots/lu %.2f partial %.2f' % (r['pivot_steps_per_lu'], r['partially_accepted_panels_per_lu']))
except Exception as e:
    print('no result', e)
PY
    tail -2 gpurun_out/ab_${l}_$name.err
  done
done
