#!/bin/bash
# A/B bench of library variants on one workload, no tests. usage: gpu_ab2.sh "<lib> ..." workload zones
libs=$1; w=$2; nz=$3
mkdir -p gpurun_out
for l in $libs; do
  if [ "$l" = default ]; then unset XNET_B200_LIB; else export XNET_B200_LIB=$PWD/xnet_b200/$l; fi
  timeout 900 python bench.py --workload $w --zones $nz --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_$l.json 2> gpurun_out/bench_$l.err
  echo "== $l $w exit $?"; python - <<PY
import json
d = json.load(open('gpurun_out/bench_$l.json'))
r = d['roofline']
print('value %.4g  cycles/iter %.0f  frac %.4f' % (d['value'], r['cycles_per_newton_iter'], r['frac']))
print(' shares', r['phase_share'])
print(' pivots/lu %.2f partial/lu %.2f dbg %s' % (r['pivot_steps_per_lu'], r['partially_accepted_panels_per_lu'], r['dbg_cycles_per_newton_iter']))
PY
done
