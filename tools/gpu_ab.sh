#!/bin/bash
# A/B pass: parity tests on the default library, then a short bench per library variant.
# usage: gpu_ab.sh "<lib1> <lib2> ..." [bench args...]   (library names relative to xnet_b200/, "default" = libxnet_b200.so)
libs=$1; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for l in $libs; do
  if [ "$l" = default ]; then unset XNET_B200_LIB; else export XNET_B200_LIB=$PWD/xnet_b200/$l; fi
  timeout 900 python bench.py --zones 2368 --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/bench_$l.json 2> gpurun_out/bench_$l.err
  echo "== $l exit $?"; python - <<PY
import json
d = json.load(open('gpurun_out/bench_$l.json'))
r = d['roofline']
print('value %.4g  cycles/iter %.0f  frac %.4f' % (d['value'], r['cycles_per_newton_iter'], r['frac']))
print(' shares', r['phase_share'])
print(' pivots/lu %.2f partial/lu %.2f dbg %s' % (r['pivot_steps_per_lu'], r['partially_accepted_panels_per_lu'], r['dbg_cycles_per_newton_iter']))
PY
  tail -2 gpurun_out/bench_$l.err
done
