#!/bin/bash
# r01f pass: new parity tests (step_be, self-heating module path, flux/ts) + A/B bench of kernel variants.
# usage: gpu_r01f.sh "<lib1> <lib2> ..."   (names relative to xnet_b200/, "default" = libxnet_b200.so)
libs=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flux_ts.py tests/test_gpu_stepbe.py -m gpu -q > gpurun_out/pytest_r01f.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r01f.log
grep -n "^E  \|^FAILED\|passed\|failed\|pytest exit" gpurun_out/pytest_r01f.log | cut -c1-260 | head -40
for l in $libs; do
  if [ "$l" = default ]; then unset XNET_B200_LIB; else export XNET_B200_LIB=$PWD/xnet_b200/$l; fi
  timeout 600 python bench.py --zones 2368 --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/bench_$l.json 2> gpurun_out/bench_$l.err
  echo "== $l exit $?"; python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$l.json'))
    r = d['roofline']
    print('value %.4g  cycles/iter %.0f  frac %.4f  launch %s' % (d['value'], r['cycles_per_newton_iter'], r['frac'], d['config']['launch']))
    print(' shares', r['phase_share'])
    print(' dbg %s' % (r['dbg_cycles_per_newton_iter'],))
except Exception as e:
    print('no result', e)
PY
  tail -2 gpurun_out/bench_$l.err
done
