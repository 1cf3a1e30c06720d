#!/bin/bash
# One script for every builder-side GPU pass (run under gpurun from the repo root).  Stages are arguments, run in order:
#   tests[:<pytest args>]   pytest -m gpu (default: the whole suite)
#   bench[:<bench args>]    python bench.py <args>  -> gpurun_out/bench_<n>.json (+ .err with wall time)
#   ref[:<bench args>]      python bench.py --impl reference <args>
#   probes                  tools/fp64_probe.bin, tools/dmma_probe.bin -> gpurun_out/*.txt
#   launches[:<bench args>] ncu launch list of a short bench run
#   ncu[:<bench args>]      ncu --set full capture of the evolve kernel of a short bench run
# e.g.  gpurun --timeout 1500 -- 'bash tools/gpu_run.sh tests "bench:--steps 20 --warmup 5" probes'
mkdir -p gpurun_out
n=0
for st in "$@"; do
  kind=${st%%:*}; arg=""; [ "$st" != "$kind" ] && arg=${st#*:}
  n=$((n+1))
  t0=$(date +%s)
  case $kind in
    tests)
      timeout ${TESTS_TIMEOUT:-1500} python -m pytest ${arg:-tests} -m gpu ${PYTEST_ARGS:--x} -q > gpurun_out/pytest_$n.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$n.log
      grep -n "^E  \|^FAILED\|^ERROR\|passed\|failed\|pytest exit\|Timeout" gpurun_out/pytest_$n.log | cut -c1-300 | head -${TESTS_LINES:-60} ;;
    bench)
      timeout 1200 python bench.py $arg > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err; echo "bench exit $? ($arg)"
      tail -3 gpurun_out/bench_$n.err | cut -c1-400
      python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_$n.json'))
    r = d['roofline']
    print('value %.4g  e2e %.4g  ms/step %.1f  cycles/iter %.0f  frac %.4f  launch %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['cycles_per_newton_iter'], r['frac'], d['config']['launch']))
    print(' shares', r['phase_share'])
    print(' also', {k: (v if not isinstance(v, dict) else {a: b for a, b in v.items() if a in ('value', 'seconds', 'ms_per_step')}) for k, v in d['also'].items()})
    for k in ('lu_isolated', 'cublas_batched', 'xb_over_cublas_speedup', 'structured', 'dense_fallbacks_per_lu', 'fallback_stats'):
        if k in r: print(' ', k, r[k])
    print(' cpu', d.get('cpu_baseline', {}).get('value'), 'clocks', d['clocks'])
except Exception as e:
    print('no result', e)
PY
      ;;
    ref)
      timeout 900 python bench.py --impl reference $arg > gpurun_out/ref_$n.json 2> gpurun_out/ref_$n.err; echo "ref exit $?"; cut -c1-300 gpurun_out/ref_$n.json ;;
    probes)
      make -C tools -s > /dev/null 2>&1
      ./tools/fp64_probe.bin > gpurun_out/fp64_probe.txt 2>&1; ./tools/dmma_probe.bin > gpurun_out/dmma_probe.txt 2>&1
      tail -5 gpurun_out/fp64_probe.txt gpurun_out/dmma_probe.txt ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$n.csv \
        python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline $arg > gpurun_out/launches_$n.log 2>&1
      tail -2 gpurun_out/launches_$n.log | cut -c1-200; wc -l gpurun_out/launches_$n.csv ;;
    ncu)
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:evolve -s 2 -c 1 -o gpurun_out/prof_$n -f \
        python bench.py --steps 1 --warmup 1 --zones 1184 --no-extras --no-cpu-baseline $arg > gpurun_out/prof_$n.log 2>&1
      tail -2 gpurun_out/prof_$n.log | cut -c1-200; ls -la gpurun_out/prof_$n.ncu-rep ;;
    *) echo "unknown stage $kind" ;;
  esac
  echo "== stage $n ($kind) took $(( $(date +%s) - t0 )) s"
done
