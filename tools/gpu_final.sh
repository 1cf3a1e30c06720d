#!/bin/bash
# End-of-round measurement pass: parity tests, the default bench (both arms), ncu launch list and full capture,
# short benches of the other workloads.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
if [ -n "$SKIP_TESTS" ]; then echo "parity tests skipped (already run on this binary)"; else
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
grep -n "^E  \|^FAILED\|passed\|failed" gpurun_out/pytest_gpu.log | cut -c1-240 | head -30
if [ $rc -ne 0 ]; then echo "parity tests failed: skipping the measurements"; exit 1; fi
fi
timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-400 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
  python bench.py --zones 592 --steps 2 --warmup 1 --kstmx 200 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:evolve_kernel -s 1 -c 1 -f -o gpurun_out/prof_final \
  python bench.py --zones 592 --steps 1 --warmup 1 --kstmx 100 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
for w in sn160:2368 sn160_heat:2368 torch47:4096 alpha:8192 xrb304:296 sn231_logft:296; do
  name=${w%%:*}; nz=${w##*:}
  timeout 600 python bench.py --workload $name --zones $nz --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "== $name exit $?"; cut -c1-260 gpurun_out/bench_$name.json
done
ls -la gpurun_out
