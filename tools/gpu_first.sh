#!/bin/bash
# First GPU pass of a session: parity tests, a short bench for timing estimate, ncu launch list + full capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --zones 2368 --steps 1 --warmup 1 --cpu-zones 8 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench exit $?"
cat gpurun_out/bench_small.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
  python bench.py --zones 592 --steps 1 --warmup 1 --kstmx 200 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:evolve_kernel -s 1 -c 1 -o gpurun_out/prof_evolve \
  python bench.py --zones 592 --steps 1 --warmup 1 --kstmx 100 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
