#!/bin/bash
# Iteration pass: parity tests + short bench.  usage: gpu_iter.sh [bench args...]
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --zones 2368 --steps 1 --warmup 1 --cpu-zones 8 "$@" > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench exit $?"
cat gpurun_out/bench_small.json; tail -3 gpurun_out/bench_small.err
