#!/bin/bash
# Stage A: full GPU parity suite on the default library, then short A/B benches against the previous build.
# usage: gpu_stage_a.sh ["<workload:zones> ..."]
wl=${1:-"sn160_scrn:2368"}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
grep -n "^E  \|^FAILED\|passed\|failed" gpurun_out/pytest_gpu.log | cut -c1-240 | head -30
bash tools/gpu_ab3.sh "libxnet_b200_prev.so default" "$wl"
