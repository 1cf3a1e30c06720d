#!/bin/bash
# ncu full capture of the evolve kernel on a short run.  usage: gpu_prof.sh <outname> [bench args]
name=$1; shift
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:evolve_kernel -s 1 -c 1 -o gpurun_out/$name -f \
  python bench.py --zones 296 --steps 1 --warmup 1 --kstmx 150 --no-cpu-baseline "$@" > gpurun_out/$name.log 2>&1
tail -3 gpurun_out/$name.log
ls -la gpurun_out/$name.ncu-rep
