#!/bin/bash
# full ncu capture of the mode-sum kernel (bench workload, B = 16); usage: gpu_ncu.sh TAG [extra bench args]
TAG=$1; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:mode_sum_kernel -s 4 -c 1 -o gpurun_out/${TAG}_modesum -f \
  python bench.py --batch 16 --steps 2 --warmup 3 --no-cpu-baseline --no-extras "$@" > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
