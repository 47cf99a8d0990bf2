#!/bin/bash
# one full ncu capture (with source) of a kernel:  bash scratch/ncu_kernel.sh TAG KERNEL_REGEX SKIP [bench args]
set -u
TAG=$1; RE=$2; SKIP=$3; shift 3
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c 1 -o gpurun_out/${TAG} -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
