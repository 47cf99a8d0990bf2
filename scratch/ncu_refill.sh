#!/bin/bash
# one full ncu capture (with source) of the refill kernel at a steady-state step:  bash scratch/ncu_refill.sh TAG [bench args]
set -u
TAG=$1; shift
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_track_refill -s 4 -c 1 -o gpurun_out/${TAG}_refill -f python bench.py --track refill --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-200
