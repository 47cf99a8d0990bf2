#!/bin/bash
# census tally: run ends handed between lanes, warp-private adds (wp2) against HEAD (cA); full suite on the tree's library
set -u
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
AB_TAG=c51_su32 BENCH_ARGS="--workload suolson_f32 --track auto --steps 10" bash scratch/ab.sh cA wp2
AB_TAG=c51_su64 BENCH_ARGS="--workload suolson_f64 --track auto --steps 10" bash scratch/ab.sh cA wp2
IMC_LIB=$PWD/variants/libimc_wp2.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c51_su32_launches.csv python bench.py --workload suolson_f32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c51_su32_launches.log 2>&1
} 2>&1 | tee gpurun_out/r2_call51.log
