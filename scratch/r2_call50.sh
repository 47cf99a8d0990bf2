#!/bin/bash
# 1-D histories under the refill schedule: one segment per trip (u1) against two (u2); static schedule for reference
set -u
mkdir -p gpurun_out
{
AB_TAG=c50_su32 BENCH_ARGS="--workload suolson_f32 --steps 10" bash scratch/ab.sh u2 u1
AB_TAG=c50_su32s BENCH_ARGS="--workload suolson_f32 --steps 10 --track history" bash scratch/ab.sh u2
AB_TAG=c50_su16 BENCH_ARGS="--workload suolson_f16 --steps 5" bash scratch/ab.sh u2 u1
} 2>&1 | tee gpurun_out/r2_call50.log
