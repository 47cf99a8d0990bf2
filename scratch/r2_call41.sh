#!/bin/bash
# A/B: Estrin evaluation of the per-segment double-precision polynomials (estrin) against Horner (dtab2)
set -u
mkdir -p gpurun_out
{
AB_TAG=c41_f64 BENCH_ARGS="--workload crookedpipe_f64" bash scratch/ab.sh dtab2 estrin
AB_TAG=c41_rw BENCH_ARGS="--workload marshak_f32_rw" bash scratch/ab.sh dtab2 estrin
AB_TAG=c41_su64 BENCH_ARGS="--workload suolson_f64 --track auto" bash scratch/ab.sh dtab2 estrin
} 2>&1 | tee gpurun_out/r2_call41.log
