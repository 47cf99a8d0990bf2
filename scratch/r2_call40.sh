#!/bin/bash
# A/B: double-precision polynomial coefficients from __constant__ tables (dtab2) against literals (base)
set -u
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
AB_TAG=c40_f64 BENCH_ARGS="--workload crookedpipe_f64" bash scratch/ab.sh base dtab2
AB_TAG=c40_f32 bash scratch/ab.sh base dtab2
AB_TAG=c40_rw BENCH_ARGS="--workload marshak_f32_rw" bash scratch/ab.sh base dtab2
} 2>&1 | tee gpurun_out/r2_call40.log
