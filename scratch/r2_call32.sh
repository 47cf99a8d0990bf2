#!/bin/bash
set -u
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -3
AB_TAG=c32 bash scratch/ab.sh prev pf
AB_TAG=c32m BENCH_ARGS="--workload marshak_f32_rw" bash scratch/ab.sh prev pf
} 2>&1 | tee gpurun_out/r2_call32.log
