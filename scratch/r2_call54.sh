#!/bin/bash
set -u
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_edge.py -m gpu -q -k "census_tally_is or lazy" 2>&1 | tail -15
bash scratch/r2_sanitize2.sh
} 2>&1 | tee gpurun_out/r2_call54.log
