#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== parity subset on the new default lib"
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -4
AB_TAG=c6 bash scratch/ab.sh nof32 u2
} 2>&1 | tee gpurun_out/r2_call6.log
