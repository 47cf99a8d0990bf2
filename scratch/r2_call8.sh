#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== parity subset on the y-major lib"
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -4
AB_TAG=c8 bash scratch/ab.sh u2 ym ym:IMC_CELL_ORDER=1 ymca ymca:IMC_CELL_ORDER=1
} 2>&1 | tee gpurun_out/r2_call8.log
