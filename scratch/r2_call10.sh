#!/bin/bash
set -u
mkdir -p gpurun_out
{
echo "== parity subset (chunked queue, y-fastest cell table)"
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -4
AB_TAG=c10 bash scratch/ab.sh ymca:IMC_CELL_ORDER=2,IMC_TALLY_ORDER=1 q q:IMC_QUEUE_MULT=1000003 q:IMC_QUEUE_MULT=7919 q:IMC_QUEUE_MULT=1000003,IMC_TALLY_ORDER=2
} 2>&1 | tee gpurun_out/r2_call10.log
