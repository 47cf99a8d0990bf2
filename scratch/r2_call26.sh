#!/bin/bash
set -u
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_edge.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -4
bash scratch/launch_list.sh r2c26_f32 2>&1 | head -12
} 2>&1 | tee gpurun_out/r2_call26.log
