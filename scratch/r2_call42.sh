#!/bin/bash
# does the ordinal-by-ordinal sharding of the new particles cost the tracking kernel its locality at N = 8?
set -u
mkdir -p gpurun_out
{
timeout 300 python scratch/stripe_probe.py 1 10 | tail -1
IMC_STRIPE=1 timeout 300 python scratch/stripe_probe.py 8 10 | tail -1
IMC_STRIPE=32 timeout 300 python scratch/stripe_probe.py 8 10 | tail -1
IMC_STRIPE=256 timeout 300 python scratch/stripe_probe.py 8 10 | tail -1
} 2>&1 | tee gpurun_out/r2_call42.log
