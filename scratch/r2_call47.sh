#!/bin/bash
# full GPU suite at HEAD; compaction A/B on the sparse-survivor deck; ncu capture of the census tally on Su-Olson
set -u
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
AB_TAG=c47_f64 BENCH_ARGS="--workload crookedpipe_f64" bash scratch/ab.sh dtab2 head47
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_census_tally -s 4 -c 1 -o gpurun_out/r2c47_census -f python bench.py --workload suolson_f32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c47_census_ncu.log 2>&1
tail -2 gpurun_out/r2c47_census_ncu.log | cut -c1-200
} 2>&1 | tee gpurun_out/r2_call47.log
