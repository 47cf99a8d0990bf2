#!/bin/bash
# compaction with 4096 particles per block + census tally with four particles per thread (clean4) against HEAD (dtab2)
set -u
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_fuzz.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
AB_TAG=c43_su32 BENCH_ARGS="--workload suolson_f32 --track auto" bash scratch/ab.sh dtab2 clean4
AB_TAG=c43_su64 BENCH_ARGS="--workload suolson_f64 --track auto" bash scratch/ab.sh dtab2 clean4
AB_TAG=c43_f32 bash scratch/ab.sh dtab2 clean4
IMC_LIB=$PWD/variants/libimc_clean4.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c43_su32_launches.csv python bench.py --workload suolson_f32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c43_su32_launches.log 2>&1
} 2>&1 | tee gpurun_out/r2_call43.log
