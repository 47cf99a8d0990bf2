#!/bin/bash
# one-barrier compaction at 4 / 3 / 2 blocks per SM against HEAD~2 (dtab2) on the sparse-survivor decks; census tally with global RED instead of shared CAS
set -u
mkdir -p gpurun_out
{
AB_TAG=c48_f64 BENCH_ARGS="--workload crookedpipe_f64" bash scratch/ab.sh dtab2 cA cB cC
AB_TAG=c48_su32 BENCH_ARGS="--workload suolson_f32 --track auto --steps 10" bash scratch/ab.sh cA cA:IMC_CENSUS_SMEM=0
AB_TAG=c48_f32 bash scratch/ab.sh dtab2 cA
} 2>&1 | tee gpurun_out/r2_call48.log
