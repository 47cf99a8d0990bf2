#!/bin/bash
# lazy compaction + the clean / census kernels of call 43 (lazy) against HEAD's predecessor (dtab2); full GPU suite on the tree's library
set -u
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
AB_TAG=c44_su32 BENCH_ARGS="--workload suolson_f32 --track auto --steps 10" bash scratch/ab.sh dtab2 lazy
AB_TAG=c44_su64 BENCH_ARGS="--workload suolson_f64 --track auto --steps 10" bash scratch/ab.sh dtab2 lazy lazy
AB_TAG=c44_f64 BENCH_ARGS="--workload crookedpipe_f64" bash scratch/ab.sh dtab2 lazy
} 2>&1 | tee gpurun_out/r2_call44.log
