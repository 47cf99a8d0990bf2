#!/bin/bash
# ncu captures of the two kernels rewritten in the second session, at HEAD
set -u
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_census_tally -s 4 -c 1 -o gpurun_out/r2h_census -f python bench.py --workload suolson_f32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_census_ncu.log 2>&1
tail -1 gpurun_out/r2h_census_ncu.log | cut -c1-120
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_compact -s 4 -c 1 -o gpurun_out/r2h_compact -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_compact_ncu.log 2>&1
tail -1 gpurun_out/r2h_compact_ncu.log | cut -c1-120
