#!/bin/bash
# A/B of build variants on one box:  bash scratch/ab.sh name1 name2 ...   (variants/libimc_<name>.so; two rounds each, interleaved)
set -u
mkdir -p gpurun_out
run() { IMC_LIB=$PWD/variants/libimc_$1.so python bench.py --track refill --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1: seg/s %.4g  ms/step %.2f  kernel_ms %.2f  frac %.4f  sph %.1f' % (d['value'], d['ms_per_step'], d['tracking_kernel_ms_per_step'], d['roofline']['frac'], d['segments_per_history']))"; }
for round in 1 2; do for v in "$@"; do run $v; done; done 2>&1 | tee -a gpurun_out/ab_${AB_TAG:-x}.log
